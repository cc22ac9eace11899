/* [x_hat, indexSet, v, targetMatrix] = OMP(A, v, m, snr)   drop-in for benchmark_algorithms/OMP.m:1
 * indexSet is a 1 x m cell of double scalars (OMP.m:13,17); v is echoed; snr is unused by the reference. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "OMP";
    if (nrhs != 3 && nrhs != 4) mexErrMsgIdAndTxt("jstsp:nargin", "%s: expected 3 or 4 input arguments, got %d", fn, nrhs);
    if (nlhs > 4) mexErrMsgIdAndTxt("jstsp:nargout", "%s: at most 4 output arguments", fn);
    int measures = (int)mxGetM(prhs[0]), size_d = (int)mxGetN(prhs[0]);
    if ((int)mxGetNumberOfElements(prhs[1]) != measures) mexErrMsgIdAndTxt("jstsp:size", "%s: length(v) must equal size(A,1)", fn);
    int m = (int)gw_scalar(prhs[2], fn, "m");
    void *t0, *t1;
    const mxComplexDouble* A = gw_complex(prhs[0], fn, "A", &t0);
    const mxComplexDouble* v = gw_complex(prhs[1], fn, "v", &t1);
    plhs[0] = mxCreateDoubleMatrix(size_d, 1, mxCOMPLEX);
    int* idx = (int*)mxMalloc(sizeof(int) * (m > 0 ? m : 1));
    mxArray* tgt = nlhs >= 4 ? mxCreateDoubleMatrix(measures, m, mxCOMPLEX) : NULL;
    int amb = 0;
    int rc = jstsp_omp(gw_handle(fn), JSTSP_F64, JSTSP_HOST, measures, size_d, m, 1, A, 0, v, measures, mxGetComplexDoubles(plhs[0]), size_d,
                       idx, tgt ? mxGetComplexDoubles(tgt) : NULL, &amb, 1e-10);
    if (nlhs >= 2 && rc == 0) {
        plhs[1] = mxCreateCellMatrix(1, m);
        for (int t = 0; t < m; ++t) mxSetCell(plhs[1], t, mxCreateDoubleScalar((double)idx[t]));
    }
    if (nlhs >= 3 && rc == 0) {                                   /* echo v (OMP.m:1 returns its own input) */
        plhs[2] = mxCreateDoubleMatrix(mxGetM(prhs[1]), mxGetN(prhs[1]), mxCOMPLEX);
        memcpy(mxGetComplexDoubles(plhs[2]), v, sizeof(mxComplexDouble) * measures);
    }
    if (tgt) plhs[3] = tgt;
    mxFree(idx); if (t0) mxFree(t0); if (t1) mxFree(t1);
    gw_status(rc, fn);
    if (amb > 0) mexWarnMsgIdAndTxt("jstsp:omp:neartie", "%s: %d selection(s) were decided by a margin below 1e-10", fn, amb);
}
