/* [Z, support, R] = jstsp_somp(A, Y, K)
 *   joint (MMV) OMP behind the +spx shim classes (jstsp19_b200/mex/matlab/+spx): stands in for
 *   spx.pursuit.joint.OrthogonalMatchingPursuit(A, K).solve(Y) of the reference's drivers
 *   (plot_errorVSsnr.m:116-118).  Z is size(A,2) x size(Y,2); support a 1 x n double row (1-based, pick order). */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "jstsp_somp";
    gw_nargs(fn, nrhs, 3, nlhs, 3);
    int N = (int)mxGetM(prhs[0]), D = (int)mxGetN(prhs[0]);
    int S = (int)mxGetN(prhs[1]);
    if ((int)mxGetM(prhs[1]) != N) mexErrMsgIdAndTxt("jstsp:size", "%s: size(Y,1) must equal size(A,1)", fn);
    int K = (int)gw_scalar(prhs[2], fn, "K");
    if (K < 1) mexErrMsgIdAndTxt("jstsp:size", "%s: K must be positive", fn);
    void *t0, *t1;
    const mxComplexDouble* A = gw_complex(prhs[0], fn, "A", &t0);
    const mxComplexDouble* Y = gw_complex(prhs[1], fn, "Y", &t1);
    plhs[0] = mxCreateDoubleMatrix(D, S, mxCOMPLEX);
    int* sup = (int*)mxMalloc(sizeof(int) * K);
    int nit = 0;
    mxArray* res = nlhs >= 3 ? mxCreateDoubleMatrix(N, S, mxCOMPLEX) : NULL;
    int rc = jstsp_somp(gw_handle(fn), JSTSP_F64, JSTSP_HOST, N, D, S, K, 1, A, 0, Y, (long long)N * S,
                        mxGetComplexDoubles(plhs[0]), (long long)D * S, sup, &nit, res ? mxGetComplexDoubles(res) : NULL, (long long)N * S, 0.0);
    if (nlhs >= 2 && rc == 0) {
        plhs[1] = mxCreateDoubleMatrix(1, nit, mxREAL);
        double* o = mxGetDoubles(plhs[1]);
        for (int t = 0; t < nit; ++t) o[t] = (double)sup[t];
    }
    if (res) plhs[2] = res;
    mxFree(sup); if (t0) mxFree(t0); if (t1) mxFree(t1);
    gw_status(rc, fn);
}
