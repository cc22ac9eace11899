/* [x_hat, indexSet, x_sel, residual] = OMP_kron(A, B, Y, m)
 *   == OMP(kron(B.', A), vec(Y), m) of benchmark_algorithms/OMP.m:1-32 on the operands the drivers build at
 *   plot_errorVSdelays.m:77-78, without forming the Kronecker dictionary (8192 x 262144 at BASELINE config 2).
 *   indexSet is a 1 x m cell of double scalars like OMP.m:13,17 (1-based linear index into the Gr x L*Gt unknown);
 *   x_hat is (Gr*L*Gt) x 1, x_sel the m coefficients in pick order, residual N x M. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "OMP_kron";
    gw_nargs(fn, nrhs, 4, nlhs, 4);
    int N = (int)mxGetM(prhs[0]), G = (int)mxGetN(prhs[0]);
    int P = (int)mxGetM(prhs[1]), M = (int)mxGetN(prhs[1]);
    if ((int)mxGetM(prhs[2]) != N || (int)mxGetN(prhs[2]) != M) mexErrMsgIdAndTxt("jstsp:size", "%s: Y must be size(A,1) x size(B,2)", fn);
    int m = (int)gw_scalar(prhs[3], fn, "m");
    void *t0, *t1, *t2;
    const mxComplexDouble* A = gw_complex(prhs[0], fn, "A", &t0);
    const mxComplexDouble* B = gw_complex(prhs[1], fn, "B", &t1);
    const mxComplexDouble* Y = gw_complex(prhs[2], fn, "Y", &t2);
    plhs[0] = mxCreateDoubleMatrix((size_t)G * P, 1, mxCOMPLEX);
    int* idx = (int*)mxMalloc(sizeof(int) * (m > 0 ? m : 1));
    mxArray* xs = nlhs >= 3 ? mxCreateDoubleMatrix(m, 1, mxCOMPLEX) : NULL;
    mxArray* res = nlhs >= 4 ? mxCreateDoubleMatrix(N, M, mxCOMPLEX) : NULL;
    int amb = 0;
    int rc = jstsp_omp_kron(gw_handle(fn), JSTSP_F64, JSTSP_HOST, N, M, G, P, m, 1, A, 0, B, 0, Y, (long long)N * M,
                            mxGetComplexDoubles(plhs[0]), (long long)G * P, idx, xs ? mxGetComplexDoubles(xs) : NULL,
                            res ? mxGetComplexDoubles(res) : NULL, (long long)N * M, &amb, 1e-10);
    if (nlhs >= 2 && rc == 0) {
        plhs[1] = mxCreateCellMatrix(1, m);
        for (int t = 0; t < m; ++t) mxSetCell(plhs[1], t, mxCreateDoubleScalar((double)idx[t]));
    }
    if (xs) plhs[2] = xs;
    if (res) plhs[3] = res;
    mxFree(idx); if (t0) mxFree(t0); if (t1) mxFree(t1); if (t2) mxFree(t2);
    gw_status(rc, fn);
    if (amb > 0) mexWarnMsgIdAndTxt("jstsp:omp:neartie", "%s: %d selection(s) were decided by a margin below 1e-10", fn, amb);
}
