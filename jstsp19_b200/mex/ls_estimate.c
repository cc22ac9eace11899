/* [S_ls, YpinvB] = ls_estimate(A, Y, B)
 *   S_ls   = pinv(A)*Y*pinv(B)     replaces the inline expression at plot_errorVSsnr.m:83 (plot_errorVSsnr_approx.m:61,67)
 *   YpinvB = Y*pinv(B)             the right-hand sides handed to the joint OMP at plot_errorVSsnr.m:117
 * A is N x G, Y is N x M, B is P x M (full rank). */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "ls_estimate";
    gw_nargs(fn, nrhs, 3, nlhs, 2);
    const int N = (int)mxGetM(prhs[0]), G = (int)mxGetN(prhs[0]), M = (int)mxGetN(prhs[1]), P = (int)mxGetM(prhs[2]);
    if ((int)mxGetM(prhs[1]) != N || (int)mxGetN(prhs[2]) != M) mexErrMsgIdAndTxt("jstsp:size", "%s: A is N x G, Y must be N x M and B P x M", fn);
    void *t0, *t1, *t2;
    const mxComplexDouble* A = gw_complex(prhs[0], fn, "A", &t0);
    const mxComplexDouble* Y = gw_complex(prhs[1], fn, "Y", &t1);
    const mxComplexDouble* B = gw_complex(prhs[2], fn, "B", &t2);
    plhs[0] = mxCreateDoubleMatrix(G, P, mxCOMPLEX);
    mxComplexDouble* YpB = NULL;
    if (nlhs > 1) { plhs[1] = mxCreateDoubleMatrix(N, P, mxCOMPLEX); YpB = mxGetComplexDoubles(plhs[1]); }
    int rc = jstsp_ls_estimate(gw_handle(fn), JSTSP_F64, JSTSP_HOST, N, M, G, P, 1, A, 0, B, 0, Y, 0, mxGetComplexDoubles(plhs[0]), 0, YpB, 0);
    if (t0) mxFree(t0); if (t1) mxFree(t1); if (t2) mxFree(t2);
    gw_status(rc, fn);
}
