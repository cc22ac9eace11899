/* X = svt(Y, tau)   drop-in for benchmark_algorithms/svt.m:1 */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "svt";
    gw_nargs(fn, nrhs, 2, nlhs, 1);
    int Mr = (int)mxGetM(prhs[0]), Mt = (int)mxGetN(prhs[0]);
    double tau = gw_scalar(prhs[1], fn, "tau");
    void* t0; const mxComplexDouble* Y = gw_complex(prhs[0], fn, "Y", &t0);
    plhs[0] = mxCreateDoubleMatrix(Mr, Mt, mxCOMPLEX);
    int rc = jstsp_svt(gw_handle(fn), JSTSP_F64, JSTSP_HOST, Mr, Mt, 1, Y, (long long)Mr * Mt, &tau, mxGetComplexDoubles(plhs[0]), (long long)Mr * Mt);
    if (t0) mxFree(t0);
    gw_status(rc, fn);
}
