/* X = mc_svt(OH, Omega, Imax, tau, rho)   drop-in for benchmark_algorithms/mc_svt.m:1 */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "mc_svt";
    gw_nargs(fn, nrhs, 5, nlhs, 1);
    int Mr = (int)mxGetM(prhs[0]), Mt = (int)mxGetN(prhs[0]);
    if ((int)mxGetM(prhs[1]) != Mr || (int)mxGetN(prhs[1]) != Mt) mexErrMsgIdAndTxt("jstsp:size", "%s: Omega must match OH", fn);
    int imax = (int)gw_scalar(prhs[2], fn, "Imax");
    double tau = gw_scalar(prhs[3], fn, "tau"), rho = gw_scalar(prhs[4], fn, "rho");
    void* t0; const mxComplexDouble* OH = gw_complex(prhs[0], fn, "OH", &t0);
    const double* om = gw_real(prhs[1], fn, "Omega");
    plhs[0] = mxCreateDoubleMatrix(Mr, Mt, mxCOMPLEX);
    long long ld = (long long)Mr * Mt;
    int rc = jstsp_mc_svt(gw_handle(fn), JSTSP_F64, JSTSP_HOST, Mr, Mt, 1, imax, OH, ld, om, ld, &tau, &rho, mxGetComplexDoubles(plhs[0]), ld);
    if (t0) mxFree(t0);
    gw_status(rc, fn);
}
