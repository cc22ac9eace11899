/* [X, convergence_error] = mc_admm(Htrue, OH, Omega, Imax, tau, rho)   drop-in for benchmark_algorithms/mc_admm.m:1 */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "mc_admm";
    gw_nargs(fn, nrhs, 6, nlhs, 2);
    int Mr = (int)mxGetM(prhs[1]), Mt = (int)mxGetN(prhs[1]);
    int imax = (int)gw_scalar(prhs[3], fn, "Imax");
    double tau = gw_scalar(prhs[4], fn, "tau"), rho = gw_scalar(prhs[5], fn, "rho");
    void *t0, *t1;
    const mxComplexDouble* Ht = gw_complex(prhs[0], fn, "Htrue", &t0);
    const mxComplexDouble* OH = gw_complex(prhs[1], fn, "OH", &t1);
    const double* om = gw_real(prhs[2], fn, "Omega");
    plhs[0] = mxCreateDoubleMatrix(Mr, Mt, mxCOMPLEX);
    mxArray* cv = nlhs >= 2 ? mxCreateDoubleMatrix(imax, 1, mxREAL) : NULL;
    long long ld = (long long)Mr * Mt;
    int rc = jstsp_mc_admm(gw_handle(fn), JSTSP_F64, JSTSP_HOST, Mr, Mt, 1, imax, cv ? Ht : NULL, ld, OH, ld, om, ld, &tau, &rho,
                           mxGetComplexDoubles(plhs[0]), ld, cv ? mxGetDoubles(cv) : NULL, imax);
    if (t0) mxFree(t0); if (t1) mxFree(t1);
    if (cv) plhs[1] = cv;
    gw_status(rc, fn);
}
