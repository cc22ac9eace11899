/* [S, Y, convergence_error] = proposed_algorithm(subY, Omega, A, B, Imax, tau_Y, tau_S, rho, type)
 * drop-in for basic_system_functions/proposed_algorithm.m:1 */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "proposed_algorithm";
    gw_nargs(fn, nrhs, 9, nlhs, 3);
    jstsp_admm_desc d; memset(&d, 0, sizeof d);
    d.N = (int)mxGetM(prhs[0]); d.M = (int)mxGetN(prhs[0]);
    d.G = (int)mxGetN(prhs[2]); d.P = (int)mxGetM(prhs[3]);
    if ((int)mxGetM(prhs[1]) != d.N || (int)mxGetN(prhs[1]) != d.M || (int)mxGetM(prhs[2]) != d.N || (int)mxGetN(prhs[3]) != d.M)
        mexErrMsgIdAndTxt("jstsp:size", "%s: need subY N x M, Omega N x M, A N x G, B P x M", fn);
    d.imax = (int)gw_scalar(prhs[4], fn, "Imax");
    double tauY = gw_scalar(prhs[5], fn, "tau_Y"), tauS = gw_scalar(prhs[6], fn, "tau_S"), rho = gw_scalar(prhs[7], fn, "rho");
    char type[32] = "";
    if (!mxIsChar(prhs[8]) || mxGetString(prhs[8], type, sizeof type)) type[0] = 0;
    d.type = strcmp(type, "approximate") == 0 ? JSTSP_APPROXIMATE : JSTSP_STD;      /* .m:23-30: anything else = exact LS */
    d.batch = 1; d.ld_subY = (long long)d.N * d.M; d.ld_omega = d.ld_subY; d.ld_S = (long long)d.G * d.P; d.ld_Y = d.ld_subY; d.ld_conv = 3LL * d.imax;
    void *t0, *t2, *t3;
    const mxComplexDouble* subY = gw_complex(prhs[0], fn, "subY", &t0);
    const double* omega = gw_real(prhs[1], fn, "Omega");
    const mxComplexDouble* A = gw_complex(prhs[2], fn, "A", &t2);
    const mxComplexDouble* B = gw_complex(prhs[3], fn, "B", &t3);
    plhs[0] = mxCreateDoubleMatrix(d.G, d.P, mxCOMPLEX);
    mxArray* Y = nlhs >= 2 ? mxCreateDoubleMatrix(d.N, d.M, mxCOMPLEX) : NULL;
    mxArray* cv = nlhs >= 3 ? mxCreateDoubleMatrix(d.imax, 3, mxREAL) : NULL;   /* diagnostics only when asked for */
    int rc = jstsp_proposed_algorithm(gw_handle(fn), &d, JSTSP_F64, JSTSP_HOST, subY, omega, A, B, &tauY, &tauS, &rho,
                                      mxGetComplexDoubles(plhs[0]), Y ? mxGetComplexDoubles(Y) : NULL, cv ? mxGetDoubles(cv) : NULL);
    if (t0) mxFree(t0); if (t2) mxFree(t2); if (t3) mxFree(t3);
    if (Y) plhs[1] = Y; if (cv) plhs[2] = cv;
    gw_status(rc, fn);
}
