/* [S, Y, convergence_error] = proposed_algorithm_pilots(subY, Omega, A, Dt, pilots, L, Imax, tau_Y, tau_S, rho, type [, indx_S])
 * proposed_algorithm_psi fed the pilot sequences themselves: pilots is Nt x M with row k = s_k, the vector the drivers hand to
 * toeplitz() (plot_errorVSsnr.m:63-67:  s = qam4mod(...); Psi_i(:,:,k) = toeplitz(s);  ->  also keep  pilots(k,:) = s).
 * Psi_bar(k,:,l) = Psi_i(l,:,k) (proposed_hbf.m:15-18) is expanded on the device, so L times fewer dictionary bytes travel. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "proposed_algorithm_pilots";
    if (nrhs != 11 && nrhs != 12) mexErrMsgIdAndTxt("jstsp:nargin", "%s: 11 or 12 inputs expected, got %d", fn, nrhs);
    gw_nargs(fn, nrhs, nrhs, nlhs, 3);
    jstsp_admm_desc d; memset(&d, 0, sizeof d);
    d.N = (int)mxGetM(prhs[0]); d.M = (int)mxGetN(prhs[0]);
    d.G = (int)mxGetN(prhs[2]);
    const int Nt = (int)mxGetM(prhs[3]), Gt = (int)mxGetN(prhs[3]);
    const int L = (int)gw_scalar(prhs[5], fn, "L");
    d.P = L * Gt;
    if ((int)mxGetM(prhs[1]) != d.N || (int)mxGetN(prhs[1]) != d.M || (int)mxGetM(prhs[2]) != d.N || (int)mxGetM(prhs[4]) != Nt || (int)mxGetN(prhs[4]) != d.M || L < 1)
        mexErrMsgIdAndTxt("jstsp:size", "%s: need subY N x M, Omega N x M, A N x G, Dt Nt x Gt, pilots Nt x M, L >= 1", fn);
    d.imax = (int)gw_scalar(prhs[6], fn, "Imax");
    double tauY = gw_scalar(prhs[7], fn, "tau_Y"), tauS = gw_scalar(prhs[8], fn, "tau_S"), rho = gw_scalar(prhs[9], fn, "rho");
    char type[32] = "";
    if (!mxIsChar(prhs[10]) || mxGetString(prhs[10], type, sizeof type)) type[0] = 0;
    d.type = strcmp(type, "approximate") == 0 ? JSTSP_APPROXIMATE : JSTSP_STD;      /* .m:23-30: anything else = exact LS */
    d.batch = 1; d.ld_subY = (long long)d.N * d.M; d.ld_omega = d.ld_subY; d.ld_S = (long long)d.G * d.P; d.ld_Y = d.ld_subY; d.ld_conv = 3LL * d.imax;
    int* indx = NULL;
    if (nrhs == 12) {                                              /* 1-based doubles -> int32 (proposed_algorithm_angles.m:36) */
        const double* ix = gw_real(prhs[11], fn, "indx_S");
        d.n_indx = (int)mxGetNumberOfElements(prhs[11]);
        indx = (int*)mxMalloc(sizeof(int) * (size_t)(d.n_indx > 0 ? d.n_indx : 1));
        for (int k = 0; k < d.n_indx; ++k) indx[k] = (int)ix[k];
    }
    void *t0, *t2, *t3, *t4;
    const mxComplexDouble* subY = gw_complex(prhs[0], fn, "subY", &t0);
    const double* omega = gw_real(prhs[1], fn, "Omega");
    const mxComplexDouble* A = gw_complex(prhs[2], fn, "A", &t2);
    const mxComplexDouble* Dt = gw_complex(prhs[3], fn, "Dt", &t3);
    const mxComplexDouble* Psi = gw_complex(prhs[4], fn, "pilots", &t4);
    plhs[0] = mxCreateDoubleMatrix(d.G, d.P, mxCOMPLEX);
    mxArray* Y = nlhs >= 2 ? mxCreateDoubleMatrix(d.N, d.M, mxCOMPLEX) : NULL;
    mxArray* cv = nlhs >= 3 ? mxCreateDoubleMatrix(d.imax, 3, mxREAL) : NULL;   /* diagnostics only when asked for */
    int rc = jstsp_proposed_algorithm_pilots(gw_handle(fn), &d, JSTSP_F64, JSTSP_HOST, subY, omega, indx, A, Dt, 0, Psi, 0, Nt, L, &tauY, &tauS, &rho,
                                          mxGetComplexDoubles(plhs[0]), Y ? mxGetComplexDoubles(Y) : NULL, cv ? mxGetDoubles(cv) : NULL);
    if (t0) mxFree(t0); if (t2) mxFree(t2); if (t3) mxFree(t3); if (t4) mxFree(t4); if (indx) mxFree(indx);
    if (Y) plhs[1] = Y; if (cv) plhs[2] = cv;
    gw_status(rc, fn);
}
