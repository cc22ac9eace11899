/* [Y_proposed_hbf, W_e, Psi_bar, Omega, Y] = proposed_hbf(H, N, Psi_i, T, Lr_e, Lr, W)
 * drop-in for basic_system_functions/proposed_hbf.m:1.  The T randperm(Lr_e) draws of :38 come from MATLAB's own
 * generator (mexCallMATLAB) in the reference's order. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "proposed_hbf";
    gw_nargs(fn, nrhs, 7, nlhs, 5);
    const mwSize* hd = mxGetDimensions(prhs[0]);
    jstsp_meas_desc d; memset(&d, 0, sizeof d);
    d.Nr = (int)hd[0]; d.Nt = (int)hd[1]; d.L = mxGetNumberOfDimensions(prhs[0]) > 2 ? (int)hd[2] : 1;
    d.T = (int)gw_scalar(prhs[3], fn, "T"); d.Wc = (int)gw_scalar(prhs[4], fn, "Lr_e"); d.Lr = (int)gw_scalar(prhs[5], fn, "Lr");
    d.psi_mode = 0; d.Tp = (int)mxGetM(prhs[2]); d.batch = 1;
    if ((int)mxGetM(prhs[1]) != d.Nr || (int)mxGetN(prhs[1]) != d.T) mexErrMsgIdAndTxt("jstsp:size", "%s: N must be Nr x T", fn);
    if ((int)mxGetM(prhs[6]) != d.Nr || (int)mxGetN(prhs[6]) < d.Wc) mexErrMsgIdAndTxt("jstsp:size", "%s: W must be Nr x (>= Lr_e)", fn);
    int* perm = (int*)mxMalloc(sizeof(int) * (size_t)d.T * d.Wc);
    for (int t = 0; t < d.T; ++t) {                                                  /* indices = randperm(Lr_e)  (.m:38) */
        mxArray* o[1] = {NULL}; mxArray* in[1] = {mxCreateDoubleScalar((double)d.Wc)};
        if (mexCallMATLAB(1, o, 1, in, "randperm") != 0 || !o[0]) mexErrMsgIdAndTxt("jstsp:rng", "%s: randperm failed", fn);
        const double* pv = mxGetDoubles(o[0]);
        for (int k = 0; k < d.Wc; ++k) perm[(size_t)t * d.Wc + k] = (int)pv[k];
        mxDestroyArray(o[0]); mxDestroyArray(in[0]);
    }
    void *t0, *t1, *t2, *t3;
    const mxComplexDouble* H = gw_complex(prhs[0], fn, "H", &t0);
    const mxComplexDouble* N = gw_complex(prhs[1], fn, "N", &t1);
    const mxComplexDouble* Psi = gw_complex(prhs[2], fn, "Psi_i", &t2);
    const mxComplexDouble* W = gw_complex(prhs[6], fn, "W", &t3);
    mwSize dp[3] = {(mwSize)d.Nt, (mwSize)d.T, (mwSize)d.L};
    mxArray* o[5];
    o[0] = mxCreateDoubleMatrix(d.Wc, d.T, mxCOMPLEX);
    o[1] = mxCreateDoubleMatrix(d.Nr, d.Wc, mxCOMPLEX);
    o[2] = mxCreateNumericArray(3, dp, mxDOUBLE_CLASS, mxCOMPLEX);
    o[3] = mxCreateDoubleMatrix(d.Wc, d.T, mxREAL);
    o[4] = mxCreateDoubleMatrix(d.Nr, d.T, mxCOMPLEX);
    int rc = jstsp_measure(gw_handle(fn), &d, JSTSP_F64, JSTSP_HOST, H, N, Psi, W, perm, mxGetComplexDoubles(o[0]), mxGetComplexDoubles(o[1]),
                           mxGetComplexDoubles(o[2]), mxGetDoubles(o[3]), mxGetComplexDoubles(o[4]));
    mxFree(perm); if (t0) mxFree(t0); if (t1) mxFree(t1); if (t2) mxFree(t2); if (t3) mxFree(t3);
    int nout = nlhs > 1 ? nlhs : 1;
    for (int k = 0; k < 5; ++k) { if (k < nout) plhs[k] = o[k]; else mxDestroyArray(o[k]); }
    gw_status(rc, fn);
}
