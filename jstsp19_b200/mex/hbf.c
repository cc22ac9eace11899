/* [Y_conventional_hbf, W_c, Psi_bar, Y] = hbf(H, N, Psi_i, T, Lr, W)   drop-in for basic_system_functions/hbf.m:1 */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "hbf";
    gw_nargs(fn, nrhs, 6, nlhs, 4);
    const mwSize* hd = mxGetDimensions(prhs[0]);
    jstsp_meas_desc d; memset(&d, 0, sizeof d);
    d.Nr = (int)hd[0]; d.Nt = (int)hd[1]; d.L = mxGetNumberOfDimensions(prhs[0]) > 2 ? (int)hd[2] : 1;
    d.T = (int)gw_scalar(prhs[3], fn, "T"); d.Wc = (int)gw_scalar(prhs[4], fn, "Lr"); d.Lr = 0;
    d.psi_mode = 0; d.Tp = (int)mxGetM(prhs[2]); d.batch = 1;
    if ((int)mxGetM(prhs[1]) != d.Nr || (int)mxGetN(prhs[1]) != d.T) mexErrMsgIdAndTxt("jstsp:size", "%s: N must be Nr x T", fn);
    void *t0, *t1, *t2, *t3;
    const mxComplexDouble* H = gw_complex(prhs[0], fn, "H", &t0);
    const mxComplexDouble* N = gw_complex(prhs[1], fn, "N", &t1);
    const mxComplexDouble* Psi = gw_complex(prhs[2], fn, "Psi_i", &t2);
    const mxComplexDouble* W = gw_complex(prhs[5], fn, "W", &t3);
    mwSize dp[3] = {(mwSize)d.Nt, (mwSize)d.T, (mwSize)d.L};
    mxArray* o[4];
    o[0] = mxCreateDoubleMatrix(d.Wc, d.T, mxCOMPLEX);
    o[1] = mxCreateDoubleMatrix(d.Nr, d.Wc, mxCOMPLEX);
    o[2] = mxCreateNumericArray(3, dp, mxDOUBLE_CLASS, mxCOMPLEX);
    o[3] = mxCreateDoubleMatrix(d.Nr, d.T, mxCOMPLEX);
    int rc = jstsp_measure(gw_handle(fn), &d, JSTSP_F64, JSTSP_HOST, H, N, Psi, W, NULL, mxGetComplexDoubles(o[0]), mxGetComplexDoubles(o[1]),
                           mxGetComplexDoubles(o[2]), NULL, mxGetComplexDoubles(o[3]));
    if (t0) mxFree(t0); if (t1) mxFree(t1); if (t2) mxFree(t2); if (t3) mxFree(t3);
    int nout = nlhs > 1 ? nlhs : 1;
    for (int k = 0; k < 4; ++k) { if (k < nout) plhs[k] = o[k]; else mxDestroyArray(o[k]); }
    gw_status(rc, fn);
}
