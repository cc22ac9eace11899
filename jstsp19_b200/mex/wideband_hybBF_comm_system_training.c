/* [Y_proposed_hbf, Y_conventional_hbf, W_tilde, Psi_bar, Omega, Lr] = wideband_hybBF_comm_system_training(H, T, snr, subSamplingRatio)
 * drop-in for basic_system_functions/wideband_hybBF_comm_system_training.m:1.  All randomness comes from MATLAB's own
 * generator through mexCallMATLAB in the reference's order: randn(Nr,T) twice (.m:16), two randn(1,T) per transmit
 * antenna (.m:20), then T x randperm(Nr) (.m:49).  The synthesis R = sum_l H_l Psi_bar_l + N, the DFT combining and the
 * masking (.m:24-56) run in the library (jstsp_measure, psi_mode 1 = Toeplitz rows built from the pilot rows). */
#include <math.h>
#include "gateway_common.h"
static mxArray* call_randn(const char* fn, int r, int c) {
    mxArray* o[1] = {NULL};
    mxArray* in[2] = {mxCreateDoubleScalar((double)r), mxCreateDoubleScalar((double)c)};
    int rc = mexCallMATLAB(1, o, 2, in, "randn");
    mxDestroyArray(in[0]); mxDestroyArray(in[1]);
    if (rc != 0 || !o[0] || (int)mxGetNumberOfElements(o[0]) != r * c) mexErrMsgIdAndTxt("jstsp:rng", "%s: randn(%d,%d) failed", fn, r, c);
    return o[0];
}
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "wideband_hybBF_comm_system_training";
    gw_nargs(fn, nrhs, 4, nlhs, 6);
    const mwSize* hd = mxGetDimensions(prhs[0]);
    jstsp_meas_desc d; memset(&d, 0, sizeof d);
    d.Nr = (int)hd[0]; d.Nt = (int)hd[1]; d.L = mxGetNumberOfDimensions(prhs[0]) > 2 ? (int)hd[2] : 1;
    d.T = (int)gw_scalar(prhs[1], fn, "T");
    const double snr = gw_scalar(prhs[2], fn, "snr"), ratio = gw_scalar(prhs[3], fn, "subSamplingRatio");
    const int Nr = d.Nr, Nt = d.Nt, T = d.T;
    if (Nr < 1 || Nt < 1 || T < 1) mexErrMsgIdAndTxt("jstsp:size", "%s: H must be Nr x Nt x L and T positive", fn);
    d.Lr = (int)floor(ratio * Nr + 0.5);                     /* round(): half away from zero, ratio*Nr >= 0 (.m:5) */
    if (d.Lr < 0 || d.Lr > Nr) mexErrMsgIdAndTxt("jstsp:size", "%s: round(subSamplingRatio*Nr) must lie in [0, Nr]", fn);
    d.Wc = Nr; d.psi_mode = 1; d.Tp = T; d.batch = 1;
    /* N = sqrt(snr/2)*(randn(Nr,T) + 1j*randn(Nr,T))   (.m:16) */
    mxComplexDouble* N = (mxComplexDouble*)mxMalloc(sizeof(mxComplexDouble) * (size_t)Nr * T);
    {
        mxArray* re = call_randn(fn, Nr, T); mxArray* im = call_randn(fn, Nr, T);
        const double *pr = mxGetDoubles(re), *pi = mxGetDoubles(im), sc = sqrt(snr / 2.0);
        for (size_t k = 0; k < (size_t)Nr * T; ++k) { N[k].real = sc * pr[k]; N[k].imag = sc * pi[k]; }
        mxDestroyArray(re); mxDestroyArray(im);
    }
    /* s_k = 1/sqrt(2)*(randn(1,T)+1j*randn(1,T)); pilot matrix row k = s_k   (.m:19-22) */
    mxComplexDouble* S = (mxComplexDouble*)mxMalloc(sizeof(mxComplexDouble) * (size_t)Nt * T);
    for (int k = 0; k < Nt; ++k) {
        mxArray* re = call_randn(fn, 1, T); mxArray* im = call_randn(fn, 1, T);
        const double *pr = mxGetDoubles(re), *pi = mxGetDoubles(im), sc = 1.0 / sqrt(2.0);
        for (int t = 0; t < T; ++t) { S[k + (size_t)Nt * t].real = sc * pr[t]; S[k + (size_t)Nt * t].imag = sc * pi[t]; }
        mxDestroyArray(re); mxDestroyArray(im);
    }
    /* indices = randperm(Nr) per training instant   (.m:48-49) */
    int* perm = (int*)mxMalloc(sizeof(int) * (size_t)T * Nr);
    for (int t = 0; t < T; ++t) {
        mxArray* o[1] = {NULL}; mxArray* in[1] = {mxCreateDoubleScalar((double)Nr)};
        if (mexCallMATLAB(1, o, 1, in, "randperm") != 0 || !o[0]) mexErrMsgIdAndTxt("jstsp:rng", "%s: randperm failed", fn);
        const double* pv = mxGetDoubles(o[0]);
        for (int k = 0; k < Nr; ++k) perm[(size_t)t * Nr + k] = (int)pv[k];
        mxDestroyArray(o[0]); mxDestroyArray(in[0]);
    }
    /* W_tilde = 1/sqrt(Nr)*fft(eye(Nr))   (.m:10) */
    mxArray* o[6];
    o[2] = mxCreateDoubleMatrix(Nr, Nr, mxCOMPLEX);
    {
        mxComplexDouble* W = mxGetComplexDoubles(o[2]);
        const double two_pi = 6.283185307179586476925286766559, sc = 1.0 / sqrt((double)Nr);
        for (int j = 0; j < Nr; ++j)
            for (int i = 0; i < Nr; ++i) {
                const double ph = -two_pi * (double)(((long long)i * j) % Nr) / (double)Nr;
                W[i + (size_t)Nr * j].real = sc * cos(ph); W[i + (size_t)Nr * j].imag = sc * sin(ph);
            }
    }
    void* t0;
    const mxComplexDouble* H = gw_complex(prhs[0], fn, "H", &t0);
    mwSize dp[3] = {(mwSize)Nt, (mwSize)T, (mwSize)d.L};
    o[0] = mxCreateDoubleMatrix(Nr, T, mxCOMPLEX);
    o[1] = mxCreateDoubleMatrix(Nr, T, mxCOMPLEX);
    o[3] = mxCreateNumericArray(3, dp, mxDOUBLE_CLASS, mxCOMPLEX);
    o[4] = mxCreateDoubleMatrix(Nr, T, mxREAL);
    o[5] = mxCreateDoubleScalar((double)d.Lr);
    jstsp_handle* h = gw_handle(fn);
    int rc = jstsp_measure(h, &d, JSTSP_F64, JSTSP_HOST, H, N, S, mxGetComplexDoubles(o[2]), perm,
                           mxGetComplexDoubles(o[0]), NULL, mxGetComplexDoubles(o[3]), mxGetDoubles(o[4]), NULL);        /* .m:53 */
    if (rc == 0 && nlhs >= 2)
        rc = jstsp_measure(h, &d, JSTSP_F64, JSTSP_HOST, H, N, S, mxGetComplexDoubles(o[2]), NULL,
                           mxGetComplexDoubles(o[1]), NULL, NULL, NULL, NULL);                                              /* .m:56 */
    mxFree(N); mxFree(S); mxFree(perm); if (t0) mxFree(t0);
    int nout = nlhs > 1 ? nlhs : 1;
    for (int k = 0; k < 6; ++k) { if (k < nout) plhs[k] = o[k]; else mxDestroyArray(o[k]); }
    gw_status(rc, fn);
}
