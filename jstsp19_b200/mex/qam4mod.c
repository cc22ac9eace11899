/* symbols = qam4mod(input, mode, N)   drop-in for basic_system_functions/qam4mod.m:1
 * 'mod': N x 1 symbols; the draw is MATLAB's own randsrc(N,1,alphabet) (qam4mod.m:8).  'demod': hard decision of `input` (:12-31). */
#include <math.h>
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "qam4mod";
    if (nrhs != 2 && nrhs != 3) mexErrMsgIdAndTxt("jstsp:nargin", "%s: expected 2 or 3 input arguments, got %d", fn, nrhs);
    if (nlhs > 1) mexErrMsgIdAndTxt("jstsp:nargout", "%s: at most 1 output argument", fn);
    char mode[16] = "";
    if (!mxIsChar(prhs[1]) || mxGetString(prhs[1], mode, sizeof mode)) mexErrMsgIdAndTxt("jstsp:type", "%s: mode must be a char array", fn);
    if (strcmp(mode, "mod") == 0) {
        if (nrhs != 3) mexErrMsgIdAndTxt("jstsp:nargin", "%s: 'mod' needs N", fn);
        const int N = (int)gw_scalar(prhs[2], fn, "N");
        if (N < 1) mexErrMsgIdAndTxt("jstsp:size", "%s: N must be positive", fn);
        const double a = 1.0 / sqrt(2.0);
        mxArray* in[3] = {mxCreateDoubleScalar(N), mxCreateDoubleScalar(1), mxCreateDoubleMatrix(1, 4, mxCOMPLEX)};
        mxComplexDouble* al = mxGetComplexDoubles(in[2]);
        al[0].real = a; al[0].imag = a; al[1].real = -a; al[1].imag = a; al[2].real = a; al[2].imag = -a; al[3].real = -a; al[3].imag = -a;
        mxArray* o[1] = {NULL};
        if (mexCallMATLAB(1, o, 3, in, "randsrc") != 0 || !o[0]) mexErrMsgIdAndTxt("jstsp:rng", "%s: randsrc failed", fn);
        void* tmp; const mxComplexDouble* v = gw_complex(o[0], fn, "randsrc output", &tmp);
        int* draws = (int*)mxMalloc(sizeof(int) * (size_t)N);
        for (int k = 0; k < N; ++k) draws[k] = (v[k].real < 0 ? 1 : 0) | (v[k].imag < 0 ? 2 : 0);
        if (tmp) mxFree(tmp);
        mxDestroyArray(o[0]); for (int k = 0; k < 3; ++k) mxDestroyArray(in[k]);
        plhs[0] = mxCreateDoubleMatrix(N, 1, mxCOMPLEX);
        int rc = jstsp_qam4mod(gw_handle(fn), JSTSP_F64, JSTSP_HOST, 0, N, draws, NULL, mxGetComplexDoubles(plhs[0]));
        mxFree(draws);
        gw_status(rc, fn);
    } else if (strcmp(mode, "demod") == 0) {
        void* t0; const mxComplexDouble* x = gw_complex(prhs[0], fn, "input", &t0);
        const size_t n = mxGetNumberOfElements(prhs[0]);
        plhs[0] = mxCreateDoubleMatrix(mxGetM(prhs[0]), mxGetN(prhs[0]), mxCOMPLEX);
        int rc = n ? jstsp_qam4mod(gw_handle(fn), JSTSP_F64, JSTSP_HOST, 1, (long long)n, NULL, x, mxGetComplexDoubles(plhs[0])) : 0;
        if (t0) mxFree(t0);
        gw_status(rc, fn);
    } else {
        mexErrMsgIdAndTxt("jstsp:type", "%s: mode must be 'mod' or 'demod'", fn);
    }
}
