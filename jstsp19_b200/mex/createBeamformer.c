/* B = createBeamformer(N, beamformer_type)   drop-in for basic_system_functions/createBeamformer.m:1
 * The random codebooks draw through MATLAB exactly where the reference does: randsrc(N,N,[1 -1 1j -1j]) for 'rand' (:8),
 * randi(32,1,N) for 'rand_ps' (:10-11), so a seeded session reproduces the reference's stream. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "createBeamformer";
    gw_nargs(fn, nrhs, 2, nlhs, 1);
    const int N = (int)gw_scalar(prhs[0], fn, "N");
    char type[32] = "";
    if (!mxIsChar(prhs[1]) || mxGetString(prhs[1], type, sizeof type)) mexErrMsgIdAndTxt("jstsp:type", "%s: beamformer_type must be a char array", fn);
    static const char* names[] = {"fft", "rand", "rand_ps", "ps", "ZC", "quantized_4", "quantized"};
    int t = -1;
    for (int k = 0; k < 7; ++k) if (strcmp(type, names[k]) == 0) t = k;
    if (t < 0 || N < 1) mexErrMsgIdAndTxt("jstsp:size", "%s: unknown beamformer_type '%s' (the reference's switch has no otherwise branch) or N < 1", fn, type);
    int* draws = NULL;
    if (t == JSTSP_BF_RAND) {
        mxArray* in[3] = {mxCreateDoubleScalar(N), mxCreateDoubleScalar(N), mxCreateDoubleMatrix(1, 4, mxCOMPLEX)};
        mxComplexDouble* al = mxGetComplexDoubles(in[2]);
        al[0].real = 1; al[0].imag = 0; al[1].real = -1; al[1].imag = 0; al[2].real = 0; al[2].imag = 1; al[3].real = 0; al[3].imag = -1;
        mxArray* o[1] = {NULL};
        if (mexCallMATLAB(1, o, 3, in, "randsrc") != 0 || !o[0]) mexErrMsgIdAndTxt("jstsp:rng", "%s: randsrc failed", fn);
        void* tmp; const mxComplexDouble* v = gw_complex(o[0], fn, "randsrc output", &tmp);
        draws = (int*)mxMalloc(sizeof(int) * (size_t)N * N);
        for (size_t k = 0; k < (size_t)N * N; ++k) draws[k] = v[k].real > 0.5 ? 0 : v[k].real < -0.5 ? 1 : v[k].imag > 0.5 ? 2 : 3;
        if (tmp) mxFree(tmp);
        mxDestroyArray(o[0]); for (int k = 0; k < 3; ++k) mxDestroyArray(in[k]);
    } else if (t == JSTSP_BF_RAND_PS) {
        mxArray* in[3] = {mxCreateDoubleScalar(32), mxCreateDoubleScalar(1), mxCreateDoubleScalar(N)};
        mxArray* o[1] = {NULL};
        if (mexCallMATLAB(1, o, 3, in, "randi") != 0 || !o[0]) mexErrMsgIdAndTxt("jstsp:rng", "%s: randi failed", fn);
        const double* v = gw_real(o[0], fn, "randi output");
        draws = (int*)mxMalloc(sizeof(int) * (size_t)N);
        for (int k = 0; k < N; ++k) draws[k] = (int)v[k];
        mxDestroyArray(o[0]); for (int k = 0; k < 3; ++k) mxDestroyArray(in[k]);
    }
    plhs[0] = mxCreateDoubleMatrix(N, N, mxCOMPLEX);
    int rc = jstsp_create_beamformer(gw_handle(fn), JSTSP_F64, JSTSP_HOST, N, t, draws, mxGetComplexDoubles(plhs[0]));
    if (draws) mxFree(draws);
    gw_status(rc, fn);
}
