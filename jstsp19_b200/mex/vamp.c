/* x = vamp(y, A, sigma, L)   drop-in for benchmark_algorithms/vamp.m:1
 * The spectral decomposition is taken from MATLAB's own svd, exactly where vamp.m:32 calls it (on the complex
 * A instead of its real embedding: every singular value then appears once instead of twice); the 100
 * VampGlmEst iterations run in the library. */
#include "gateway_common.h"
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    const char* fn = "vamp";
    gw_nargs(fn, nrhs, 4, nlhs, 1);
    int m = (int)mxGetM(prhs[1]), n = (int)mxGetN(prhs[1]);
    if ((int)mxGetNumberOfElements(prhs[0]) != m) mexErrMsgIdAndTxt("jstsp:size", "%s: length(y) must equal size(A,1)", fn);
    double sigma = gw_scalar(prhs[2], fn, "sigma"), L = gw_scalar(prhs[3], fn, "L");
    mxArray* out[3] = {NULL, NULL, NULL};
    mxArray* in[1] = {(mxArray*)prhs[1]};
    if (mexCallMATLAB(3, out, 1, in, "svd") != 0) mexErrMsgIdAndTxt("jstsp:svd", "%s: svd(A) failed", fn);   /* [U,S,V] = svd(A) */
    void *t0, *t1, *tu;
    const mxComplexDouble* y = gw_complex(prhs[0], fn, "y", &t0);
    const mxComplexDouble* A = gw_complex(prhs[1], fn, "A", &t1);
    /* m <= n: U (m x m) of A A' (VampGlmEst.m:402-406); m > n: V (n x n) of A'A, which VampGlmEst.m:72-86 derives itself (:407-411) */
    const mxComplexDouble* U = gw_complex(out[m <= n ? 0 : 2], fn, "U", &tu);
    double* d = (double*)mxCalloc(m, sizeof(double));
    {   /* d = diag(S).^2, zero padded to m (vamp.m:33-34); the first min(m,n) entries are the eigenvalues either way */
        void* ts; const mxComplexDouble* S = gw_complex(out[1], fn, "S", &ts);
        int k = m < n ? m : n;
        for (int i = 0; i < k; ++i) d[i] = S[i + (size_t)m * i].real * S[i + (size_t)m * i].real;
        if (ts) mxFree(ts);
    }
    plhs[0] = mxCreateDoubleMatrix(n, 1, mxCOMPLEX);
    int rc = jstsp_vamp(gw_handle(fn), JSTSP_F64, JSTSP_HOST, m, n, 1, 100, 0.85, y, m, A, 0, &sigma, &L, U, 0, d, 0, mxGetComplexDoubles(plhs[0]), n);
    mxFree(d); if (t0) mxFree(t0); if (t1) mxFree(t1); if (tu) mxFree(tu);
    for (int k = 0; k < 3; ++k) if (out[k]) mxDestroyArray(out[k]);
    gw_status(rc, fn);
}
