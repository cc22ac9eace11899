/* gateway_common.h - shared plumbing of the MEX gateways (one gateway source per MATLAB function
 * name of the reference; build with `mex -R2018a <name>.c -ljstsp_b200`).
 * All gateways run the fp64 path of the library so results match the reference's fp64 arithmetic;
 * outputs are allocated by MATLAB (mxCreate*), the library never returns memory it owns. */
#ifndef JSTSP_GATEWAY_COMMON_H
#define JSTSP_GATEWAY_COMMON_H
#include <string.h>
#include "mex.h"
#include "../../include/jstsp_b200.h"

static jstsp_handle* g_handle = NULL;
static void jstsp_gateway_cleanup(void) { if (g_handle) { jstsp_destroy(g_handle); g_handle = NULL; } }

static jstsp_handle* gw_handle(const char* fn) {
    if (!g_handle) {
        if (jstsp_create(&g_handle, 0) != JSTSP_OK) {
            g_handle = NULL;
            mexErrMsgIdAndTxt("jstsp:nogpu", "%s: no usable B200-class CUDA device (libjstsp_b200 has no CPU fallback)", fn);
        }
        mexAtExit(jstsp_gateway_cleanup);
    }
    return g_handle;
}

/* interleaved complex view of a double array; real inputs are widened (MATLAB passes real arrays
 * wherever the imaginary part is all zero).  *tmp receives a buffer to mxFree (or NULL). */
static const mxComplexDouble* gw_complex(const mxArray* a, const char* fn, const char* what, void** tmp) {
    *tmp = NULL;
    if (!mxIsDouble(a)) mexErrMsgIdAndTxt("jstsp:type", "%s: %s must be a double array", fn, what);
    if (mxIsComplex(a)) return mxGetComplexDoubles(a);
    size_t n = mxGetNumberOfElements(a);
    mxComplexDouble* c = (mxComplexDouble*)mxMalloc((n ? n : 1) * sizeof(mxComplexDouble));
    const double* r = mxGetDoubles(a);
    for (size_t i = 0; i < n; ++i) { c[i].real = r[i]; c[i].imag = 0.0; }
    *tmp = c;
    return c;
}
static const double* gw_real(const mxArray* a, const char* fn, const char* what) {
    if (!mxIsDouble(a) || mxIsComplex(a)) mexErrMsgIdAndTxt("jstsp:type", "%s: %s must be a real double array", fn, what);
    return mxGetDoubles(a);
}
static double gw_scalar(const mxArray* a, const char* fn, const char* what) {
    if (!mxIsDouble(a) || mxGetNumberOfElements(a) != 1) mexErrMsgIdAndTxt("jstsp:type", "%s: %s must be a scalar", fn, what);
    return mxGetScalar(a);
}
static void gw_nargs(const char* fn, int nrhs, int want_rhs, int nlhs, int max_lhs) {
    if (nrhs != want_rhs) mexErrMsgIdAndTxt("jstsp:nargin", "%s: expected %d input arguments, got %d", fn, want_rhs, nrhs);
    if (nlhs > max_lhs) mexErrMsgIdAndTxt("jstsp:nargout", "%s: at most %d output arguments", fn, max_lhs);
}
/* status mapping: <0 -> MATLAB error (after temporaries are released by the caller), >0 -> warning */
static void gw_status(int rc, const char* fn) {
    if (rc < 0) mexErrMsgIdAndTxt("jstsp:failed", "%s: %s (code %d)", fn, g_handle ? jstsp_last_error(g_handle) : "no handle", rc);
    if (rc > 0) mexWarnMsgIdAndTxt("jstsp:nonfinite", "%s: %d trial(s) produced non-finite values", fn, rc);
}
#endif
