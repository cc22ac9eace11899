/* mex_shim.c - fake MEX runtime behind shim/mex.h (unit-test infrastructure for the gateways). */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct mxArray_tag {
    mxClassID cls;
    int is_complex;
    mwSize ndim;
    mwSize dims[4];
    void* data;          /* doubles / mxComplexDouble / char / mxArray* cells */
};

static size_t numel(const mxArray* a) { size_t n = 1; for (mwSize i = 0; i < a->ndim; ++i) n *= a->dims[i]; return n; }

mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity flag) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = cls; a->is_complex = flag == mxCOMPLEX; a->ndim = ndim < 2 ? 2 : ndim;
    a->dims[0] = a->dims[1] = 1; a->dims[2] = a->dims[3] = 1;
    for (mwSize i = 0; i < ndim && i < 4; ++i) a->dims[i] = dims[i];
    size_t n = numel(a);
    a->data = calloc(n ? n : 1, a->is_complex ? sizeof(mxComplexDouble) : sizeof(double));
    return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag) { mwSize d[2] = {m, n}; return mxCreateNumericArray(2, d, mxDOUBLE_CLASS, flag); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); ((double*)a->data)[0] = v; return a; }
mxArray* mxCreateString(const char* s) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = mxCHAR_CLASS; a->ndim = 2; a->dims[0] = 1; a->dims[1] = strlen(s); a->dims[2] = a->dims[3] = 1;
    a->data = strdup(s);
    return a;
}
mxArray* mxCreateCellMatrix(mwSize m, mwSize n) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = mxCELL_CLASS; a->ndim = 2; a->dims[0] = m; a->dims[1] = n; a->dims[2] = a->dims[3] = 1;
    a->data = calloc((m * n) != 0 ? m * n : 1, sizeof(mxArray*));
    return a;
}
void mxSetCell(mxArray* c, mwIndex i, mxArray* v) { ((mxArray**)c->data)[i] = v; }
mxArray* mxGetCell(const mxArray* c, mwIndex i) { return ((mxArray**)c->data)[i]; }
void mxDestroyArray(mxArray* a) {
    if (!a) return;
    if (a->cls == mxCELL_CLASS) { size_t n = numel(a); for (size_t i = 0; i < n; ++i) mxDestroyArray(((mxArray**)a->data)[i]); }
    free(a->data); free(a);
}
mwSize mxGetM(const mxArray* a) { return a->dims[0]; }
mwSize mxGetN(const mxArray* a) { size_t n = 1; for (mwSize i = 1; i < a->ndim; ++i) n *= a->dims[i]; return n; }
mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->ndim; }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims; }
size_t mxGetNumberOfElements(const mxArray* a) { return numel(a); }
int mxIsComplex(const mxArray* a) { return a->is_complex; }
int mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
int mxIsCell(const mxArray* a) { return a->cls == mxCELL_CLASS; }
mxDouble* mxGetDoubles(const mxArray* a) { return (a->cls == mxDOUBLE_CLASS && !a->is_complex) ? (mxDouble*)a->data : NULL; }
mxComplexDouble* mxGetComplexDoubles(const mxArray* a) { return (a->cls == mxDOUBLE_CLASS && a->is_complex) ? (mxComplexDouble*)a->data : NULL; }
double mxGetScalar(const mxArray* a) { return a->is_complex ? ((mxComplexDouble*)a->data)[0].real : ((double*)a->data)[0]; }
int mxGetString(const mxArray* a, char* buf, mwSize buflen) {
    if (a->cls != mxCHAR_CLASS) return 1;
    size_t n = a->dims[1];
    if (n + 1 > buflen) { memcpy(buf, a->data, buflen - 1); buf[buflen - 1] = 0; return 1; }
    memcpy(buf, a->data, n); buf[n] = 0;
    return 0;
}
void* mxMalloc(size_t n) { return malloc(n ? n : 1); }
void* mxCalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz ? sz : 1); }
void mxFree(void* p) { free(p); }

static char g_err_id[256], g_err_msg[1024], g_warn_id[256];
static jmp_buf g_jmp;
static int g_jmp_armed = 0;
static jstsp_shim_callback g_cb = NULL;
static void (*g_atexit)(void) = NULL;

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    snprintf(g_err_id, sizeof g_err_id, "%s", id);
    vsnprintf(g_err_msg, sizeof g_err_msg, fmt, ap);
    va_end(ap);
    if (g_jmp_armed) longjmp(g_jmp, 1);      /* MATLAB unwinds out of mexFunction; so does the shim */
    fprintf(stderr, "mexErrMsgIdAndTxt outside jstsp_shim_call: %s: %s\n", g_err_id, g_err_msg);
    abort();
}
void mexWarnMsgIdAndTxt(const char* id, const char* fmt, ...) { (void)fmt; snprintf(g_warn_id, sizeof g_warn_id, "%s", id); }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
int mexCallMATLAB(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name) {
    if (!g_cb) return 1;
    return g_cb(nlhs, plhs, nrhs, prhs, name);
}
void jstsp_shim_set_callback(jstsp_shim_callback cb) { g_cb = cb; }
const char* jstsp_shim_last_error_id(void) { return g_err_id; }
const char* jstsp_shim_last_error_msg(void) { return g_err_msg; }
const char* jstsp_shim_last_warning_id(void) { return g_warn_id; }
void jstsp_shim_clear(void) { g_err_id[0] = g_err_msg[0] = g_warn_id[0] = 0; }
int jstsp_shim_call(void (*fn)(int, mxArray**, int, const mxArray**), int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    jstsp_shim_clear();
    g_jmp_armed = 1;
    if (setjmp(g_jmp)) { g_jmp_armed = 0; return 1; }
    fn(nlhs, plhs, nrhs, prhs);
    g_jmp_armed = 0;
    return 0;
}
void jstsp_shim_run_atexit(void) { if (g_atexit) { g_atexit(); g_atexit = NULL; } }
