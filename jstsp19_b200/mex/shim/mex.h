/* mex.h - minimal stand-in for MATLAB's MEX C API (interleaved-complex, -R2018a flavour).
 *
 * Neither MATLAB nor Octave exists in the build image, so the gateways in ../ are compiled against
 * this shim for unit tests (tests/test_mex_gateways.py drives mexFunction through ctypes).  It
 * declares exactly the subset of the documented API the gateways use, with the documented
 * semantics; building with a real MATLAB (`mex -R2018a`) uses MATLAB's own mex.h instead. */
#ifndef JSTSP_MEX_SHIM_H
#define JSTSP_MEX_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;
typedef size_t mwIndex;
typedef struct { double real, imag; } mxComplexDouble;
typedef double mxDouble;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxCHAR_CLASS = 4, mxCELL_CLASS = 1 } mxClassID;
typedef struct mxArray_tag mxArray;

mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity flag);
mxArray* mxCreateNumericArray(mwSize ndim, const mwSize* dims, mxClassID cls, mxComplexity flag);
mxArray* mxCreateDoubleScalar(double v);
mxArray* mxCreateString(const char* s);
mxArray* mxCreateCellMatrix(mwSize m, mwSize n);
void mxSetCell(mxArray* c, mwIndex i, mxArray* v);
mxArray* mxGetCell(const mxArray* c, mwIndex i);
void mxDestroyArray(mxArray* a);
mwSize mxGetM(const mxArray* a);
mwSize mxGetN(const mxArray* a);               /* product of dims 2..end, like MATLAB */
mwSize mxGetNumberOfDimensions(const mxArray* a);
const mwSize* mxGetDimensions(const mxArray* a);
size_t mxGetNumberOfElements(const mxArray* a);
int mxIsComplex(const mxArray* a);
int mxIsDouble(const mxArray* a);
int mxIsChar(const mxArray* a);
int mxIsCell(const mxArray* a);
mxDouble* mxGetDoubles(const mxArray* a);
mxComplexDouble* mxGetComplexDoubles(const mxArray* a);
double mxGetScalar(const mxArray* a);
int mxGetString(const mxArray* a, char* buf, mwSize buflen);
void* mxMalloc(size_t n);
void* mxCalloc(size_t n, size_t sz);
void mxFree(void* p);

void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...);
void mexWarnMsgIdAndTxt(const char* id, const char* fmt, ...);
int mexAtExit(void (*fn)(void));
int mexCallMATLAB(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name);

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);

/* ---- shim-only test hooks (not part of MATLAB's API) ---- */
typedef int (*jstsp_shim_callback)(int nlhs, mxArray* plhs[], int nrhs, mxArray* prhs[], const char* name);
void jstsp_shim_set_callback(jstsp_shim_callback cb);       /* serves mexCallMATLAB (randn, rand, randperm, svd) */
const char* jstsp_shim_last_error_id(void);
const char* jstsp_shim_last_error_msg(void);
const char* jstsp_shim_last_warning_id(void);
void jstsp_shim_clear(void);
int jstsp_shim_call(void (*fn)(int, mxArray**, int, const mxArray**), int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);  /* returns 1 if mexErrMsgIdAndTxt fired */
void jstsp_shim_run_atexit(void);

#ifdef __cplusplus
}
#endif
#endif
