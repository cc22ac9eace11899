/*
 * jstsp_b200.h - C ABI of libjstsp_b200.so, the B200-native engine for the
 * per-trial channel-estimation hot path of vlaxose/jstsp19.
 *
 * Every entry point replaces one MATLAB function of the reference (cited as
 * file:line relative to the reference root) and keeps its argument meaning.
 * A MATLAB MEX gateway / ctypes stub binds exactly these symbols (see
 * INTEGRATION.md).  No torch / C++ types cross this boundary.
 *
 * Conventions (identical for all entry points)
 *   - Matrices are COLUMN-MAJOR, complex values are INTERLEAVED (re,im), like
 *     MATLAB -R2018a arrays.  `dtype` selects the element type of every
 *     floating-point buffer of the call AND the arithmetic precision:
 *       JSTSP_F32: float  / interleaved complex float   (throughput path)
 *       JSTSP_F64: double / interleaved complex double  (reference precision)
 *   - `mem` says where ALL data buffers of the call live (JSTSP_HOST: pageable or
 *     pinned host memory, copied in/out by the library; JSTSP_DEVICE: device
 *     memory of the handle's GPU, used in place).
 *   - `batch` independent problems ("trials") are solved per call.  `ld_*`
 *     arguments are the distance, in ELEMENTS of that buffer's element type,
 *     between consecutive trials; 0 means "the same buffer for every trial".
 *   - Return value: 0 = ok; <0 = error (JSTSP_E_*), text via jstsp_last_error;
 *     >0 = number of trials whose result contains a non-finite value (the
 *     reference propagates NaN silently; the count lets the caller warn).
 *   - All work is enqueued on the handle's stream; calls with JSTSP_HOST buffers
 *     return after the results are in the caller's memory, calls with
 *     JSTSP_DEVICE buffers are asynchronous unless stated otherwise.
 *   - There is no CPU fallback: every entry point fails with JSTSP_E_CUDA if
 *     no sm_100-class device is usable.
 */
#ifndef JSTSP_B200_H
#define JSTSP_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jstsp_handle jstsp_handle;

enum { JSTSP_F32 = 0, JSTSP_F64 = 1 };
enum { JSTSP_HOST = 0, JSTSP_DEVICE = 1 };
/* `type` argument of proposed_algorithm.m:23-30: the string 'approximate' selects one
 * exact-line-search gradient step per iteration, anything else the exact LS solve. */
enum { JSTSP_APPROXIMATE = 0, JSTSP_STD = 1 };

enum {
    JSTSP_OK = 0,
    JSTSP_E_ARG = -1,         /* bad argument / NULL pointer / non-positive size */
    JSTSP_E_CUDA = -2,        /* CUDA runtime error (including "no device")      */
    JSTSP_E_UNSUPPORTED = -3, /* shape outside what the kernels cover            */
    JSTSP_E_NOMEM = -4        /* workspace allocation failed                     */
};

/* ---- handle --------------------------------------------------------------------- */
int jstsp_create(jstsp_handle** out, int device);
void jstsp_destroy(jstsp_handle* h);
const char* jstsp_last_error(const jstsp_handle* h);
const char* jstsp_version(void);
/* Use an existing cudaStream_t (passed as void*) instead of the handle's own stream. */
int jstsp_set_stream(jstsp_handle* h, void* cuda_stream);
/* Block until everything enqueued on the handle's stream has finished. */
int jstsp_synchronize(jstsp_handle* h);
/* Number of kernel launches issued through this handle since creation. */
long long jstsp_launch_count(const jstsp_handle* h);
/* Upper bound on the trials processed per internal pass (0 = automatic). */
int jstsp_set_chunk(jstsp_handle* h, int max_trials_per_pass);

/* Per-kernel-class device timing with CUDA events on the handle's stream (for roofline
 * reports).  enable: 0 = off, 1 = on, 2 = on and reset the accumulators.
 * jstsp_profile_read returns 0 and fills total_ms / launches / name for `slot`, or 1 when
 * `slot` is past the last class; it synchronises the stream. */
int jstsp_profile(jstsp_handle* h, int enable);
int jstsp_profile_read(jstsp_handle* h, int slot, double* total_ms, long long* launches, const char** name);

/* Developer hook: device buffer (8 x int64 per CTA of the largest grid) that receives in-kernel
 * clock64() phase timestamps of the ADMM fast-path kernels; NULL (default) disables. */
int jstsp_debug_buffer(jstsp_handle* h, void* device_buffer);

/* ---- proposed ADMM matrix completion ------------------------------------------- */
typedef struct {
    int N, M;          /* subY is N x M  (rows = RF-chain domain, cols = training instants) */
    int G, P;          /* A is N x G, B is P x M, S is G x P                               */
    int imax;          /* iterations (no early exit, proposed_algorithm.m:32)              */
    int type;          /* JSTSP_APPROXIMATE or JSTSP_STD                                   */
    int batch;
    long long ld_subY, ld_omega, ld_A, ld_B;  /* trial strides of the inputs (0 = shared)  */
    long long ld_S, ld_Y, ld_conv;            /* trial strides of the outputs              */
    int n_indx;        /* _angles only: length of each trial's indx_S ranking              */
    long long ld_indx; /* _angles only: trial stride of indx_S (0 = shared)                */
} jstsp_admm_desc;

/* [S,Y,convergence_error] = proposed_algorithm(subY,Omega,A,B,Imax,tau_Y,tau_S,rho,type)
 *   replaces basic_system_functions/proposed_algorithm.m:1-73.
 *   subY  N x M complex; omega N x M REAL (0/1 values, any real weights accepted);
 *   A N x G complex; B P x M complex; tau_Y,tau_S,rho: one double per trial
 *   (always double, in `mem` space); S G x P complex (out); Y N x M complex (out,
 *   may be NULL); conv imax x 3 real (out, may be NULL - the spectral-norm
 *   diagnostics of proposed_algorithm.m:51,67-69 are only computed when requested). */
int jstsp_proposed_algorithm(jstsp_handle* h, const jstsp_admm_desc* d, int dtype, int mem,
                             const void* subY, const void* omega, const void* A, const void* B,
                             const double* tau_Y, const double* tau_S, const double* rho,
                             void* S, void* Y, void* conv);

/* [S,Y,convergence_error] = proposed_algorithm_angles(subY,Omega,indx_S,A,B,Imax,tau_Y,tau_S,rho,type,greedy_nnz)
 *   replaces basic_system_functions/proposed_algorithm_angles.m:1-85.
 *   indx_S: d->n_indx 1-based column-major linear indices into the G x P grid, as int32
 *   (the gateway converts MATLAB doubles); entry k joins the support at iteration
 *   ceil((k-10)/5) (proposed_algorithm_angles.m:36).  greedy_nnz is unused by the
 *   reference and therefore absent here. */
int jstsp_proposed_algorithm_angles(jstsp_handle* h, const jstsp_admm_desc* d, int dtype, int mem,
                                    const void* subY, const void* omega, const int* indx_S,
                                    const void* A, const void* B,
                                    const double* tau_Y, const double* tau_S, const double* rho,
                                    void* S, void* Y, void* conv);

/* ---- singular-value thresholding and the SVT-based benchmark solvers -------------- */
/* X = svt(Y, tau)   replaces benchmark_algorithms/svt.m:1-15 (returns zeros when a
 * singular value is exactly 0, svt.m:7-13).  Y, X: Mr x Mt complex; tau: one double per trial. */
int jstsp_svt(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch,
              const void* Y, long long ld_Y, const double* tau, void* X, long long ld_X);

/* X = mc_svt(OH, Omega, Imax, tau, rho)   replaces benchmark_algorithms/mc_svt.m:1-12. */
int jstsp_mc_svt(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch, int imax,
                 const void* OH, long long ld_OH, const void* omega, long long ld_omega,
                 const double* tau, const double* rho, void* X, long long ld_X);

/* [X, convergence_error] = mc_admm(Htrue, OH, Omega, Imax, tau, rho)
 *   replaces benchmark_algorithms/mc_admm.m:1-34.  Htrue and conv (imax x 1 real) may
 *   both be NULL; conv needs Htrue. */
int jstsp_mc_admm(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch, int imax,
                  const void* Htrue, long long ld_H, const void* OH, long long ld_OH,
                  const void* omega, long long ld_omega, const double* tau, const double* rho,
                  void* X, long long ld_X, void* conv, long long ld_conv);

#ifdef __cplusplus
}
#endif
#endif /* JSTSP_B200_H */
