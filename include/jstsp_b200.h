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
/* Use an existing cudaStream_t (passed as void*) instead of the handle's own stream.  The handle owns ONE workspace: when the stream
 * changes, the new stream is made to wait (event) for the work already queued on the old one, so calls issued under different streams
 * are serialised rather than left to overwrite each other's state. */
int jstsp_set_stream(jstsp_handle* h, void* cuda_stream);
/* Number of trials of the last proposed_algorithm* / solver call whose output held a non-finite value (the reference's silent NaN
 * propagation).  HOST-buffer calls return the same number themselves; DEVICE-buffer calls are asynchronous and
 * return 0, so this function synchronises the handle's stream and reads the device counter.  < 0 on error. */
long long jstsp_nonfinite_count(jstsp_handle* h);
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

/* The same estimator with the dictionary given by its factors, exactly as the reference's drivers hold them before
 * they form B (plot_errorVSsnr.m:132-136, plot_errorVSdelays.m:130-134):
 *     B((l-1)*Gt+1 : l*Gt, :) = Dt' * Psi_bar(:,:,l),   l = 1..L
 *   Dt       Nt x Gt complex (output 6 of wideband_mmwave_channel.m:1), Gt = d->P / L, trial stride ld_Dt (0 = shared);
 *   Psi_bar  Nt x M x L complex (output of proposed_hbf.m:1 / wideband_hybBF_comm_system_training.m:1), trial stride ld_Psi;
 *   indx_S   NULL for proposed_algorithm, the ranking of proposed_algorithm_angles otherwise (d->n_indx, d->ld_indx);
 *   d->ld_B is ignored; everything else as in jstsp_proposed_algorithm.
 * When Psi_bar has the structure the drivers give it - Psi_bar(k,j,l) = e_k(j-l) (rows of toeplitz(s_k), proposed_hbf.m:15-18)
 * with components that are exact in bf16 after one common scaling (4-QAM pilots, plot_errorVSsnr.m:63-67) - and the call is
 * fp32 'approximate' with N = 16, Nt = 64, M % 128 == 0 and conv == NULL, both dictionary products of every iteration run on
 * tcgen05 tensor cores out of a 34 KiB bf16 pilot tile per 128 columns (csrc/admm_psi.cuh) and the dense dictionary is never
 * streamed.  Any other input is served by materialising B on the device and running the dense kernels of
 * jstsp_proposed_algorithm; results agree to rounding either way.  jstsp_last_path reports which one ran. */
int jstsp_proposed_algorithm_psi(jstsp_handle* h, const jstsp_admm_desc* d, int dtype, int mem,
                                 const void* subY, const void* omega, const int* indx_S, const void* A,
                                 const void* Dt, long long ld_Dt, const void* Psi_bar, long long ld_Psi, int Nt, int L,
                                 const double* tau_Y, const double* tau_S, const double* rho,
                                 void* S, void* Y, void* conv);
/* Same estimator, fed the pilot SEQUENCES instead of Psi_bar: pilots is Nt x M per trial, row k = s_k, the vector the drivers
 * pass to toeplitz() at plot_errorVSsnr.m:63-67 (= Psi_i(1,:,k)).  Psi_bar(k,:,l) = row l of toeplitz(s_k) (proposed_hbf.m:15-18,
 * MATLAB's Hermitian rule below the diagonal) is expanded on the device, so a HOST call moves L times fewer dictionary bytes.
 * Results are identical to jstsp_proposed_algorithm_psi on the expanded Psi_bar (same kernels from there on). */
int jstsp_proposed_algorithm_pilots(jstsp_handle* h, const jstsp_admm_desc* d, int dtype, int mem,
                                    const void* subY, const void* omega, const int* indx_S, const void* A,
                                    const void* Dt, long long ld_Dt, const void* pilots, long long ld_pilots, int Nt, int L,
                                    const double* tau_Y, const double* tau_S, const double* rho,
                                    void* S, void* Y, void* conv);

/* Path taken by the last jstsp_proposed_algorithm_psi call on this handle: 1 = dense kernels on the materialised
 * dictionary, 2 = Psi-domain tcgen05 kernel (0 = no call yet). */
int jstsp_last_path(const jstsp_handle* h);
/* Form of the Psi-domain path taken by the last pass: 1 = the whole solve (all Imax iterations of proposed_algorithm.m:32-70) ran
 * in one persistent kernel per pass (csrc/admm_mega.cuh), 0 = four kernels per iteration (csrc/admm_psi.cuh). */
int jstsp_last_variant(const jstsp_handle* h);

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

/* ---- orthogonal matching pursuit ------------------------------------------------------ */
/* [x_hat, indexSet, v, targetMatrix] = OMP(A, v, m, snr)   replaces benchmark_algorithms/OMP.m:1-32
 *   (snr is unused by the reference and absent here; v is echoed by the gateway).
 *   A measures x size_d complex; v measures x 1; x_hat size_d x 1 (out); index_set m int32 per
 *   trial, 1-based, first-maximum tie rule of OMP.m:17 (out); target_matrix measures x m (out, may
 *   be NULL); ambiguous (may be NULL): per trial, the number of iterations whose best/runner-up
 *   correlation magnitudes differ by less than margin_tol (relative) - 0 means the support is
 *   robust against evaluation-order rounding. */
int jstsp_omp(jstsp_handle* h, int dtype, int mem, int measures, int size_d, int m, int batch,
              const void* A, long long ld_A, const void* v, long long ld_v,
              void* x_hat, long long ld_x, int* index_set, void* target_matrix, int* ambiguous, double margin_tol);

/* OMP over the Kronecker dictionary Phi = kron(B.', A) without materialising it:
 *   [x_hat, indexSet] = OMP(kron(B.', A), vec(Y), m)   (benchmark_algorithms/OMP.m:1-32 on the operands the
 *   drivers build at plot_errorVSdelays.m:77-78 / plot_errorVSframelength.m:78-79; BASELINE config 2 would
 *   need an 8192 x 262144 Phi).  A N x G, B P x M, Y N x M complex (ld_A / ld_B = 0: shared by all trials).
 *   index_set: m int32 per trial, 1-based linear index g + G*(p-1) into the G x P unknown, first-maximum
 *   rule of OMP.m:17.  x_hat (G*P per trial, may be NULL - it is 2 MiB per trial at config 2), x_sel (m per
 *   trial, may be NULL): coefficient of pick t as OMP.m:29-31 scatters it, residual (N x M, may be NULL):
 *   v - T x of the last iteration (OMP.m:20-21), ambiguous / margin_tol as in jstsp_omp. */
int jstsp_omp_kron(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int m, int batch,
                   const void* A, long long ld_A, const void* B, long long ld_B, const void* Y, long long ld_Y,
                   void* x_hat, long long ld_x, int* index_set, void* x_sel, void* residual, long long ld_r,
                   int* ambiguous, double margin_tol);

/* Joint (MMV) OMP - replaces the reference's external sparse-plex call
 *   spx.pursuit.joint.OrthogonalMatchingPursuit(A, K).solve(Y).Z
 *   (plot_errorVSsnr.m:116-118, plot_errorVSdelays.m:114-115, plot_time_comparisions.m:101-102).
 *   sparse-plex is neither vendored nor version-pinned by the reference (README.md:9): this is the textbook
 *   row-l2 SOMP, a documented replacement rather than a parity target.  A N x D, Y N x S, Z D x S (out);
 *   support K int32 per trial, 1-based in selection order, 0 = unused (out); n_iters per trial (out, may be
 *   NULL); residual N x S (out, may be NULL).  Stops after K picks, at min(N, D) atoms, on a repeated pick,
 *   or when ||R||_F <= res_tol * ||Y||_F. */
int jstsp_somp(jstsp_handle* h, int dtype, int mem, int N, int D, int S, int K, int batch,
               const void* A, long long ld_A, const void* Y, long long ld_Y,
               void* Z, long long ld_Z, int* support, int* n_iters, void* residual, long long ld_R, double res_tol);

/* [S, convergence_error] = sparse_admm(Htrue, OH, Dr, Dt, Imax)
 *   replaces benchmark_algorithms/sparse_admm.m:1-36 (rho = 0.01, tau_s = 1e-4 hard-coded, :12-13).
 *   Like the reference this needs Gr == Mr and Gt == Mt (Dr Mr x Mr, Dt Mt x Mt); Mr, Mt <= 64.
 *   Htrue and conv (imax x 1 real) may both be NULL. */
int jstsp_sparse_admm(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch, int imax,
                      const void* Htrue, long long ld_H, const void* OH, long long ld_OH,
                      const void* Dr, long long ld_Dr, const void* Dt, long long ld_Dt,
                      void* S, long long ld_S, void* conv, long long ld_conv);

/* ---- VAMP ------------------------------------------------------------------------------- */
/* x = vamp(y, A, sigma, L)   replaces benchmark_algorithms/vamp.m:1-55 + VampGlmEst.m:350-521 for
 *   nit iterations (vamp.m: 100) with damping `damp` (vamp.m: 0.85).  y m x 1, A m x n complex;
 *   sigma (noise variance) and L (expected non-zeros) one double per trial.  The spectral basis is an
 *   input, computed by the caller exactly where vamp.m:32 calls svd: for m <= n `basis` is the m x m
 *   complex eigenvector matrix U of A A^H and d its m eigenvalues (= squared singular values of A,
 *   each of which appears twice in the reference's real-embedded d).  For m > n (VampGlmEst.m:407-411,
 *   which works in the eigenbasis of A'A, :72-86) `basis` is the n x n eigenvector matrix V of A^H A
 *   (the V of the same svd) and d its n eigenvalues. */
int jstsp_vamp(jstsp_handle* h, int dtype, int mem, int m, int n, int batch, int nit, double damp,
               const void* y, long long ld_y, const void* A, long long ld_A, const double* sigma, const double* L,
               const void* basis, long long ld_basis, const void* d, long long ld_d, void* x, long long ld_x);

/* ---- measurement model ------------------------------------------------------------------ */
/* [H,Zbar,Ar,At,Dr,Dt] = wideband_mmwave_channel(L,Mr,Mt,Ncl,Nray,Gr,Gt)
 *   replaces basic_system_functions/wideband_mmwave_channel.m:1-62, quirks included (page-1
 *   steering vectors :24-25, cumulative cluster sum :29).  The draws are inputs, per trial and in
 *   the reference's order: normals[2*(l*Np+ray)+{0,1}] = the two randn of :19,
 *   uniforms[2*(l*Np+ray)+{0,1}] = the rand of :20 and of :22.  Any output may be NULL.
 *   H Mr x Mt x L, Zbar Gr x (L*Gt), Ar Mr x Np x L, At Mt x Np x L, Dr Mr x Gr, Dt Mt x Gt. */
int jstsp_wideband_mmwave_channel(jstsp_handle* h, int dtype, int mem, int L, int Mr, int Mt, int ncl, int nray, int Gr, int Gt, int batch,
                                  const double* normals, const double* uniforms,
                                  void* H, void* Zbar, void* Ar, void* At, void* Dr, void* Dt);

typedef struct {
    int Nr, Nt, L, T;      /* H is Nr x Nt x L, T training columns                                 */
    int Wc;                /* W_e = W(:, 1:Wc)  (Lr_e of proposed_hbf.m:11, Lr of hbf.m:23)          */
    int Lr;                /* ones per mask column (proposed_hbf.m:39); ignored when perm == NULL   */
    int psi_mode;          /* 0: Psi is Psi_i (Tp x Tp x Nt, only rows 1..L are read);              */
    int Tp;                /* 1: Psi is the pilot matrix (Nt x T, row k = s_k), Toeplitz rows built on the fly */
    int batch;
    long long ld_H, ld_N, ld_Psi, ld_W;   /* trial strides (0 = shared)                            */
} jstsp_meas_desc;

/* Measurement synthesis shared by
 *   [Y_proposed_hbf,W_e,Psi_bar,Omega,Y] = proposed_hbf(H,N,Psi_i,T,Lr_e,Lr,W)   (proposed_hbf.m:1-44)
 *   [Y_conventional_hbf,W_c,Psi_bar,Y]   = hbf(H,N,Psi_i,T,Lr,W)                 (hbf.m:1-26)
 *   and the synthesis part of wideband_hybBF_comm_system_training.m:24-56:
 *     Y = sum_l H(:,:,l) Psi_bar(:,:,l), Psi_bar(k,:,l) = Psi_i(l,:,k);  Y_out = [Omega .*] (W_e' (Y + N)).
 *   perm: the randperm draws of proposed_hbf.m:38, T x Wc int32 per trial (1-based), or NULL for
 *   the unmasked (hbf) output.  Outputs (any may be NULL): Y_out Wc x T, W_e Nr x Wc,
 *   Psi_bar Nt x T x L, Omega Wc x T real, Y_noiseless Nr x T. */
int jstsp_measure(jstsp_handle* h, const jstsp_meas_desc* d, int dtype, int mem,
                  const void* H, const void* N, const void* Psi, const void* W, const int* perm,
                  void* Y_out, void* W_e, void* Psi_bar, void* Omega, void* Y_noiseless);

/* ---- combiner codebooks and pilot symbols -------------------------------------- */
enum { JSTSP_BF_FFT = 0, JSTSP_BF_RAND = 1, JSTSP_BF_RAND_PS = 2, JSTSP_BF_PS = 3, JSTSP_BF_ZC = 4, JSTSP_BF_QUANTIZED_4 = 5, JSTSP_BF_QUANTIZED = 6 };
/* B = createBeamformer(N, beamformer_type)   replaces basic_system_functions/createBeamformer.m:1-35; B is N x N (out).
 *   draws (random codebooks only, in `mem` space): 'rand' N*N int32 in 0..3 = index of each randsrc(N,N,[1 -1 1j -1j]) entry
 *   into that alphabet, column-major (:8); 'rand_ps' N int32 in 1..32 = randi(32,1,N) (:10-11).  NULL otherwise. */
int jstsp_create_beamformer(jstsp_handle* h, int dtype, int mem, int N, int type, const int* draws, void* B);
/* symbols = qam4mod(input, mode, N)   replaces basic_system_functions/qam4mod.m:1-33.
 *   mode 0 ('mod'): draws = n int32 in 0..3, the index of each randsrc draw into the alphabet of qam4mod.m:7; input unused.
 *   mode 1 ('demod'): hard decision of the n soft symbols in `input` (qam4mod.m:12-31); draws unused. */
int jstsp_qam4mod(jstsp_handle* h, int dtype, int mem, int mode, long long n, const int* draws, const void* input, void* symbols);

/* ---- driver-side metric and parameters ("next" rows) ---------------------- */
/* nmse[b] = min(1, norm(S-Zbar)^2/norm(Zbar)^2) with matrix 2-norms (plot_errorVSsnr.m:138-141). */
int jstsp_nmse(jstsp_handle* h, int dtype, int mem, int G, int P, int batch,
               const void* S, long long ld_S, const void* Zbar, long long ld_Z, double* nmse);
/* tau_Y = 1/norm(Y,'fro')^2, tau_Z = 1/norm(Zbar,'fro')^2/2, rho = sqrt(lambda_k(Y'Y)/norm(Y,'fro')^2)
 * with k = kth_eig: 6 for min(eigs(.)) (plot_errorVSsnr.m:127-130), 1 for max(eigs(.))
 * (plot_errorVSdelays.m:127-128).  Zbar / tau_Z may be NULL. */
int jstsp_admm_parameters(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int batch, int kth_eig,
                          const void* Y, long long ld_Y, const void* Zbar, long long ld_Z,
                          double* tau_Y, double* tau_Z, double* rho);

/* rate[b] = real(log2(det(eye(n) + scale[b] * X_b * X_b')))  for X n x m complex, n <= 64: the achievable-rate / capacity
 * metric of the sweep drivers:
 *   plot_rateVSframelength.m:113,130,135   X = Zbar,     scale = 1 / (Nr * (sigma^2 + NMSE))
 *   plot_capacity.m:47-66, plot_ee.m:36-87 X = W_c' * Y, scale = 1 / (sigma^2 * Nt)
 * scale and rate are doubles in `mem` space. */
int jstsp_log2det_rate(jstsp_handle* h, int dtype, int mem, int n, int m, int batch,
                       const void* X, long long ld_X, const double* scale, double* rate);

/* Spectral efficiency of one receiver design of the capacity / energy-efficiency sweeps:
 *   rate[b] = real(log2(det(eye(Mr) + scale[b] * Wsel' * (Y*Y') * Wsel))),  Wsel = W(:, cols(1:Mr))
 * replaces plot_capacity.m:47,52,57,64 and plot_ee.m:47,52,57,64 (scale = 1/square_noise_variance * 1/Nt).  Y is the noiseless received
 * block Nr x T (4th output of hbf.m:1 / 5th of proposed_hbf.m:1), W the Nr x Wc combiner, cols the 1-based column choice of the design
 * (ind(1:Mr) of ind = randperm(Mr_e), plot_capacity.m:63), Mr entries per trial (ld_cols = 0: shared); NULL = the first Mr columns
 * (W_c = W(:, 1:Lr), hbf.m:24).  Mr <= 64. */
int jstsp_capacity(jstsp_handle* h, int dtype, int mem, int Nr, int T, int Wc, int Mr, int batch,
                   const void* Y, long long ld_Y, const void* W, long long ld_W, const int* cols, long long ld_cols,
                   const double* scale, double* rate);

/* The loop body of plot_capacity.m:35-66 / plot_ee.m:36-66 for a whole Mr range in one call: the four receiver designs on the same noiseless blocks Y.
 *   design 0 digital beamforming (all Nr columns of W_zc, plot_capacity.m:45-47), 1 conventional HBF with phase shifters (W_q(:, 1:Mr), :50-52),
 *   2 conventional HBF with ZC (W_zc(:, 1:Mr), :55-57), 3 proposed (W_q(:, ind(1:Mr)) with ind = randperm(Mr_e), :61-64).
 * mr_range: n_mr host integers (Mr_range, plot_capacity.m:16); W_zc / W_q: the Nr x Nr codebooks createBeamformer(Nr,'ZC') / (Nr,'quantized'), shared by the batch;
 * ind: 1-based permutations, at least max(Mr) entries per trial, trial stride ld_ind (0: shared); scale[b] = 1/(square_noise_variance * Nt).
 * out[(i_mr * 4 + design) * batch + b] (doubles in `mem` space); mean over b gives mean_Capacity_*(mr_index), and rate ./ jstsp_power_model the energy efficiency
 * (plot_ee.m:84-87). */
int jstsp_capacity_sweep(jstsp_handle* h, int dtype, int mem, int Nr, int T, int batch, int n_mr, const int* mr_range,
                         const void* Y, long long ld_Y, const void* W_zc, const void* W_q, const int* ind, long long ld_ind,
                         const double* scale, double* out);

/* Power consumption of the four receiver designs of plot_ee.m:69-77 (Pcirc = 0, Psw = 5 mW, Pps = 15 mW, Plna = 20 mW, Pps_zc = 60 mW):
 * power4 = { digital beamforming, conventional HBF with phase shifters, conventional HBF with ZC, proposed }.  Energy efficiency is
 * mean(rate) / power (plot_ee.m:84-87).  Host arithmetic, no handle. */
int jstsp_power_model(int Nr, int Mr, int Mr_e, double* power4);

/* Random draws of one batch of Monte-Carlo trials, generated on the device (csrc/rng.cu).  Replaces the calls to MATLAB's global stream in
 * the trial loop: randn / rand of wideband_mmwave_channel.m:19-22, randn of plot_errorVSsnr.m:60, randsrc of qam4mod.m:7-8 (via
 * plot_errorVSsnr.m:63-67) and randperm of proposed_hbf.m:37.  The reference fixes no seed, so the distributions are its behaviour, not the
 * numbers.  Every value is Philox4x32-10 of counter = (block index, stream id, global trial index) under key = seed: trial t gets the same
 * numbers whichever batch, shard or GPU count computes it.  DEVICE buffers only; any output may be NULL.
 *   sigma2   [batch] noise variances 10^(-snr/10)                                   (in)
 *   normals  [batch][L][Np][2] double, N(0,1): (randn, randn) of each path gain      -> jstsp_wideband_mmwave_channel
 *   uniforms [batch][L][Np][2] double, U(0,1): (rand, rand) of each path's angles    -> jstsp_wideband_mmwave_channel
 *   pilots   [batch][T][Nt] complex of `dtype`: s_k(t) at [t][k], (+-1 +-j)/sqrt 2    -> jstsp_measure (psi_mode 1), jstsp_proposed_algorithm_pilots
 *   noise    [batch][T][Nr] complex of `dtype`: sqrt(sigma2/2) (randn + j randn)      -> jstsp_measure
 *   perm     [batch][T][Nr] int32, 1-based: randperm(Nr) of each training instant    -> jstsp_measure */
int jstsp_draw_trials(jstsp_handle* h, int dtype, unsigned long long seed, long long first_trial, int batch, int Nr, int Nt, int L, int Np, int T,
                      const double* sigma2, double* normals, double* uniforms, void* pilots, void* noise, int* perm);
/* The generator itself on the host (known-answer tests): out[4] = Philox4x32-10(ctr[4], key[2]).  No handle, no device. */
void jstsp_philox4x32_10(const unsigned* ctr, const unsigned* key, unsigned* out);

/* Least-squares baseline of the drivers, batched:  S = pinv(A) * Y * pinv(B)   (plot_errorVSsnr.m:83, plot_errorVSsnr_approx.m:61,67)
 * and / or  YpinvB = Y * pinv(B)  (the right-hand sides handed to the joint OMP, plot_errorVSsnr.m:117).
 * A N x G, B P x M, Y N x M, S G x P, YpinvB N x P; either output may be NULL (A may be NULL when S is).  Full-rank operands: pinv is
 * formed through the Gram matrix of the short side (no singular-value truncation).  Returns the number of trials with a non-finite S. */
int jstsp_ls_estimate(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int batch,
                      const void* A, long long ld_A, const void* B, long long ld_B, const void* Y, long long ld_Y,
                      void* S, long long ld_S, void* YpinvB, long long ld_YpinvB);

#ifdef __cplusplus
}
#endif
#endif /* JSTSP_B200_H */
