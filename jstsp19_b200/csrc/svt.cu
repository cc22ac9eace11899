// svt.cu - singular-value thresholding and the two SVT-based benchmark solvers.
//
//   jstsp_svt      replaces benchmark_algorithms/svt.m:1-15
//   jstsp_mc_svt   replaces benchmark_algorithms/mc_svt.m:1-12
//   jstsp_mc_admm  replaces benchmark_algorithms/mc_admm.m:1-34  (its dense (MrMt)^2 `A\`
//                  solve, mc_admm.m:11-17,24, is the element-wise divide by Omega + rho)
//
// These solvers are HBM-bound: per iteration one Jacobi kernel on the Mr x Mr Gram matrix
// and ONE fused streaming kernel that applies the spectral weights (X = W Z), performs the
// element-wise primal/dual updates and accumulates the Gram matrix of the next SVT input,
// so every state matrix is read once and written once per iteration.
#include "common.cuh"
#include "gemm_cores.cuh"
#include "jacobi.cuh"

namespace jstsp {

enum { MODE_SVT = 0, MODE_MCSVT = 1, MODE_MCADMM = 2 };

template <typename T>
struct SvtP {
    int N, M, RP, MC, nmc, iter, imax;
    const cx<T>* in; long long ld_in;        // svt: Y ; mc_*: OH
    const T* omega;  long long ld_omega;
    const cx<T>* Htrue; long long ld_H;
    const double *tau, *rho;                 // per trial; threshold = tau / rho (rho == nullptr -> 1)
    cx<T> *Ys, *Zs;                          // state (N x M per trial)
    cx<T>* W; double* gram;                  // weights, partial Grams [b][nmc][2NN]
    double* Uprev;                           // eigenvectors of the previous iteration [b][2NN] (warm start of the Jacobi solve)
    double* cgram;                           // conv: [b][2][nmc][2NN]  (X - Htrue, Htrue)
    double* convd;                           // conv: [b][imax] ; slot [b][imax] holds sigma_max(Htrue)^2
    cx<T>* out; long long ld_out;
};

constexpr int kEigThreads = 256;    // one 2 x 2 block of a 32 x 32 problem per thread and rotation step

template <typename T>
__device__ __forceinline__ void weights_body(const SvtP<T>& p, int b, unsigned char* smem) {
    JacobiSmem sm; sm.carve(smem, p.N);
    const int n = p.N, nn = n * n;
    const double* g = p.gram + (size_t)b * p.nmc * 2 * nn;
    for (int t = threadIdx.x; t < nn; t += blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int c = 0; c < p.nmc; ++c) { re += g[(size_t)c * 2 * nn + 2 * t]; im += g[(size_t)c * 2 * nn + 2 * t + 1]; }
        sm.Are[t] = re; sm.Aim[t] = im;
    }
    __syncthreads();
    // Warm start from the previous iteration's eigenvectors (the mc_svt / mc_admm iterates move slowly, so Q^H G Q is nearly
    // diagonal and one or two sweeps suffice); every 16th iteration restarts cold to shed drift.  Same scheme as k_svt_weights.
    double* Up = p.Uprev ? p.Uprev + (size_t)b * 2 * nn : nullptr;
    const bool warm = Up && p.iter > 0 && (p.iter % 16) != 0;
    if (warm) {
        for (int t = threadIdx.x; t < nn; t += blockDim.x) { sm.Ure[t] = Up[t]; sm.Uim[t] = Up[nn + t]; }
        __syncthreads();
        jacobi_similarity_block(sm, n, reinterpret_cast<double*>(smem + JacobiSmem::bytes(n)));
    }
    // fp32 solves: W is rounded to fp32 (6e-8), so a sweep that starts at a relative off-diagonal norm of 1e-5 (and ends near 1e-10) is the last
    jacobi_hermitian_block(sm, n, 24, warm, sizeof(T) == 4 ? 1e-10 : 1e-20);
    if (Up) for (int t = threadIdx.x; t < nn; t += blockDim.x) { Up[t] = sm.Ure[t]; Up[nn + t] = sm.Uim[t]; }
    const double tau = p.rho ? p.tau[b] / p.rho[b] : p.tau[b];
    cx<T>* W = p.W + (size_t)b * nn;
    svt_weights_block(sm, n, tau, [&](int i, int j, double re, double im) { W[i + n * j] = mk<T>((T)re, (T)im); });
}
template <typename T>
__global__ void __launch_bounds__(kEigThreads) k_weights(SvtP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    weights_body<T>(p, blockIdx.x, smem);
}

// sigma_max^2 of the summed partial Gram `which` (0: X - Htrue, 1: Htrue); grid (1, batch)
template <typename T>
__global__ void __launch_bounds__(128) k_mc_conv(SvtP<T> p, int which) {
    extern __shared__ __align__(16) unsigned char smem[];
    JacobiSmem sm; sm.carve(smem, p.N);
    const int b = blockIdx.x, n = p.N, nn = n * n;
    const double* g = p.cgram + ((size_t)b * 2 + which) * p.nmc * 2 * nn;
    for (int t = threadIdx.x; t < nn; t += blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int c = 0; c < p.nmc; ++c) { re += g[(size_t)c * 2 * nn + 2 * t]; im += g[(size_t)c * 2 * nn + 2 * t + 1]; }
        sm.Are[t] = re; sm.Aim[t] = im;
    }
    __syncthreads();
    jacobi_hermitian_block(sm, n);
    if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int k = 0; k < n; ++k) mx = fmax(mx, sm.Are[k + n * k]);
        double* c = p.convd + (size_t)b * (p.imax + 1);
        if (which == 1) c[p.imax] = mx;
        else c[p.iter] = mx / c[p.imax];       // norm(X-Htrue)^2/norm(Htrue)^2   (mc_admm.m:28)
    }
}

// partial Gram of a global matrix chunk (used for the first SVT input and for Htrue)
template <typename T>
__global__ void __launch_bounds__(kThreads) k_gram_of(SvtP<T> p, const cx<T>* src, long long ld, double* dst, int slots_per_trial, int slot) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.y, chunk = blockIdx.x, N = p.N, RP = p.RP, MC = p.MC;
    const int c0 = chunk * MC, ncols = (p.M - c0) < MC ? (p.M - c0) : MC;
    T* Zre = reinterpret_cast<T*>(smem); T* Zim = Zre + (size_t)RP * MC;
    const cx<T>* s = src + (long long)b * ld + (size_t)c0 * N;
    for (int t = threadIdx.x; t < RP * MC; t += kThreads) {
        int r = t % RP, c = t / RP;
        cx<T> v = mk<T>(T(0), T(0));
        if (r < N && c < ncols) v = s[(size_t)c * N + r];
        Zre[t] = v.re; Zim[t] = v.im;
    }
    __syncthreads();
    gram_partial<T>(Zre, Zim, RP, N, ncols, dst + (((size_t)b * slots_per_trial + slot) * p.nmc + chunk) * 2 * N * N);
}

template <typename T, int MODE>
__device__ __forceinline__ void step_body(const SvtP<T>& p, int b, int chunk, unsigned char* smem) {
    const int N = p.N, RP = p.RP, MC = p.MC;
    const int c0 = chunk * MC, ncols = (p.M - c0) < MC ? (p.M - c0) : MC;
    const bool conv = (MODE == MODE_MCADMM) && p.convd != nullptr;
    T* Zre = reinterpret_cast<T*>(smem);
    T* Zim = Zre + (size_t)RP * MC;
    T* Nre = Zim + (size_t)RP * MC;        // next SVT input
    T* Nim = Nre + (size_t)RP * MC;
    T* Ere = Nim + (size_t)RP * MC;        // conv: X - Htrue
    T* Eim = Ere + (size_t)RP * MC;
    const int planes = (MODE == MODE_SVT) ? 2 : (conv ? 6 : 4);
    cx<T>* Ws = reinterpret_cast<cx<T>*>(Zre + (size_t)RP * MC * planes);
    const size_t off = (size_t)b * N * p.M + (size_t)c0 * N;
    const cx<T>* in = p.in + (long long)b * p.ld_in + (size_t)c0 * N;
    const T rho = (MODE == MODE_SVT) ? T(1) : (T)p.rho[b];
    const T irho = T(1) / rho;
    const cx<T>* Wg = p.W + (size_t)b * N * N;
    for (int t = threadIdx.x; t < N * N; t += kThreads) Ws[t] = Wg[t];
    for (int t = threadIdx.x; t < RP * MC; t += kThreads) {
        int r = t % RP, c = t / RP;
        T zr = 0, zi = 0;
        if (r < N && c < ncols) {
            size_t gi = (size_t)c * N + r;
            if (MODE == MODE_SVT) { cx<T> v = in[gi]; zr = v.re; zi = v.im; }
            else if (MODE == MODE_MCSVT) { cx<T> v = p.Ys[off + gi]; zr = v.re; zi = v.im; }                       // mc_svt.m:8
            else { cx<T> y = p.Ys[off + gi], z = p.Zs[off + gi]; zr = y.re - irho * z.re; zi = y.im - irho * z.im; }  // mc_admm.m:22
        }
        Zre[t] = zr; Zim[t] = zi;
        if (MODE != MODE_SVT) { Nre[t] = 0; Nim[t] = 0; }
        if (conv) { Ere[t] = 0; Eim[t] = 0; }
    }
    __syncthreads();
    const bool last = p.iter == p.imax - 1;
    // X = W Z, 2 rows x 2 columns per thread (two 8-byte reads of W and four of Z per 8 complex MACs), then the element-wise updates
    auto elem = [&](int r, int c, T xr, T xi) {
        size_t gi = (size_t)c * N + r;
        if (MODE == MODE_SVT) { p.out[(long long)b * p.ld_out + (size_t)(c0 + c) * N + r] = mk<T>(xr, xi); return; }
        const T om = p.omega[(long long)b * p.ld_omega + (size_t)(c0 + c) * N + r];
        cx<T> oh = in[gi];
        if (MODE == MODE_MCSVT) {
            cx<T> y = p.Ys[off + gi];
            T nr = y.re + rho * (oh.re - om * xr), ni = y.im + rho * (oh.im - om * xi);     // mc_svt.m:9
            p.Ys[off + gi] = mk<T>(nr, ni);
            Nre[c * RP + r] = nr; Nim[c * RP + r] = ni;
        } else {
            cx<T> z = p.Zs[off + gi];
            T d = T(1) / (om + rho);                                                         // mc_admm.m:11-17,24
            T yr = (oh.re + z.re + rho * xr) * d, yi = (oh.im + z.im + rho * xi) * d;
            T zr = z.re + rho * (xr - yr), zi = z.im + rho * (xi - yi);                      // mc_admm.m:26
            p.Ys[off + gi] = mk<T>(yr, yi); p.Zs[off + gi] = mk<T>(zr, zi);
            Nre[c * RP + r] = yr - irho * zr; Nim[c * RP + r] = yi - irho * zi;
            if (conv) { cx<T> ht = p.Htrue[(long long)b * p.ld_H + (size_t)(c0 + c) * N + r]; Ere[c * RP + r] = xr - ht.re; Eim[c * RP + r] = xi - ht.im; }
        }
        if (last) p.out[(long long)b * p.ld_out + (size_t)(c0 + c) * N + r] = mk<T>(xr, xi);
    };
    const int hN = (N + 1) / 2, hC = (ncols + 1) / 2;
    for (int t = threadIdx.x; t < hN * hC; t += kThreads) {
        const int r0 = 2 * (t % hN), ca = 2 * (t / hN);
        const int r1 = r0 + 1 < N ? r0 + 1 : r0, cb = ca + 1 < ncols ? ca + 1 : ca;
        T ar[2][2] = {{0, 0}, {0, 0}}, ai[2][2] = {{0, 0}, {0, 0}};                           // [row][column]
        for (int k = 0; k < N; ++k) {
            const cx<T> w0 = Ws[r0 + N * k], w1 = Ws[r1 + N * k];
            const T zar = Zre[ca * RP + k], zai = Zim[ca * RP + k], zbr = Zre[cb * RP + k], zbi = Zim[cb * RP + k];
            cmac<T>(ar[0][0], ai[0][0], w0.re, w0.im, zar, zai); cmac<T>(ar[1][0], ai[1][0], w1.re, w1.im, zar, zai);
            cmac<T>(ar[0][1], ai[0][1], w0.re, w0.im, zbr, zbi); cmac<T>(ar[1][1], ai[1][1], w1.re, w1.im, zbr, zbi);
        }
        elem(r0, ca, ar[0][0], ai[0][0]);
        if (r1 != r0) elem(r1, ca, ar[1][0], ai[1][0]);
        if (cb != ca) { elem(r0, cb, ar[0][1], ai[0][1]); if (r1 != r0) elem(r1, cb, ar[1][1], ai[1][1]); }
    }
    if (MODE == MODE_SVT) return;
    __syncthreads();
    gram_partial<T>(Nre, Nim, RP, N, ncols, p.gram + ((size_t)b * p.nmc + chunk) * 2 * N * N);
    if (conv) gram_partial<T>(Ere, Eim, RP, N, ncols, p.cgram + (((size_t)b * 2 + 0) * p.nmc + chunk) * 2 * N * N);
}
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) k_svt_step(SvtP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    step_body<T, MODE>(p, blockIdx.y, blockIdx.x, smem);
}
// One CTA per trial runs the whole solve (mc_svt.m:7-10 / mc_admm.m:20-27): per iteration the Jacobi solve of the Mr x Mr Gram matrix in shared
// memory, then the fused step over the trial's column chunks.  One launch per call instead of two per iteration; the state (a few hundred KB per
// trial) stays L2-resident between iterations, and co-resident CTAs overlap one trial's latency-bound eigen-solve with another's streaming step.
// The two phases share the same shared memory; partial Gram matrices and W travel through global memory as in the two-kernel form, so the
// arithmetic - and every result bit - is the same.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads, 4) k_svt_persist(SvtP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x;
    for (int it = 0; it < p.imax; ++it) {
        p.iter = it;
        weights_body<T>(p, b, smem);
        __syncthreads();
        for (int chunk = 0; chunk < p.nmc; ++chunk) {
            step_body<T, MODE>(p, b, chunk, smem);
            __syncthreads();
        }
    }
}

template <typename T>
static int run_svt_family(Handle* h, int mode, int mem, int N, int M, int batch, int imax,
                          const void* Htrue_, long long ld_H, const void* in_, long long ld_in,
                          const void* omega_, long long ld_omega, const double* tau_, const double* rho_,
                          void* out_, long long ld_out, void* conv_, long long ld_conv) {
    if (N <= 0 || M <= 0 || batch <= 0 || imax < 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if (!in_ || !tau_ || !out_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    if (mode != MODE_SVT && (!omega_ || !rho_)) return fail(h, JSTSP_E_ARG, "NULL buffer");
    if (conv_ && !Htrue_) return fail(h, JSTSP_E_ARG, "convergence_error needs Htrue");
    if (N > 64) return fail(h, JSTSP_E_UNSUPPORTED, "SVT kernels cover Mr <= 64 rows");
    const bool host = mem == JSTSP_HOST, want_conv = conv_ != nullptr;
    cudaStream_t st = h->stream;
    if (ld_out == 0) ld_out = (long long)N * M;
    SvtP<T> p{};
    p.N = N; p.M = M; p.RP = round_up8(N); p.imax = (mode == MODE_SVT) ? 1 : imax;
    int MC = 128;
    const int planes = (mode == MODE_SVT) ? 2 : (want_conv ? 6 : 4);
    auto smem_of = [&](int mc) { return (size_t)planes * p.RP * mc * sizeof(T) + sizeof(cx<T>) * (size_t)N * N; };
    while (MC > 16 && smem_of(MC) > 64 * 1024) MC /= 2;
    p.MC = MC; p.nmc = ceil_div(M, MC);
    const size_t NM = (size_t)N * M, NN2 = 2 * (size_t)N * N;
    int chunk_trials = batch;
    if (h->max_chunk > 0 && chunk_trials > h->max_chunk) chunk_trials = h->max_chunk;
    auto layout = [&](Arena& a, int nb, SvtP<T>& q) {
        q.W = a.take<cx<T>>((size_t)N * N * nb);
        q.gram = a.take<double>((size_t)nb * q.nmc * NN2);
        if (mode != MODE_SVT) { q.Ys = a.take<cx<T>>(NM * nb); q.Uprev = a.take<double>((size_t)nb * NN2); }
        if (mode == MODE_MCADMM) q.Zs = a.take<cx<T>>(NM * nb);
        if (want_conv) { q.cgram = a.take<double>((size_t)nb * 2 * q.nmc * NN2); q.convd = a.take<double>((size_t)nb * (imax + 1)); }
        if (host) {
            q.in = a.take<cx<T>>(ld_in ? NM * nb : NM);
            if (mode != MODE_SVT) q.omega = a.take<T>(ld_omega ? NM * nb : NM);
            if (want_conv) q.Htrue = a.take<cx<T>>(ld_H ? NM * nb : NM);
            q.tau = a.take<double>(nb);
            if (mode != MODE_SVT) q.rho = a.take<double>(nb);
            q.out = a.take<cx<T>>(NM * nb);
        }
    };
    size_t freeb = 0, totalb = 0;
    JSTSP_CUDA(h, cudaMemGetInfo(&freeb, &totalb));
    size_t budget = (size_t)((freeb + h->ws_bytes) * 0.7);
    for (;;) {
        Arena probe(nullptr, 0); SvtP<T> q = p; layout(probe, chunk_trials, q);
        if (probe.off <= budget || chunk_trials == 1) { int rc = ensure_workspace(h, probe.off); if (rc) return rc; break; }
        chunk_trials = (chunk_trials + 1) / 2;
    }
    const size_t sm_step = smem_of(MC), sm_gram = 2 * (size_t)p.RP * MC * sizeof(T), sm_j = JacobiSmem::bytes(N), sm_w = JacobiSmem::bytes(N) + 2 * sizeof(double) * (size_t)N * N + 16;
    int rc;
    if ((rc = set_smem(h, k_svt_step<T, MODE_SVT>, sm_step))) return rc;
    if ((rc = set_smem(h, k_svt_step<T, MODE_MCSVT>, sm_step))) return rc;
    if ((rc = set_smem(h, k_svt_step<T, MODE_MCADMM>, sm_step))) return rc;
    if ((rc = set_smem(h, k_gram_of<T>, sm_gram))) return rc;
    if ((rc = set_smem(h, k_weights<T>, sm_w))) return rc;
    if ((rc = set_smem(h, k_mc_conv<T>, sm_j))) return rc;
    JSTSP_CUDA(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), st));
    const size_t esz = sizeof(cx<T>);
    for (int b0 = 0; b0 < batch; b0 += chunk_trials) {
        const int nb = (batch - b0) < chunk_trials ? (batch - b0) : chunk_trials;
        Arena ar(h->ws, h->ws_bytes);
        SvtP<T> q = p;
        layout(ar, nb, q);
        q.ld_in = ld_in; q.ld_omega = ld_omega; q.ld_H = ld_H; q.ld_out = ld_out;
        if (host) {
            auto up = [&](const void* dst, const void* src, size_t per, long long ld, size_t el) -> cudaError_t {
                if (ld == 0) return cudaMemcpyAsync(const_cast<void*>(dst), src, per * el, cudaMemcpyHostToDevice, st);
                if ((size_t)ld == per) return cudaMemcpyAsync(const_cast<void*>(dst), (const char*)src + (size_t)b0 * ld * el, per * el * nb, cudaMemcpyHostToDevice, st);
                return cudaMemcpy2DAsync(const_cast<void*>(dst), per * el, (const char*)src + (size_t)b0 * ld * el, (size_t)ld * el, per * el, nb, cudaMemcpyHostToDevice, st);
            };
            JSTSP_CUDA(h, up(q.in, in_, NM, ld_in, esz));
            if (mode != MODE_SVT) JSTSP_CUDA(h, up(q.omega, omega_, NM, ld_omega, sizeof(T)));
            if (want_conv) JSTSP_CUDA(h, up(q.Htrue, Htrue_, NM, ld_H, esz));
            JSTSP_CUDA(h, cudaMemcpyAsync(const_cast<double*>(q.tau), tau_ + b0, sizeof(double) * nb, cudaMemcpyHostToDevice, st));
            if (mode != MODE_SVT) JSTSP_CUDA(h, cudaMemcpyAsync(const_cast<double*>(q.rho), rho_ + b0, sizeof(double) * nb, cudaMemcpyHostToDevice, st));
            if (q.ld_in) q.ld_in = NM; if (q.ld_omega) q.ld_omega = NM; if (q.ld_H) q.ld_H = NM;
            q.ld_out = NM;
        } else {
            q.in = (const cx<T>*)in_ + (long long)b0 * ld_in;
            q.omega = omega_ ? (const T*)omega_ + (long long)b0 * ld_omega : nullptr;
            q.Htrue = Htrue_ ? (const cx<T>*)Htrue_ + (long long)b0 * ld_H : nullptr;
            q.tau = tau_ + b0; q.rho = rho_ ? rho_ + b0 : nullptr;
            q.out = (cx<T>*)out_ + (long long)b0 * ld_out;
        }
        dim3 grid(q.nmc, nb);
        if (mode == MODE_SVT) {
            q.rho = nullptr;
            JSTSP_LAUNCH(h, PK_OTHER, (k_gram_of<T><<<grid, kThreads, sm_gram, st>>>(q, q.in, q.ld_in, q.gram, 1, 0)));
            JSTSP_LAUNCH(h, PK_EIG, (k_weights<T><<<nb, kEigThreads, sm_w, st>>>(q)));
            q.iter = 0;
            JSTSP_LAUNCH(h, PK_SVT_STEP, (k_svt_step<T, MODE_SVT><<<grid, kThreads, sm_step, st>>>(q)));
        } else {
            JSTSP_CUDA(h, cudaMemsetAsync(q.Ys, 0, esz * NM * nb, st));                       // Y = 0 (mc_svt.m:5, mc_admm.m:7)
            if (mode == MODE_MCADMM) JSTSP_CUDA(h, cudaMemsetAsync(q.Zs, 0, esz * NM * nb, st));
            JSTSP_CUDA(h, cudaMemsetAsync(q.gram, 0, sizeof(double) * (size_t)nb * q.nmc * NN2, st));
            if (imax == 0) JSTSP_CUDA(h, cudaMemsetAsync(q.out, 0, esz * NM * nb, st));
            if (want_conv) {
                JSTSP_LAUNCH(h, PK_OTHER, (k_gram_of<T><<<grid, kThreads, sm_gram, st>>>(q, q.Htrue, q.ld_H, q.cgram, 2, 1)));
                JSTSP_LAUNCH(h, PK_OTHER, (k_mc_conv<T><<<nb, 128, sm_j, st>>>(q, 1)));
            }
            const bool persist = !want_conv && imax > 0 && kEigThreads == kThreads && getenv("JSTSP_SVT_PERSIST_OFF") == nullptr;
            if (persist) {
                const size_t sm_p = sm_w > sm_step ? sm_w : sm_step;
                if (mode == MODE_MCSVT) {
                    if ((rc = set_smem(h, k_svt_persist<T, MODE_MCSVT>, sm_p))) return rc;
                    JSTSP_LAUNCH(h, PK_SVT_STEP, (k_svt_persist<T, MODE_MCSVT><<<nb, kThreads, sm_p, st>>>(q)));
                } else {
                    if ((rc = set_smem(h, k_svt_persist<T, MODE_MCADMM>, sm_p))) return rc;
                    JSTSP_LAUNCH(h, PK_SVT_STEP, (k_svt_persist<T, MODE_MCADMM><<<nb, kThreads, sm_p, st>>>(q)));
                }
            }
            for (int it = 0; it < imax && !persist; ++it) {
                q.iter = it;
                JSTSP_LAUNCH(h, PK_EIG, (k_weights<T><<<nb, kEigThreads, sm_w, st>>>(q)));
                if (mode == MODE_MCSVT) JSTSP_LAUNCH(h, PK_SVT_STEP, (k_svt_step<T, MODE_MCSVT><<<grid, kThreads, sm_step, st>>>(q)));
                else JSTSP_LAUNCH(h, PK_SVT_STEP, (k_svt_step<T, MODE_MCADMM><<<grid, kThreads, sm_step, st>>>(q)));
                if (want_conv) { JSTSP_LAUNCH(h, PK_OTHER, (k_mc_conv<T><<<nb, 128, sm_j, st>>>(q, 0))); }
            }
        }
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            long long ldo = ld_out;
            if ((size_t)ldo == NM || nb == 1) JSTSP_CUDA(h, cudaMemcpyAsync((char*)out_ + (size_t)b0 * ldo * esz, q.out, esz * NM * nb, cudaMemcpyDeviceToHost, st));
            else JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)out_ + (size_t)b0 * ldo * esz, (size_t)ldo * esz, q.out, NM * esz, NM * esz, nb, cudaMemcpyDeviceToHost, st));
            if (want_conv) {
                std::vector<double> tmp((size_t)nb * (imax + 1));
                JSTSP_CUDA(h, cudaMemcpyAsync(tmp.data(), q.convd, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost, st));
                JSTSP_CUDA(h, cudaStreamSynchronize(st));
                long long ldc = ld_conv ? ld_conv : imax;
                for (int bb = 0; bb < nb; ++bb)
                    for (int it = 0; it < imax; ++it) ((T*)conv_)[(size_t)(b0 + bb) * ldc + it] = (T)tmp[(size_t)bb * (imax + 1) + it];
            }
            JSTSP_CUDA(h, cudaStreamSynchronize(st));
        } else if (want_conv) {
            return fail(h, JSTSP_E_UNSUPPORTED, "mc_admm convergence_error output is only available for JSTSP_HOST calls");
        }
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

#define DISPATCH(call_f, call_d)                                      \
    if (!h) return JSTSP_E_ARG;                                       \
    JSTSP_CUDA(h, cudaSetDevice(h->device));                          \
    if (dtype == JSTSP_F32) return call_f;                            \
    if (dtype == JSTSP_F64) return call_d;                            \
    return fail(h, JSTSP_E_ARG, "unknown dtype");

extern "C" int jstsp_svt(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch,
                         const void* Y, long long ld_Y, const double* tau, void* X, long long ld_X) {
    DISPATCH((run_svt_family<float>(h, MODE_SVT, mem, Mr, Mt, batch, 1, nullptr, 0, Y, ld_Y, nullptr, 0, tau, nullptr, X, ld_X, nullptr, 0)),
             (run_svt_family<double>(h, MODE_SVT, mem, Mr, Mt, batch, 1, nullptr, 0, Y, ld_Y, nullptr, 0, tau, nullptr, X, ld_X, nullptr, 0)))
}

extern "C" int jstsp_mc_svt(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch, int imax,
                            const void* OH, long long ld_OH, const void* omega, long long ld_omega,
                            const double* tau, const double* rho, void* X, long long ld_X) {
    DISPATCH((run_svt_family<float>(h, MODE_MCSVT, mem, Mr, Mt, batch, imax, nullptr, 0, OH, ld_OH, omega, ld_omega, tau, rho, X, ld_X, nullptr, 0)),
             (run_svt_family<double>(h, MODE_MCSVT, mem, Mr, Mt, batch, imax, nullptr, 0, OH, ld_OH, omega, ld_omega, tau, rho, X, ld_X, nullptr, 0)))
}

extern "C" int jstsp_mc_admm(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch, int imax,
                             const void* Htrue, long long ld_H, const void* OH, long long ld_OH,
                             const void* omega, long long ld_omega, const double* tau, const double* rho,
                             void* X, long long ld_X, void* conv, long long ld_conv) {
    DISPATCH((run_svt_family<float>(h, MODE_MCADMM, mem, Mr, Mt, batch, imax, Htrue, ld_H, OH, ld_OH, omega, ld_omega, tau, rho, X, ld_X, conv, ld_conv)),
             (run_svt_family<double>(h, MODE_MCADMM, mem, Mr, Mt, batch, imax, Htrue, ld_H, OH, ld_OH, omega, ld_omega, tau, rho, X, ld_X, conv, ld_conv)))
}
