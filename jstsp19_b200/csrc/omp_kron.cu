// omp_kron.cu - OMP over the beamspace-delay Kronecker dictionary, never materialised.
//
//   jstsp_omp_kron(A, B, Y, m) == OMP(kron(B.', A), vec(Y), m)       (benchmark_algorithms/OMP.m:1-32)
//
// The reference's drivers hand the sparse solvers Phi = kron(B.', A), y = vec(Y)
// (plot_errorVSdelays.m:77-78, plot_errorVSframelength.m:78-79); at BASELINE config 2
// (Nt = Nr = 64, 4x oversampled grids) Phi would be 8192 x 262144 complex - 32 GiB in fp64 - so the
// factors are the input here.  Column j = g + G*p (0-based, column-major over the G x P unknown) of
// Phi is vec(A(:,g) * B(p,:)), hence
//     Phi' * r   = vec(A^H R B^H)            (OMP.m:17; R = reshape(r, N, M))
//     atom j     = A(:,g) B(p,:)             (OMP.m:18)
// and everything else (growing least squares, residual, scatter) is OMP.m unchanged.
//
// Per iteration, two launches:
//   k_kron_corr   grid (P/32, batch): T_c = R B_c^H (N x 32), C_c = A^H T_c (G x 32) out of shared
//                 memory, |C|^2 arg-max with the first-maximum rule and the runner-up -> one
//                 candidate per (chunk, trial).  C is never written.  FP32 FMA bound.
//   k_kron_update grid (batch): final arg-max, duplicate handling, then the least-squares re-solve through an incrementally
//                 grown fp64 Cholesky factor of the Gram matrix of the picked atoms (rank-one atoms make a Gram entry N + M
//                 terms, so neither the atoms nor an orthonormal basis are stored) and r = Y - sum_j x_j a_j b_j.
// k_kron_finish scatters x_hat (OMP.m:27-31).
// For fp32 operands at N = 64 the correlation runs on the tensor cores instead (k_kron_corr_tc below): a tf32 screen
// followed by an fp64 re-evaluation of every candidate inside the rounding band, with k_kron_corr as the fallback.
#include "common.cuh"
#include "umma_prims.cuh"
#include <cstdlib>

namespace jstsp {

constexpr int KC_PC = 32;      // delay-beam columns (p) per CTA
constexpr int KC_RB = 64;      // row block (n in phase 1, g in phase 2)
constexpr int KC_KT = 32;      // contraction tile

template <typename T>
struct KronP {
    int N, M, G, P, m, t, nchunk;
    const cx<T>* A;  long long ld_A;     // N x G
    const cx<T>* AH; long long ld_AH;    // G x N (conjugate transpose, built once per call)
    const cx<T>* B;  long long ld_B;     // P x M
    const cx<T>* Y;  long long ld_Y;     // N x M
    cx<T>* res;                          // [b][N*M] residual
    cx<double>* K;                       // [b][m*m] inverse of the Cholesky factor L of the Gram matrix of the picked atoms (row-major, lower)
    cx<double>* rhsv;                    // [b][m]   T' v
    cx<double>* yv;                      // [b][m]   y = K T' v
    cx<double>* xv;                      // [b][m]   x = K' y : coefficients of the unique picks
    cx<T>* Asel;                         // [b][m][N] A(:, g_j)
    cx<T>* Bsel;                         // [b][m][M] B(p_j, :)
    int* state;                          // [b][2 + 3m]: nuniq, amb, sel[m], uniq_idx[m], mult[m]
    double* cand_val;                    // [b][nchunk][2] best, second
    int* cand_idx;                       // [b][nchunk]    index of the best
    int* cand_idx2;                      // [b][nchunk]    index of the runner-up of the chunk
    float* abmax;                        // [b] (max |A entry|)(max |B entry|): scale of the rounding error of the tf32 screen
    double margin_tol;
    // tensor-core screening path (fp32, N == 64): packed K-major operands and per-tile candidates
    float* Bt;                           // [nB][P][2M]   row p = B(p,:) interleaved (re,im)
    float* Ach;                          // [nA][2G][2N]  rows (g,0) = (Ar,Ai)(:,g), (g,1) = (-Ai,Ar)(:,g)
    float* Rp;                           // [b][2N][2M]   rows (n,0) = (Rr,Ri)(n,:), (n,1) = (Ri,-Rr)(n,:)
    float* row_v1;                       // [b][P] largest |C(:,p)|^2 of row p (tf32 accuracy)
    int* row_g1;                         // [b][P] its g
    float* row_v2;                       // [b][P] runner-up value of the row (its index is not kept)
    int ntile;
    int* flag;                           // [b] 1 = the tf32 screen could not isolate the maximum: redo this iteration with the fp32 kernel
    int screen;                          // k_kron_update: 1 = candidates come from the tf32 screen, 0 = from k_kron_corr
    float band;                          // relative width of the candidate band on |C|^2
    long long* dbg;                      // developer hook (jstsp_debug_buffer): per-CTA wait-cycle counters of k_kron_corr_tc
    // outputs
    cx<T>* x_hat; long long ld_x;
    int* index_set;
    cx<T>* x_sel;
    cx<T>* r_out; long long ld_r;
    int* ambiguous;
};

// rows owned by a thread inside a 64-row block: {2q, 2q+1, 32+2q, 32+2q+1} - two 16-byte (fp32)
// shared-memory reads at a 16-byte stride across the quarter-warp, i.e. conflict-free.
__device__ __forceinline__ int kc_row(int q, int i) { return (i >> 1) * 32 + 2 * q + (i & 1); }

template <typename T>
__global__ void __launch_bounds__(256) k_kron_setup(KronP<T> p, int nA, int nB) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t NM = (size_t)p.N * p.M;
    const cx<T>* Y = p.Y + (long long)b * p.ld_Y;
    cx<T>* r = p.res + (size_t)b * NM;
    for (size_t i = tid; i < NM; i += 256) r[i] = Y[i];                       // r = v (OMP.m:10)
    int* st = p.state + (size_t)b * (2 + 3 * p.m);
    for (int i = tid; i < 2 + 3 * p.m; i += 256) st[i] = 0;
    if (b < nA) {                                                             // AH(g,n) = conj(A(n,g))
        const cx<T>* A = p.A + (long long)b * p.ld_A;
        cx<T>* AH = const_cast<cx<T>*>(p.AH) + (long long)b * p.ld_AH;
        for (int i = tid; i < p.N * p.G; i += 256) { const int g = i % p.G, n = i / p.G; AH[i] = conj(A[n + (size_t)p.N * g]); }
        if (p.Ach) {
            float* ac = p.Ach + (size_t)b * 4 * p.N * p.G;
            for (int i = tid; i < p.N * p.G; i += 256) {
                const int n = i % p.N, g = i / p.N; const cx<T> a = A[i];
                *reinterpret_cast<float2*>(ac + (size_t)(2 * g) * 2 * p.N + 2 * n) = make_float2((float)a.re, (float)a.im);
                *reinterpret_cast<float2*>(ac + (size_t)(2 * g + 1) * 2 * p.N + 2 * n) = make_float2(-(float)a.im, (float)a.re);
            }
        }
    }
    if (p.flag && tid == 0) p.flag[b] = 0;
    if (p.abmax) {                                                            // entrywise maxima of this trial's A and B
        __shared__ float s_am[8], s_bm[8];
        const cx<T>* A = p.A + (long long)b * p.ld_A; const cx<T>* B = p.B + (long long)b * p.ld_B;
        float am = 0.f, bm = 0.f;
        for (int i = tid; i < p.N * p.G; i += 256) am = fmaxf(am, (float)(A[i].re * A[i].re + A[i].im * A[i].im));
        for (size_t i = tid; i < (size_t)p.P * p.M; i += 256) bm = fmaxf(bm, (float)(B[i].re * B[i].re + B[i].im * B[i].im));
        for (int o = 16; o > 0; o >>= 1) { am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o)); bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o)); }
        if (tid % 32 == 0) { s_am[tid / 32] = am; s_bm[tid / 32] = bm; }
        __syncthreads();
        if (tid == 0) { for (int w = 1; w < 8; ++w) { am = fmaxf(am, s_am[w]); bm = fmaxf(bm, s_bm[w]); } p.abmax[b] = sqrtf(am) * sqrtf(bm); }
    }
    if (p.Bt && b < nB) {
        const cx<T>* B = p.B + (long long)b * p.ld_B;
        float* bt = p.Bt + (size_t)b * 2 * p.P * p.M;
        for (size_t i = tid; i < (size_t)p.P * p.M; i += 256) {
            const int pp = (int)(i % p.P), mm = (int)(i / p.P); const cx<T> v = B[i];
            *reinterpret_cast<float2*>(bt + (size_t)pp * 2 * p.M + 2 * mm) = make_float2((float)v.re, (float)v.im);
        }
    }
}

template <typename T>
__device__ __forceinline__ void kron_corr_item(const KronP<T>& p, int chunk, int b, unsigned char* smem) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, M = p.M, G = p.G, P = p.P;
    const int Npad = ceil_div(N, KC_KT) * KC_KT;
    cx<T>* Tc = reinterpret_cast<cx<T>*>(smem);                   // [Npad][PC]
    cx<T>* Lt = Tc + (size_t)Npad * KC_PC;                        // [KT][RB]  left operand tile
    cx<T>* Bt = Lt + KC_KT * KC_RB;                               // [KT][PC]
    __shared__ double s_best[8], s_second[8];
    const cx<T>* R = p.res + (size_t)b * N * M;
    const cx<T>* B = p.B + (long long)b * p.ld_B;
    const cx<T>* AH = p.AH + (long long)b * p.ld_AH;
    const int p0 = chunk * KC_PC;
    const int q = tid & 15, tp = tid >> 4;                         // rows kc_row(q,0..3), columns 2tp, 2tp+1
    const cx<T> zero = mk<T>(T(0), T(0));

    // ---- phase 1: T_c(n, p) = sum_m R(n, m) conj(B(p0+p, m)) ----
    for (int n0 = 0; n0 < Npad; n0 += KC_RB) {
        T ar[4][2], ai[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ar[i][0] = ar[i][1] = ai[i][0] = ai[i][1] = T(0); }
        for (int m0 = 0; m0 < M; m0 += KC_KT) {
            __syncthreads();
            for (int e = tid; e < KC_KT * KC_RB; e += 256) {
                const int k = e / KC_RB, i = e % KC_RB, n = n0 + i, mm = m0 + k;
                Lt[e] = (n < N && mm < M) ? R[n + (size_t)N * mm] : zero;
            }
            for (int e = tid; e < KC_KT * KC_PC; e += 256) {
                const int k = e / KC_PC, j = e % KC_PC, pp = p0 + j, mm = m0 + k;
                Bt[e] = (pp < P && mm < M) ? B[pp + (size_t)P * mm] : zero;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < KC_KT; ++k) {
                cx<T> l[4], r2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) l[i] = Lt[k * KC_RB + kc_row(q, i)];
                r2[0] = Bt[k * KC_PC + 2 * tp]; r2[1] = Bt[k * KC_PC + 2 * tp + 1];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) cmac<T>(ar[i][j], ai[i][j], l[i].re, l[i].im, r2[j].re, -r2[j].im);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + kc_row(q, i);
            if (n < Npad) { Tc[(size_t)n * KC_PC + 2 * tp] = mk<T>(ar[i][0], ai[i][0]); Tc[(size_t)n * KC_PC + 2 * tp + 1] = mk<T>(ar[i][1], ai[i][1]); }
        }
    }
    // ---- phase 2: C_c(g, p) = sum_n AH(g, n) T_c(n, p); arg-max of |C|^2, lowest j = g + G p on ties ----
    double best = -1.0, second = -1.0; long long bidx = 0x7fffffffffffLL, sidx = 0x7fffffffffffLL;
    for (int g0 = 0; g0 < G; g0 += KC_RB) {
        T ar[4][2], ai[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ar[i][0] = ar[i][1] = ai[i][0] = ai[i][1] = T(0); }
        for (int n0 = 0; n0 < Npad; n0 += KC_KT) {
            __syncthreads();
            for (int e = tid; e < KC_KT * KC_RB; e += 256) {
                const int k = e / KC_RB, i = e % KC_RB, g = g0 + i, n = n0 + k;
                Lt[e] = (g < G && n < N) ? AH[g + (size_t)G * n] : zero;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < KC_KT; ++k) {
                cx<T> l[4], r2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) l[i] = Lt[k * KC_RB + kc_row(q, i)];
                r2[0] = Tc[(size_t)(n0 + k) * KC_PC + 2 * tp]; r2[1] = Tc[(size_t)(n0 + k) * KC_PC + 2 * tp + 1];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) cmac<T>(ar[i][j], ai[i][j], l[i].re, l[i].im, r2[j].re, r2[j].im);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int g = g0 + kc_row(q, i), pp = p0 + 2 * tp + j;
                if (g < G && pp < P) {
                    const double mag = (double)ar[i][j] * ar[i][j] + (double)ai[i][j] * ai[i][j];
                    const long long jj = g + (long long)G * pp;
                    if (mag > best || (mag == best && jj < bidx)) { second = best; sidx = bidx; best = mag; bidx = jj; }
                    else if (mag > second || (mag == second && jj < sidx)) { second = mag; sidx = jj; }
                }
            }
    }
    // ---- reduce (best, second, index) over the CTA ----
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o), osi = __shfl_xor_sync(0xffffffffu, sidx, o);
        // merge (best, runner-up) of the two lanes: the two largest of the four records, lowest index first on equal values
        if (ob > best || (ob == best && oi < bidx)) {
            if (best > os || (best == os && bidx < osi)) { second = best; sidx = bidx; } else { second = os; sidx = osi; }
            best = ob; bidx = oi;
        } else if (ob > second || (ob == second && oi < sidx)) { second = ob; sidx = oi; }
    }
    __shared__ long long s_lidx[8], s_lidx2[8];
    if (lane == 0) { s_best[warp] = best; s_second[warp] = second; s_lidx[warp] = bidx; s_lidx2[warp] = sidx; }
    __syncthreads();
    if (tid == 0) {
        double gb = -1.0, gs = -1.0; long long gi = 0x7fffffffffffLL, gi2 = 0x7fffffffffffLL;
        for (int w = 0; w < 8; ++w) {
            const double wv[2] = {s_best[w], s_second[w]}; const long long wi[2] = {s_lidx[w], s_lidx2[w]};
            for (int k = 0; k < 2; ++k) {
                if (wv[k] > gb || (wv[k] == gb && wi[k] < gi)) { gs = gb; gi2 = gi; gb = wv[k]; gi = wi[k]; }
                else if (wv[k] > gs || (wv[k] == gs && wi[k] < gi2)) { gs = wv[k]; gi2 = wi[k]; }
            }
        }
        p.cand_val[((size_t)b * p.nchunk + chunk) * 2 + 0] = gb;
        p.cand_val[((size_t)b * p.nchunk + chunk) * 2 + 1] = gs;
        p.cand_idx[(size_t)b * p.nchunk + chunk] = (int)gi;
        p.cand_idx2[(size_t)b * p.nchunk + chunk] = (int)gi2;
    }
    __syncthreads();
}

// grid-stride over (chunk, trial) items: as the fallback of the tensor-core screen it is launched with a few CTAs per SM and
// skips every trial whose flag is clear, so an iteration without undecided trials costs a few microseconds.
template <typename T>
__global__ void __launch_bounds__(256) k_kron_corr(KronP<T> p, int nb) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int items = p.nchunk * nb;
    for (int w = blockIdx.x; w < items; w += gridDim.x) {
        const int b = w / p.nchunk;
        if (p.flag && !p.flag[b]) continue;
        kron_corr_item<T>(p, w % p.nchunk, b, smem);
    }
}


// residual -> K-major operand image of the tensor-core screen: Rp rows (n,0) = (Rr,Ri)(n,:), (n,1) = (Ri,-Rr)(n,:);
// 32 columns m per CTA through a padded shared-memory tile so that both the read (along n) and the write (along m) coalesce.
template <typename T>
__global__ void __launch_bounds__(256) k_kron_pack_r(KronP<T> p) {
    __shared__ float2 tl[32][65];
    const int b = blockIdx.y, m0 = blockIdx.x * 32, tid = threadIdx.x, N = p.N, M = p.M;      // N == 64 on this path
    const cx<T>* r = p.res + (size_t)b * N * M;
    for (int e = tid; e < 32 * 64; e += 256) {
        const int n = e & 63, k = e >> 6;
        if (m0 + k < M) { const cx<T> x = r[n + (size_t)N * (m0 + k)]; tl[k][n] = make_float2((float)x.re, (float)x.im); }
    }
    __syncthreads();
    float* rp = p.Rp + (size_t)b * 4 * N * M;
    for (int e = tid; e < 64 * 2 * 32; e += 256) {
        const int k = e & 31, row = e >> 5, n = row >> 1;
        if (m0 + k < M) {
            const float2 x = tl[k][n];
            *reinterpret_cast<float2*>(rp + (size_t)row * 2 * M + 2 * (m0 + k)) = (row & 1) ? make_float2(x.y, -x.x) : x;
        }
    }
}

// ---- tcgen05 screening kernel (fp32 storage, tf32 products, N == 64) -----------------------------------------------------------
// One CTA per (128 delay-beam indices p, trial).  All operands are K-major SWIZZLE_128B tiles [128 rows][32 floats] delivered by
// TMA tensor copies; complex arithmetic rides on real MMAs through the sign-embedded small operands:
//   GEMM 1   D1[p][(n,c)] = sum_{k=(m,c')} Bt[p][k] Rp[(n,c)][k]      = (T_re, T_im)(n,p),  T = R B^H        K = 2M
//   epilogue the 128 threads copy their D1 row (= row p of T^T, interleaved) into a K-major tile image in shared memory
//   GEMM 2   D2[p][(g,c)] = sum_{k=(n,c')} Tt[p][k] Ach[(g,c)][k]     = (C_re, C_im)(g,p),  C = A^H T        K = 2N = 128
//   epilogue |C|^2, per-thread / per-CTA top-2 (with indices) and third value -> one candidate record per tile.
// tf32 keeps ~10 mantissa bits, so the result only SCREENS: k_kron_update re-evaluates every candidate inside the error band in
// fp64 and picks among them (or hands the iteration to the fp32 kernel when a tile holds more than two in-band entries).
namespace kt {
constexpr int PT = 128, TILE = 16384, SLOT = 2 * TILE, NST = 3, THREADS = 320, NACC = 3;
constexpr size_t SMEM = (size_t)NST * SLOT + 8 * TILE + 512;
}  // namespace kt

__global__ void __launch_bounds__(kt::THREADS, 1) k_kron_corr_tc(KronP<float> p, const __grid_constant__ CUtensorMap mapBt, const __grid_constant__ CUtensorMap mapR,
                                                                const __grid_constant__ CUtensorMap mapA, int b_shared, int a_shared, int nb) {
    using namespace kt;
    using namespace um;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    unsigned char* Tt = smem + NST * SLOT;                                                    // two images of T^T (4 K-slice tiles each)
    uint64_t* bars = reinterpret_cast<uint64_t*>(Tt + 8 * TILE);
    uint64_t *full = bars, *empty = bars + NST, *d1_full = bars + 2 * NST, *t_ready = d1_full + 1, *d2_full = t_ready + 1, *d2_empty = d2_full + NACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + NACC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KS1 = (2 * p.M) / 32, NGRP = (2 * p.G) / 128;
    const int total = p.ntile * nb;
    const int nitem = blockIdx.x < total ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(d1_full, 1); mbar_init(t_ready, 128);
        for (int r = 0; r < NACC; ++r) { mbar_init(&d2_full[r], 1); mbar_init(&d2_empty[r], 128); }
        mbar_fence_init();
    }
    if (warp == 4) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    // Persistent: CTA c takes work items c, c + gridDim.x, ... (item = trial * ntile + tile).  Tensor-pipe order
    //     G1(0) | G1(1) G2(0) | G1(2) G2(1) | ...
    // so the copy of D1 into the T image of item k+1 (warps 0-3) and the arg-max scan of item k (warps 6-9) both run under MMAs.
    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            int gi = 0;
            long long w_empty = 0, t0 = clock64();
            auto load = [&](int k, bool second) {
                const int w = blockIdx.x + k * gridDim.x, b = w / p.ntile, p0 = (w % p.ntile) * PT;
                const int bb = b_shared ? 0 : b, aa = a_shared ? 0 : b;
                const int n = second ? 2 * NGRP : KS1;
                for (int i = 0; i < n; ++i, ++gi) {
                    const int slot = gi % NST, use = gi / NST;
                    if (use > 0) { const long long tq = clock64(); mbar_wait(&empty[slot], (use - 1) & 1); w_empty += clock64() - tq; }
                    mbar_expect_tx(&full[slot], SLOT);
                    unsigned char* dst = ring + slot * SLOT;
                    if (!second) {
                        tma_3d(dst, &mapBt, 32 * i, p0, bb, &full[slot]);
                        tma_3d(dst + TILE, &mapR, 32 * i, 0, b, &full[slot]);
                    } else {
                        const int q = i >> 1, u = i & 1;
                        tma_3d(dst, &mapA, 32 * (2 * u), 128 * q, aa, &full[slot]);
                        tma_3d(dst + TILE, &mapA, 32 * (2 * u + 1), 128 * q, aa, &full[slot]);
                    }
                }
            };
            if (nitem > 0) load(0, false);
            for (int k = 0; k < nitem; ++k) { if (k + 1 < nitem) load(k + 1, false); load(k, true); }
            if (p.dbg) { p.dbg[blockIdx.x * 8 + 5] = w_empty; p.dbg[blockIdx.x * 8 + 6] = clock64() - t0; }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(128, 128);
            int gi = 0, gq = 0;
            long long w_full = 0, w_tr = 0, w_d2e = 0, t0 = clock64(), tq;
            auto gemm1 = [&]() {
                for (int i = 0; i < KS1; ++i, ++gi) {
                    const int slot = gi % NST;
                    tq = clock64(); mbar_wait(&full[slot], (gi / NST) & 1); w_full += clock64() - tq;
                    tc_fence_after();
                    const uint32_t a = smem_u32(ring + slot * SLOT), bq = a + TILE;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_tf32(tm, desc_k128(a, ks), desc_k128(bq, ks), idesc, (i | ks) ? 1u : 0u);
                    umma_commit(&empty[slot]);
                }
                umma_commit(d1_full);
            };
            if (nitem > 0) gemm1();
            for (int k = 0; k < nitem; ++k) {
                tq = clock64(); mbar_wait(t_ready, k & 1); w_tr += clock64() - tq;   // D1 of item k drained into T image k & 1
                tc_fence_after();
                if (k + 1 < nitem) gemm1();
                const uint32_t tt = smem_u32(Tt + (k & 1) * 4 * TILE);
                for (int q = 0; q < NGRP; ++q, ++gq) {
                    const int r = gq % NACC;
                    if (gq >= NACC) { tq = clock64(); mbar_wait(&d2_empty[r], ((gq / NACC) - 1) & 1); w_d2e += clock64() - tq; tc_fence_after(); }
                    const uint32_t d2 = tm + 128 * (1 + r);
                    for (int u = 0; u < 2; ++u, ++gi) {
                        const int slot = gi % NST;
                        tq = clock64(); mbar_wait(&full[slot], (gi / NST) & 1); w_full += clock64() - tq;
                        tc_fence_after();
                        const uint32_t bq = smem_u32(ring + slot * SLOT);
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                umma_tf32(d2, desc_k128(tt + (2 * u + hf) * TILE, ks), desc_k128(bq + hf * TILE, ks), idesc, (u | hf | ks) ? 1u : 0u);
                        umma_commit(&empty[slot]);
                    }
                    umma_commit(&d2_full[r]);
                }
            }
            if (p.dbg) { long long* o = p.dbg + blockIdx.x * 8; o[0] = clock64() - t0; o[1] = w_full; o[2] = w_tr; o[3] = w_d2e; o[4] = nitem; }
        }
    } else if (warp < 4) {
        // ===== staging warps: D1 row (= row p of T^T, (re,im) interleaved along n) -> K-major SWIZZLE_128B image =====
        const int pl = tid;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int k = 0; k < nitem; ++k) {
            mbar_wait(d1_full, k & 1);
            tc_fence_after();
            unsigned char* img = Tt + (k & 1) * 4 * TILE;
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t v[32];
                tmem_ld32_nowait(tm + lane_base + 32 * c4, v);
                tmem_ld_wait();
                unsigned char* row = img + c4 * TILE + pl * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(row + ((j ^ (pl & 7)) << 4)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            fence_async_smem();
            tc_fence_before();
            mbar_arrive(t_ready);
        }
    } else {
        // ===== scan warps (6-9): thread = TMEM lane = local p; |C|^2 and the running top three =====
        const int quarter = warp & 3, pl = quarter * 32 + lane;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        int gq = 0;
        for (int k = 0; k < nitem; ++k) {
            const int w = blockIdx.x + k * gridDim.x, b = w / p.ntile, tile = w % p.ntile;
            const int pp = tile * PT + pl;
            const bool valid = pp < p.P;
            // branch-free running maxima, four independent chains (g mod 4): largest value with its index and the runner-up VALUE of
            // each chain.  Entries that lose inside their chain only ever matter through that value (-> the tile's "third").
            float cv1[4], cv2[4]; int ci1[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { cv1[c] = -1.f; cv2[c] = -1.f; ci1[c] = 0x7fffffff; }
            for (int q = 0; q < NGRP; ++q, ++gq) {
                const int r = gq % NACC;
                mbar_wait(&d2_full[r], (gq / NACC) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint32_t v[32];
                    tmem_ld32_nowait(tm + lane_base + 128 * (1 + r) + 32 * c4, v);
                    tmem_ld_wait();
                    if (valid) {
                        const int jb = 64 * q + 16 * c4;
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float re = __uint_as_float(v[2 * e]), im = __uint_as_float(v[2 * e + 1]);
                            const float mg = fmaf(re, re, im * im);
                            const int c = e & 3;
                            const bool gt = mg > cv1[c];                              // strict: the first (lowest-index) maximum stays
                            cv2[c] = fmaxf(cv2[c], gt ? cv1[c] : mg);
                            ci1[c] = gt ? jb + e : ci1[c];
                            cv1[c] = gt ? mg : cv1[c];
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&d2_empty[r]);
            }
            if (valid) {                                                          // row record: maximum (first index on ties), runner-up value
                float v1 = cv1[0], v2 = cv2[0]; int g1 = ci1[0];
#pragma unroll
                for (int c = 1; c < 4; ++c) {
                    const bool gt = cv1[c] > v1 || (cv1[c] == v1 && ci1[c] < g1);
                    v2 = fmaxf(fmaxf(v2, cv2[c]), gt ? v1 : cv1[c]);
                    g1 = gt ? ci1[c] : g1; v1 = gt ? cv1[c] : v1;
                }
                p.row_v1[(size_t)b * p.P + pp] = v1; p.row_g1[(size_t)b * p.P + pp] = g1; p.row_v2[(size_t)b * p.P + pp] = v2;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tm, 512);
}

constexpr int KU_T = 512, KU_W = KU_T / 32;      // threads / warps of k_kron_update (latency-bound: more loads in flight per trial)

template <typename T>
__global__ void __launch_bounds__(KU_T) k_kron_update(KronP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, G = p.G, P = p.P, m = p.m, t = p.t;
    const size_t NM = (size_t)N * p.M;
    __shared__ double s_red[KU_W][2];
    __shared__ int s_pick, s_dup;
    cx<T>* r = p.res + (size_t)b * NM;
    int* st = p.state + (size_t)b * (2 + 3 * m);
    int* sel = st + 2; int* uniq_idx = sel + m; int* mult = uniq_idx + m;
    const int nu = st[0];
    if (p.flag && !p.screen && !p.flag[b]) return;                            // fallback launch: only trials the screen left undecided (the flag is rewritten by the next screen)
    // ---- candidate list: every entry that the correlation pass could not rule out as the first maximum of |Phi' r|^2 (OMP.m:17).
    //      More than one candidate -> all of them are re-evaluated in fp64 (a_g' R b_p', N M terms) and the first maximum of the
    //      exact values is the pick, so the support equals an fp64 evaluation's whichever pass produced the list. ----
    constexpr int MAXC = 128, MAXR = 8;
    __shared__ int s_cand[MAXC];
    __shared__ double s_cmag[MAXC];
    __shared__ int s_rows[MAXR];
    __shared__ int s_nc, s_nrow, s_fail;
    __shared__ float s_vmax, s_band;
    __shared__ double s_lim2;
    if (tid == 0) { s_nc = 0; s_nrow = 0; s_fail = 0; s_lim2 = 0.0; }
    if (p.screen) {
        // tf32 screen: every row maximum within the band of the largest value; a row whose runner-up is in the band too is recomputed whole
        // (exactly), since the screen does not keep that index.  The band is the larger of p.band and four times the screen's own rounding
        // error relative to |C|max: 16 x 2^-10 (tf32 truncation of three operands, random-sign accumulation with a wide margin) x |R|_F x
        // max|A| x max|B|.  A residual that is mostly noise therefore widens the band by itself; a list that overflows hands the
        // iteration to the fp32 pass below.
        const float* rv1 = p.row_v1 + (size_t)b * P; const int* rg1 = p.row_g1 + (size_t)b * P; const float* rv2 = p.row_v2 + (size_t)b * P;
        float vm = -1.f; double r2 = 0.0;
        for (int q = tid; q < P; q += KU_T) vm = fmaxf(vm, rv1[q]);
        for (size_t i = tid; i < NM; i += KU_T) { const cx<T> x = r[i]; r2 += (double)x.re * x.re + (double)x.im * x.im; }
        for (int o = 16; o > 0; o >>= 1) { vm = fmaxf(vm, __shfl_xor_sync(0xffffffffu, vm, o)); r2 += __shfl_xor_sync(0xffffffffu, r2, o); }
        if (lane == 0) { s_red[warp][0] = vm; s_red[warp][1] = r2; }
        __syncthreads();
        if (tid == 0) {
            float a = -1.f; double rr = 0.0;
            for (int w = 0; w < KU_W; ++w) { a = fmaxf(a, (float)s_red[w][0]); rr += s_red[w][1]; }
            s_vmax = a;
            const float e = 16.f * 9.765625e-4f * (float)sqrt(rr) * (p.abmax ? p.abmax[b] : 1.f);
            const float rel = a > 0.f ? 4.f * e * rsqrtf(a) : 1.f;
            s_band = fminf(fmaxf(p.band, rel), 0.5f);
        }
        __syncthreads();
        const float vmax = s_vmax, lim = vmax * (1.f - s_band);
        for (int q = tid; q < P; q += KU_T) {
            const bool whole = rv2[q] >= lim && rv2[q] >= 0.f;
            if (whole) { const int k = atomicAdd(&s_nrow, 1); if (k < MAXR) s_rows[k] = q; }
            else if (rv1[q] >= lim) { const int k = atomicAdd(&s_nc, 1); if (k < MAXC) s_cand[k] = rg1[q] + G * q; }
        }
        __syncthreads();
        if (tid == 0) {
            const int failed = (s_nc > MAXC || s_nrow > MAXR || !(vmax > 0.f)) ? 1 : 0;     // overflow / all-zero / non-finite screen: the fp32 kernel decides
            if (p.dbg) {
                unsigned long long* ctr = reinterpret_cast<unsigned long long*>(p.dbg) + 1200;
                atomicAdd(ctr + 0, 1ull); atomicAdd(ctr + 1, (unsigned long long)failed); atomicAdd(ctr + 2, (unsigned long long)s_nc); atomicAdd(ctr + 3, (unsigned long long)s_nrow);
            }
            s_fail = failed; p.flag[b] = failed;
            s_lim2 = (double)lim * (1.0 - (double)s_band);
        }
        __syncthreads();
        if (s_fail) return;
    } else {
        // fp32 / fp64 correlation pass (k_kron_corr): best and runner-up of every chunk, with their indices.  Every record within the pass's own
        // rounding band of the largest value (1e-4 relative on |C|^2 for fp32 sums of N M terms, 1e-11 for fp64, or the caller's margin_tol if
        // wider) is a candidate.  (A chunk keeps two records, so a THIRD entry of one chunk inside a 1e-4 band would go unseen.)
        if (tid == 0) {
            double gb = -1.0;
            for (int c = 0; c < p.nchunk; ++c) gb = fmax(gb, p.cand_val[((size_t)b * p.nchunk + c) * 2]);
            const double band = fmax(p.margin_tol, sizeof(T) == 4 ? 1e-4 : 1e-11), lim = gb * (1.0 - band);
            int nc = 0;
            for (int c = 0; c < p.nchunk; ++c) {                              // chunks ascend in p, i.e. in j
                const double wb = p.cand_val[((size_t)b * p.nchunk + c) * 2], wsd = p.cand_val[((size_t)b * p.nchunk + c) * 2 + 1];
                if (wb >= lim && nc < MAXC) s_cand[nc++] = p.cand_idx[(size_t)b * p.nchunk + c];
                if (wsd >= lim && wsd >= 0.0 && nc < MAXC) s_cand[nc++] = p.cand_idx2[(size_t)b * p.nchunk + c];
            }
            s_nc = nc; s_nrow = 0;
            if (nc == 0) { s_cand[0] = 0; s_nc = 1; }                         // non-finite correlations: index 1, like max() of a NaN vector would not be - flagged by the caller's checks
        }
        __syncthreads();
    }
    {
        const int nc0 = s_nc < MAXC ? s_nc : MAXC, nrow = s_nrow < MAXR ? s_nrow : MAXR;
        if (nc0 + nrow > 1 || nrow > 0) {
            for (int c = 0; c < nc0; ++c) {                                      // exact |a_g' R b_p'|^2 (OMP.m:17 for one column)
                const int j = s_cand[c], g = j % G, pq = j / G;
                const cx<T>* a = p.A + (long long)b * p.ld_A + (size_t)N * g;
                const cx<T>* bq = p.B + (long long)b * p.ld_B + pq;
                double re = 0.0, im = 0.0;
                for (size_t i = tid; i < NM; i += KU_T) {
                    const int n = (int)(i % N), mm = (int)(i / N);
                    const cx<T> av = a[n], bv = bq[(size_t)P * mm], rv = r[i];
                    const double wr = (double)av.re * bv.re - (double)av.im * bv.im, wi = (double)av.re * bv.im + (double)av.im * bv.re;   // a b
                    re += wr * rv.re + wi * rv.im; im += wr * rv.im - wi * rv.re;                                                       // conj(a b) r
                }
                for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
                if (lane == 0) { s_red[warp][0] = re; s_red[warp][1] = im; }
                __syncthreads();
                if (tid == 0) { double cr = 0, ci = 0; for (int w = 0; w < KU_W; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; } s_cmag[c] = cr * cr + ci * ci; }
                __syncthreads();
            }
            // whole rows: u = R b_p' (N values), c_g = a_g' u for every g; entries inside a slightly wider band join the list with exact values
            double* u = reinterpret_cast<double*>(smem);                         // 2 N doubles (the least-squares scratch is not live yet)
            const double lim2 = s_lim2;
            for (int rr = 0; rr < nrow; ++rr) {
                const int pq = s_rows[rr];
                const cx<T>* bq = p.B + (long long)b * p.ld_B + pq;
                for (int n = tid; n < 2 * N; n += KU_T) u[n] = 0.0;
                __syncthreads();
                for (size_t i = tid; i < NM; i += KU_T) {
                    const int n = (int)(i % N), mm = (int)(i / N);
                    const cx<T> bv = bq[(size_t)P * mm], x = r[i];
                    atomicAdd(&u[2 * n], (double)x.re * bv.re + (double)x.im * bv.im);
                    atomicAdd(&u[2 * n + 1], (double)x.im * bv.re - (double)x.re * bv.im);
                }
                __syncthreads();
                for (int g = tid; g < G; g += KU_T) {
                    const cx<T>* a = p.A + (long long)b * p.ld_A + (size_t)N * g;
                    double re = 0.0, im = 0.0;
                    for (int n = 0; n < N; ++n) { const cx<T> av = a[n]; re += (double)av.re * u[2 * n] + (double)av.im * u[2 * n + 1]; im += (double)av.re * u[2 * n + 1] - (double)av.im * u[2 * n]; }
                    const double mag = re * re + im * im;
                    if (mag >= lim2) { const int k = atomicAdd(&s_nc, 1); if (k < MAXC) { s_cand[k] = g + G * pq; s_cmag[k] = mag; } }
                }
                __syncthreads();
            }
        }
        // the whole-row pass may have overflowed the list: nothing has been written yet, the fp32 pass decides this iteration
        if (p.screen && s_nc > MAXC) { if (tid == 0) p.flag[b] = 1; return; }
        if (tid == 0) {
            const int nc = s_nc < MAXC ? s_nc : MAXC;
            double gb = -1.0, gs = -1.0; int gi = nc > 0 ? s_cand[0] : 0;
            if (nc > 1 || nrow > 0) {
                gi = 0x7fffffff;
                for (int c = 0; c < nc; ++c) {
                    const double wb = s_cmag[c]; const int wi = s_cand[c];
                    if (wb > gb || (wb == gb && wi < gi)) { if (gb > gs) gs = gb; gb = wb; gi = wi; }
                    else if (wb > gs) gs = wb;
                }
                // "ambiguous" now means: two EXACT (fp64) values closer than fp64 rounding of an N M-term sum - a genuine tie, resolved like max() by the lowest index
                if (gs >= 0.0 && gb - gs <= 1e-12 * gb) st[1]++;
            }
            p.index_set[(size_t)b * m + t] = gi + 1;                              // 1-based (OMP.m:17)
            int dup = -1;
            for (int k = 0; k < nu; ++k) if (uniq_idx[k] == gi) dup = k;
            if (dup >= 0) { sel[t] = dup; mult[dup]++; }
            s_pick = gi; s_dup = dup;
        }
    }
    __syncthreads();
    if (s_dup >= 0) return;                                                   // no new direction (see omp.cu)
    // ---- least squares x = pinv(T) v (OMP.m:19) through the Cholesky factor of the Gram matrix T'T, all in fp64.  Atoms are
    //      rank-one, so <atom_j, atom_t> = (a_j' a_t)(b_j b_t') costs N + M terms and T itself is never stored. ----
    const int M = p.M, pick = s_pick, g = pick % G, pp = pick / G;
    constexpr int J = 32;
    cx<double>* gcol = reinterpret_cast<cx<double>*>(smem);                   // Gram column -> w = L^-1 g
    cx<double>* xtmp = gcol + (m + 1);
    double* dg = reinterpret_cast<double*>(xtmp + (m + 1));                   // diagonal of L
    cx<T>* a_s = reinterpret_cast<cx<T>*>(dg + ((m + 3) & ~1));             // keeps cx<double> 16-byte aligned for odd m
    cx<T>* b_s = a_s + N;
    cx<T>* As = b_s + M;
    cx<T>* Bs = As + J * N;
    cx<T>* Asel = p.Asel + (size_t)b * m * N;
    cx<T>* Bsel = p.Bsel + (size_t)b * m * M;
    cx<double>* K = p.K + (size_t)b * m * m;
    cx<double>* yv = p.yv + (size_t)b * m;
    const cx<T>* Ag = p.A + (long long)b * p.ld_A + (size_t)N * g;
    const cx<T>* Bg = p.B + (long long)b * p.ld_B + pp;
    const cx<T>* Y = p.Y + (long long)b * p.ld_Y;
    __shared__ cx<double> s_rhs;
    for (int i = tid; i < N; i += KU_T) { const cx<T> a = Ag[i]; a_s[i] = a; Asel[(size_t)nu * N + i] = a; }               // atom = A(:,g) B(p,:)  (OMP.m:18)
    for (int i = tid; i < M; i += KU_T) { const cx<T> v = Bg[(size_t)P * i]; b_s[i] = v; Bsel[(size_t)nu * M + i] = v; }
    __syncthreads();
    {   // rhs_t = atom' v
        double re = 0.0, im = 0.0;
        for (size_t i = tid; i < NM; i += KU_T) {
            const int n = (int)(i % N), mm = (int)(i / N);
            const cx<T> av = a_s[n], bv = b_s[mm], y = Y[i];
            const double wr = (double)av.re * bv.re - (double)av.im * bv.im, wi = (double)av.re * bv.im + (double)av.im * bv.re;
            re += wr * y.re + wi * y.im; im += wr * y.im - wi * y.re;
        }
        for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
        if (lane == 0) { s_red[warp][0] = re; s_red[warp][1] = im; }
    }
    for (int j = warp; j <= nu; j += KU_W) {                                     // Gram column g_j = <atom_j, atom_t>, j = nu is the new atom itself
        const cx<T>* aj = j == nu ? a_s : Asel + (size_t)j * N;
        const cx<T>* bj = j == nu ? b_s : Bsel + (size_t)j * M;
        double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0;
        for (int n = lane; n < N; n += 32) { const cx<T> x = aj[n], y = a_s[n]; ar += (double)x.re * y.re + (double)x.im * y.im; ai += (double)x.re * y.im - (double)x.im * y.re; }
        for (int mm = lane; mm < M; mm += 32) { const cx<T> x = bj[mm], y = b_s[mm]; br += (double)x.re * y.re + (double)x.im * y.im; bi += (double)x.re * y.im - (double)x.im * y.re; }
        for (int o = 16; o > 0; o >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, o); ai += __shfl_xor_sync(0xffffffffu, ai, o);
            br += __shfl_xor_sync(0xffffffffu, br, o); bi += __shfl_xor_sync(0xffffffffu, bi, o);
        }
        if (lane == 0) gcol[j] = mk<double>(ar * br - ai * bi, ar * bi + ai * br);
    }
    __syncthreads();
    if (tid == 0) { double cr = 0, ci = 0; for (int w = 0; w < KU_W; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; } s_rhs = mk<double>(cr, ci); }
    __syncthreads();
    // K = L^-1 (row-major, lower) is kept instead of L: the new row is -(w' K)/l_tt with w = K g, so every step below is a
    // parallel matrix-vector product (one memory round trip) instead of a latency chain of triangular-solve steps.
    __shared__ double s_ltt;
    __shared__ int s_dep;
    __shared__ cx<double> s_yt;
    for (int j = warp; j < nu; j += KU_W) {                                      // w = K g
        const cx<double>* kr = K + (size_t)j * m;
        cx<double> acc = mk<double>(0.0, 0.0);
        for (int i = lane; i <= j; i += 32) acc = acc + kr[i] * gcol[i];
        for (int o = 16; o > 0; o >>= 1) { acc.re += __shfl_xor_sync(0xffffffffu, acc.re, o); acc.im += __shfl_xor_sync(0xffffffffu, acc.im, o); }
        if (lane == 0) xtmp[j] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        double s2 = 0.0;
        for (int j = lane; j < nu; j += 32) s2 += xtmp[j].re * xtmp[j].re + xtmp[j].im * xtmp[j].im;
        for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const double gtt = gcol[nu].re, d2 = gtt - s2;
        const double dep_tol = sizeof(T) == 4 ? 1e-10 : 1e-24;               // a column numerically inside the span of the chosen ones carries no new direction
        if (lane == 0) { s_dep = !(d2 > dep_tol * gtt); s_ltt = s_dep ? 0.0 : sqrt(d2); }
    }
    __syncthreads();
    const double iltt = s_dep ? 0.0 : 1.0 / s_ltt;                            // dependent column: zero row, x_t = 0, nothing else changes
    cx<double>* rhsv = p.rhsv + (size_t)b * m;
    {
        cx<double> part = mk<double>(0.0, 0.0);
        for (int i = tid; i < nu; i += KU_T) {                                 // K(t, i) = -(1/l_tt) sum_{j >= i} conj(w_j) K(j, i)
            cx<double> acc = mk<double>(0.0, 0.0);
            for (int j = i; j < nu; ++j) acc = acc + conj(xtmp[j]) * K[(size_t)j * m + i];
            const cx<double> kt = mk<double>(-iltt * acc.re, -iltt * acc.im);
            K[(size_t)nu * m + i] = kt;
            part = part + kt * rhsv[i];
        }
        for (int o = 16; o > 0; o >>= 1) { part.re += __shfl_xor_sync(0xffffffffu, part.re, o); part.im += __shfl_xor_sync(0xffffffffu, part.im, o); }
        if (lane == 0) { s_red[warp][0] = part.re; s_red[warp][1] = part.im; }
    }
    __syncthreads();
    if (tid == 0) {
        double cr = 0, ci = 0; for (int w = 0; w < KU_W; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; }
        const cx<double> yt = mk<double>(cr + iltt * s_rhs.re, ci + iltt * s_rhs.im);                                   // y_t = K(t,:) rhs
        K[(size_t)nu * m + nu] = mk<double>(iltt, 0.0);
        rhsv[nu] = s_rhs; yv[nu] = yt; s_yt = yt;
    }
    __syncthreads();
    {
        cx<double>* xv = p.xv + (size_t)b * m;
        for (int i = tid; i <= nu; i += KU_T) {                                // x = K' y
            cx<double> acc = mk<double>(0.0, 0.0);
            for (int j = i; j <= nu; ++j) acc = acc + conj(K[(size_t)j * m + i]) * (j == nu ? s_yt : yv[j]);
            xtmp[i] = acc; xv[i] = acc;
        }
    }
    // ---- r = v - T x (OMP.m:20-21) as a rank-(nu+1) update of Y, J atoms per pass out of shared memory ----
    for (int j0 = 0; j0 <= nu; j0 += J) {
        const int jn = (nu + 1 - j0) < J ? (nu + 1 - j0) : J;
        __syncthreads();
        for (int e = tid; e < jn * N; e += KU_T) { const int j = e / N, n = e % N; As[e] = (j0 + j == nu) ? a_s[n] : Asel[(size_t)(j0 + j) * N + n]; }
        for (int e = tid; e < jn * M; e += KU_T) {
            const int j = e / M, mm = e % M;
            const cx<double> x = xtmp[j0 + j];
            const cx<T> v = (j0 + j == nu) ? b_s[mm] : Bsel[(size_t)(j0 + j) * M + mm];
            Bs[e] = mk<T>((T)(x.re * v.re - x.im * v.im), (T)(x.re * v.im + x.im * v.re));
        }
        __syncthreads();
        for (size_t i = tid; i < NM; i += KU_T) {
            const int n = (int)(i % N), mm = (int)(i / N);
            T ar = 0, ai = 0;
            for (int j = 0; j < jn; ++j) { const cx<T> a = As[j * N + n], bb = Bs[j * M + mm]; cmac<T>(ar, ai, a.re, a.im, bb.re, bb.im); }
            const cx<T> base = j0 == 0 ? Y[i] : r[i];
            r[i] = mk<T>(base.re - ar, base.im - ai);
        }
    }
    if (tid == 0) { sel[t] = nu; uniq_idx[nu] = pick; mult[nu] = 1; st[0] = nu + 1; }
}

template <typename T>
__global__ void __launch_bounds__(256) k_kron_finish(KronP<T> p) {
    const int b = blockIdx.x, tid = threadIdx.x, m = p.m;
    const size_t NM = (size_t)p.N * p.M;
    const cx<double>* z = p.xv + (size_t)b * m;
    int* st = p.state + (size_t)b * (2 + 3 * m);
    int* sel = st + 2; int* uniq_idx = sel + m; int* mult = uniq_idx + m;
    const int nu = st[0];
    if (tid == 0 && p.ambiguous) p.ambiguous[b] = st[1];
    if (p.x_sel) for (int k = tid; k < m; k += 256) { const int u = sel[k]; const T s = T(1) / (T)mult[u]; p.x_sel[(size_t)b * m + k] = mk<T>((T)z[u].re * s, (T)z[u].im * s); }
    if (p.r_out) { cx<T>* ro = p.r_out + (long long)b * p.ld_r; const cx<T>* r = p.res + (size_t)b * NM; for (size_t i = tid; i < NM; i += 256) ro[i] = r[i]; }
    if (p.x_hat) {
        cx<T>* xh = p.x_hat + (long long)b * p.ld_x;
        const size_t D = (size_t)p.G * p.P;
        for (size_t j = tid; j < D; j += 256) xh[j] = mk<T>(T(0), T(0));       // OMP.m:27
        __syncthreads();
        for (int k = tid; k < nu; k += 256) { const T s = T(1) / (T)mult[k]; xh[uniq_idx[k]] = mk<T>((T)z[k].re * s, (T)z[k].im * s); }   // OMP.m:29-31, pinv's even split over duplicates
    }
}

template <typename T>
static int run_omp_kron(Handle* h, int mem, int N, int M, int G, int P, int m, int batch,
                        const void* A_, long long ld_A, const void* B_, long long ld_B, const void* Y_, long long ld_Y,
                        void* x_, long long ld_x, int* idx_, void* xsel_, void* r_, long long ld_r, int* amb_, double margin_tol) {
    if (N <= 0 || M <= 0 || G <= 0 || P <= 0 || m <= 0 || batch <= 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if ((long long)G * P > 0x7fffffffLL) return fail(h, JSTSP_E_ARG, "G*P exceeds the int32 index range of index_set");
    if (!A_ || !B_ || !Y_ || !idx_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    const size_t esz = sizeof(cx<T>);
    const size_t NM = (size_t)N * M, NG = (size_t)N * G, PM = (size_t)P * M, GP = (size_t)G * P;
    if (ld_Y == 0) ld_Y = (long long)NM;
    if (ld_x == 0) ld_x = (long long)GP;
    if (ld_r == 0) ld_r = (long long)NM;
    const int Npad = ceil_div(N, KC_KT) * KC_KT;
    const size_t smem_corr = esz * ((size_t)Npad * KC_PC + KC_KT * KC_RB + KC_KT * KC_PC);
    int rc = set_smem(h, k_kron_corr<T>, smem_corr);
    if (rc) return rc;
    const size_t smem_upd = 2 * sizeof(cx<double>) * (size_t)(m + 1) + sizeof(double) * (size_t)((m + 3) & ~1) + esz * (size_t)(33 * (N + M));
    rc = set_smem(h, k_kron_update<T>, smem_upd);
    if (rc) return rc;
    const int nchunk = ceil_div(P, KC_PC);
    // tensor-core screen: fp32, the shapes of BASELINE config 2 (N = 64 rows; (g,c) groups of 128; K slices of 32 floats)
    const char* env_tc = getenv("JSTSP_OMP_TC");
    const bool use_tc = sizeof(T) == 4 && N == 64 && G % 64 == 0 && M % 16 == 0 && P >= 128 && !(env_tc && env_tc[0] == '0');
    const int ntile = ceil_div(P, kt::PT);
    if (use_tc) {
        cudaError_t e = cudaFuncSetAttribute(k_kron_corr_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kt::SMEM);
        if (e != cudaSuccess) return fail(h, JSTSP_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    }
    int chunk = batch;
    if (h->max_chunk > 0 && chunk > h->max_chunk) chunk = h->max_chunk;
    const bool sharedA = ld_A == 0, sharedB = ld_B == 0;
    auto layout = [&](Arena& a, int nb, KronP<T>& q) {
        q.res = a.take<cx<T>>(NM * nb);
        q.K = a.take<cx<double>>((size_t)nb * m * m);
        q.rhsv = a.take<cx<double>>((size_t)nb * m);
        q.yv = a.take<cx<double>>((size_t)nb * m);
        q.xv = a.take<cx<double>>((size_t)nb * m);
        q.Asel = a.take<cx<T>>((size_t)nb * m * N);
        q.Bsel = a.take<cx<T>>((size_t)nb * m * M);
        q.state = a.take<int>((size_t)nb * (2 + 3 * m));
        q.cand_val = a.take<double>((size_t)nb * nchunk * 2);
        q.cand_idx = a.take<int>((size_t)nb * nchunk);
        q.cand_idx2 = a.take<int>((size_t)nb * nchunk);
        q.abmax = a.take<float>(nb);
        q.AH = a.take<cx<T>>(sharedA ? NG : NG * nb);
        if (use_tc) {
            q.Bt = a.take<float>((sharedB ? 1 : (size_t)nb) * 2 * PM);
            q.Ach = a.take<float>((sharedA ? 1 : (size_t)nb) * 4 * NG);
            q.Rp = a.take<float>((size_t)nb * 4 * NM);
            q.row_v1 = a.take<float>((size_t)nb * P);
            q.row_g1 = a.take<int>((size_t)nb * P);
            q.row_v2 = a.take<float>((size_t)nb * P);
            q.flag = a.take<int>(nb);
        }
        if (host) {
            q.A = a.take<cx<T>>(sharedA ? NG : NG * nb);
            q.B = a.take<cx<T>>(sharedB ? PM : PM * nb);
            q.Y = a.take<cx<T>>(NM * nb);
            q.index_set = a.take<int>((size_t)m * nb);
            if (x_) q.x_hat = a.take<cx<T>>(GP * nb);
            if (xsel_) q.x_sel = a.take<cx<T>>((size_t)m * nb);
            if (r_) q.r_out = a.take<cx<T>>(NM * nb);
            if (amb_) q.ambiguous = a.take<int>(nb);
        }
    };
    size_t freeb = 0, totalb = 0;
    JSTSP_CUDA(h, cudaMemGetInfo(&freeb, &totalb));
    const size_t budget = (size_t)((freeb + h->ws_bytes) * 0.7);
    for (;;) {
        Arena probe(nullptr, 0); KronP<T> q{}; layout(probe, chunk, q);
        if (probe.off <= budget || chunk == 1) { rc = ensure_workspace(h, probe.off); if (rc) return rc; break; }
        chunk = (chunk + 1) / 2;
    }
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = (batch - b0) < chunk ? (batch - b0) : chunk;
        Arena ar(h->ws, h->ws_bytes);
        KronP<T> q{};
        q.N = N; q.M = M; q.G = G; q.P = P; q.m = m; q.nchunk = nchunk; q.margin_tol = margin_tol;
        q.ntile = ntile; q.band = 0.015f; q.dbg = h->dbg;
        layout(ar, nb, q);
        q.ld_AH = sharedA ? 0 : (long long)NG;
        if (host) {
            auto up = [&](const cx<T>* dst, const void* src, size_t per, long long ld, bool shared) -> cudaError_t {
                if (shared) return cudaMemcpyAsync(const_cast<cx<T>*>(dst), src, per * esz, cudaMemcpyHostToDevice, st);
                return cudaMemcpy2DAsync(const_cast<cx<T>*>(dst), per * esz, (const char*)src + (size_t)b0 * ld * esz, (size_t)ld * esz, per * esz, nb, cudaMemcpyHostToDevice, st);
            };
            JSTSP_CUDA(h, up(q.A, A_, NG, ld_A, sharedA));
            JSTSP_CUDA(h, up(q.B, B_, PM, ld_B, sharedB));
            JSTSP_CUDA(h, up(q.Y, Y_, NM, ld_Y, false));
            q.ld_A = sharedA ? 0 : (long long)NG; q.ld_B = sharedB ? 0 : (long long)PM; q.ld_Y = (long long)NM;
            q.ld_x = (long long)GP; q.ld_r = (long long)NM;
        } else {
            q.A = (const cx<T>*)A_ + (long long)b0 * ld_A; q.ld_A = ld_A;
            q.B = (const cx<T>*)B_ + (long long)b0 * ld_B; q.ld_B = ld_B;
            q.Y = (const cx<T>*)Y_ + (long long)b0 * ld_Y; q.ld_Y = ld_Y;
            q.x_hat = x_ ? (cx<T>*)x_ + (long long)b0 * ld_x : nullptr; q.ld_x = ld_x;
            q.index_set = idx_ + (size_t)b0 * m;
            q.x_sel = xsel_ ? (cx<T>*)xsel_ + (size_t)b0 * m : nullptr;
            q.r_out = r_ ? (cx<T>*)r_ + (long long)b0 * ld_r : nullptr; q.ld_r = ld_r;
            q.ambiguous = amb_ ? amb_ + b0 : nullptr;
        }
        JSTSP_LAUNCH(h, PK_SETUP, (k_kron_setup<T><<<nb, 256, 0, st>>>(q, sharedA ? 1 : nb, sharedB ? 1 : nb)));
        if (use_tc) JSTSP_LAUNCH(h, PK_SETUP, (k_kron_pack_r<T><<<dim3(ceil_div(M, 32), nb), 256, 0, st>>>(q)));
        CUtensorMap mapBt, mapR, mapA;
        if (use_tc) {
            const bool ok = um::make_map_k128(q.Bt, 2 * M, P, sharedB ? 1 : nb, 2 * (long long)PM, kt::PT, &mapBt) &&
                            um::make_map_k128(q.Rp, 2 * M, 2 * N, nb, 4 * (long long)NM, 128, &mapR) &&
                            um::make_map_k128(q.Ach, 2 * N, 2 * G, sharedA ? 1 : nb, 4 * (long long)NG, 128, &mapA);
            if (!ok) return fail(h, JSTSP_E_CUDA, "cuTensorMapEncodeTiled failed for the Kronecker OMP operands");
        }
        for (int t = 0; t < m; ++t) {
            q.t = t;
            if (use_tc) {
                if constexpr (sizeof(T) == 4) {
                    q.screen = 1;
                    const int items = ntile * nb;
                    JSTSP_LAUNCH(h, PK_OMP_CORR_TC, (k_kron_corr_tc<<<items < h->sm_count ? items : h->sm_count, kt::THREADS, kt::SMEM, st>>>(q, mapBt, mapR, mapA, sharedB ? 1 : 0, sharedA ? 1 : 0, nb)));
                    JSTSP_LAUNCH(h, PK_OMP, (k_kron_update<T><<<nb, KU_T, smem_upd, st>>>(q)));
                    q.screen = 0;                                             // trials the screen left undecided (rare) redo the iteration in fp32
                }
            }
            {
                const int items = nchunk * nb, cap = 4 * h->sm_count;
                JSTSP_LAUNCH(h, PK_OMP_CORR, (k_kron_corr<T><<<(use_tc && items > cap) ? cap : items, 256, smem_corr, st>>>(q, nb)));
            }
            JSTSP_LAUNCH(h, use_tc ? PK_OTHER : PK_OMP, (k_kron_update<T><<<nb, KU_T, smem_upd, st>>>(q)));
            if (use_tc && t + 1 < m) JSTSP_LAUNCH(h, PK_SETUP, (k_kron_pack_r<T><<<dim3(ceil_div(M, 32), nb), 256, 0, st>>>(q)));
        }
        JSTSP_LAUNCH(h, PK_OMP, (k_kron_finish<T><<<nb, 256, 0, st>>>(q)));
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            JSTSP_CUDA(h, cudaMemcpyAsync(idx_ + (size_t)b0 * m, q.index_set, sizeof(int) * (size_t)m * nb, cudaMemcpyDeviceToHost, st));
            if (x_) JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)x_ + (size_t)b0 * ld_x * esz, (size_t)ld_x * esz, q.x_hat, GP * esz, GP * esz, nb, cudaMemcpyDeviceToHost, st));
            if (xsel_) JSTSP_CUDA(h, cudaMemcpyAsync((char*)xsel_ + (size_t)b0 * m * esz, q.x_sel, esz * (size_t)m * nb, cudaMemcpyDeviceToHost, st));
            if (r_) JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)r_ + (size_t)b0 * ld_r * esz, (size_t)ld_r * esz, q.r_out, NM * esz, NM * esz, nb, cudaMemcpyDeviceToHost, st));
            if (amb_) JSTSP_CUDA(h, cudaMemcpyAsync(amb_ + b0, q.ambiguous, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
            JSTSP_CUDA(h, cudaStreamSynchronize(st));
        }
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" int jstsp_omp_kron(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int m, int batch,
                              const void* A, long long ld_A, const void* B, long long ld_B, const void* Y, long long ld_Y,
                              void* x_hat, long long ld_x, int* index_set, void* x_sel, void* residual, long long ld_r,
                              int* ambiguous, double margin_tol) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_omp_kron<float>(h, mem, N, M, G, P, m, batch, A, ld_A, B, ld_B, Y, ld_Y, x_hat, ld_x, index_set, x_sel, residual, ld_r, ambiguous, margin_tol);
    if (dtype == JSTSP_F64) return run_omp_kron<double>(h, mem, N, M, G, P, m, batch, A, ld_A, B, ld_B, Y, ld_Y, x_hat, ld_x, index_set, x_sel, residual, ld_r, ambiguous, margin_tol);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
