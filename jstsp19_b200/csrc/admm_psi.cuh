// admm_psi.cuh - structured-dictionary ("Psi-domain") tcgen05 path of the proposed ADMM iteration.
//
// The reference's driver builds the estimator's dictionary as B((l-1)Gt+1:lGt,:) = Dt' * Psi_bar(:,:,l)
// (plot_errorVSsnr.m:133-136) where Psi_bar(k,:,l) is row l of toeplitz(s_k) of the 4-QAM pilot sequence s_k
// (plot_errorVSsnr.m:63-67, proposed_hbf.m:15-18).  Two facts about that operand are used here:
//   * Psi_bar(k,j,l) = e_k(j-l): every delay tap is the SAME Nt x (M+L-1) array e shifted by l columns
//     (e_k(t) = s_k(t) for t >= 0 and conj(s_k(-t)) for t < 0 - MATLAB's Hermitian toeplitz), and
//   * its entries are +-scale +- j scale: exactly representable in bf16 after one common scaling.
// So instead of streaming the dense 2 MiB fp32 dictionary B twice per iteration (admm_fast.cuh / admm_tc.cuh) the fused
// kernel keeps the (128+L-1)-column slice of e that a 128-column chunk of X needs - 34 KiB in bf16 - in shared memory and
// runs BOTH big products of the iteration out of it:
//   pass 1   Xs(:,chunk) = sum_l Q_l e(:, chunk - l),          Q_l = (A S_l) Dt'          (proposed_algorithm.m:58)
//   pass 2   T1'_l      = K(:,chunk) e(:, chunk - l)^H  ;      K B^H = [T1'_l Dt]_l        (proposed_algorithm.m:47)
// The products with the pilots are EXACT in the dictionary operand; the small operands (Q, K) are split into three bf16
// terms (24 mantissa bits) that ride on the N side of one kind::f16 MMA (N = 2 * 16 rows * 3 = 96), so the result has fp32
// accuracy.  The tile is stored [kc / 8][column][kc % 8] (kc = 2 * antenna + (re|im)) with SWIZZLE_NONE, which makes it a
// K-major operand for pass 1 and an MN-major operand for pass 2 at the same time, and makes a delay tap a start-address
// offset of 16 bytes per column (layouts and the shift verified on the B200 by tools/umma_probe4.cu: 56.7 cycles per
// M=128 N=96 K=16 MMA).
//
// The line search of the gradient step (proposed_algorithm.m:47-50) is restated so that it needs no B B^H at all:
//   res = K2'(k - K2 v)  ->  Res = A' (K - XV) B',  XV = A V B carried across iterations
//   res' R res = |K2 res|^2 = |A Res B|_F^2 = |G|_F^2,  G = (A Res) B      (a third pilot product, pass-1 shaped)
//   v += alpha res       ->  V += alpha Res,  XV += alpha G
// Per iteration: k_fused_psi (Xs, element-wise update, T1'), k_psi_res (Res, operand of G), k_psi_g (G, |G|^2),
// k_psi_step (alpha, V, S, XV, operand of Xs).  The Dt rotations (16 x 64 x 64 per tap) run in the two small kernels.
// State tiles (16 KiB per array and 128 columns) travel by TMA tensor copies with SWIZZLE_128B, so HBM access is fully
// coalesced and each thread reads its 64-byte column segment from shared memory without bank conflicts.
//
// Inputs that do not have this structure (Gaussian pilots, non-Toeplitz Psi_bar, Nt != 64, fp64) are served by the dense
// kernels after materialising B = (I (x) Dt') Psi on the device (k_build_b) - a different GPU kernel, never a CPU path.
#pragma once
#include <cuda_bf16.h>
#include "admm_tc.cuh"

namespace jstsp {
namespace psi {

using tc::MC; using tc::WORKERS; using tc::NWW; using tc::THREADS; using tc::ZP;

constexpr int N = 16;                    // rows of X (RF-chain domain)
constexpr int KC = 128;                  // 2 * Nt: (antenna, re|im) rows of the pilot tile = MMA M of pass 2
constexpr int NT = KC / 2;
constexpr int NKG = KC / 8;              // 8-element groups along kc
constexpr int ROWS = MC + 8;             // columns of e held per chunk (>= 128 + L - 1)
constexpr int RS = ROWS * 16;            // bytes between kc groups inside the tile
constexpr int TILE = NKG * RS;           // 34816
constexpr int MAXL = 8;
constexpr int NS = 6 * N;                // rows of the small operands: (split, n, c)
constexpr int NRG = NS / 8;              // 12 row groups
constexpr int QKS = NRG * 2 * 128;       // bytes of one K = 16 step of the pass-1 operand image
constexpr int QTAP = (KC / 16) * QKS;    // bytes of one tap of the pass-1 operand image (24576)
constexpr int QSLOT = QTAP / 2;          // the image is streamed in half taps
constexpr int KOP_LBO = NRG * 128 + 16;  // pass-2 operand: stride between 8-column k groups (+16: bank spread for the worker stores)
constexpr int KOP = (MC / 8) * KOP_LBO;  // 24832
constexpr int OPREG = 25 * 1024;         // two pass-1 half-tap slots; then the XV tile; then the pass-2 operand
constexpr int SLOT = N * MC * 8;         // one state tile: 128 columns x 16 rows complex fp32 = 16 KiB, SWIZZLE_128B
constexpr int ZREG = 2 * N * ZP * 4;
constexpr int WREG = N * N * 8;
constexpr int BARS = 256;
constexpr int TMEM_COLS = 256;           // two 96-column accumulators, shared by pass 1 (even/odd taps) and pass 2
constexpr int OFF_OP = TILE, OFF_S0 = OFF_OP + OPREG, OFF_S1 = OFF_S0 + SLOT, OFF_Z = OFF_S1 + SLOT, OFF_W = OFF_Z + ZREG, OFF_BAR = OFF_W + WREG;
constexpr size_t SMEM = (size_t)OFF_BAR + BARS;
static_assert(KOP <= OPREG && 2 * QSLOT <= OPREG && SLOT <= OPREG, "operand region too small");
static_assert(OFF_OP % 1024 == 0 && OFF_S0 % 1024 == 0 && OFF_S1 % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
// G kernel: tile | two half-tap slots | output tile; 128 TMEM columns -> three CTAs per SM
constexpr int G_NSLOT = 2;
constexpr int G_TMEM_COLS = 128;
constexpr int G_OFF_OP = TILE, G_OFF_OUT = G_OFF_OP + G_NSLOT * QSLOT, G_OFF_BAR = G_OFF_OUT + SLOT;
constexpr size_t G_SMEM = (size_t)G_OFF_BAR + BARS;
static_assert(G_OFF_OUT % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");

__host__ __device__ constexpr uint32_t instr_desc_bf16(int Mm, int Nn, int a_mn, int b_mn = 0) {   // kind::f16, bf16 x bf16 -> fp32; a_mn / b_mn: operand is MN-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(Mm >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned short bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf16_val(unsigned short b) { return __uint_as_float((uint32_t)b << 16); }
// x = hi + mid + lo (+ <= 2^-24 |x|), each term a bf16
__device__ __forceinline__ void split3(float x, unsigned short (&s)[3]) {
    s[0] = bf16_bits(x); x -= bf16_val(s[0]);
    s[1] = bf16_bits(x); x -= bf16_val(s[1]);
    s[2] = bf16_bits(x);
}
// this thread's 8 rows (64 bytes) of column m in a SWIZZLE_128B state tile: 16-byte chunk j of row m sits at chunk j ^ (m % 8)
__device__ __forceinline__ void tile_read8(const unsigned char* slot, int m, int half, cx<float> (&v)[8]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const float4 a = *reinterpret_cast<const float4*>(slot + m * 128 + (((4 * half + u) ^ (m & 7)) << 4));
        v[2 * u] = mk<float>(a.x, a.y); v[2 * u + 1] = mk<float>(a.z, a.w);
    }
}
__device__ __forceinline__ void tile_write8(unsigned char* slot, int m, int half, const cx<float> (&v)[8]) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
        *reinterpret_cast<float4*>(slot + m * 128 + (((4 * half + u) ^ (m & 7)) << 4)) = make_float4(v[2 * u].re, v[2 * u].im, v[2 * u + 1].re, v[2 * u + 1].im);
}

struct In {                      // the structured dictionary of one pass (device pointers)
    const cx<float>* Psi; long long ld_Psi;     // Psi_bar: Nt x M x L per trial (proposed_hbf.m:7,17)
    const cx<float>* Dt;  long long ld_Dt;      // Nt x Gt
    int Nt, Gt, L;
    unsigned short* E;           // [nE][NKG][Mext][8] bf16 pilot image, Mext = M + 8
    float* scale;                // [b] common magnitude of the pilot components
    unsigned short* omask;       // [b][M] bit n = Omega(n, m) != 0
    unsigned char* QopS;         // [b][L][QTAP] pass-1 operand image of Q_S = scale (A S)(I (x) Dt')
    unsigned char* QopG;         // [b][L][QTAP] same for (A Res)
    cx<float>* T1p;              // [b][N][L*Nt] (K - XV) Psi^H, accumulated over the chunks of an iteration
    cx<float>* XV;               // [b][M][N] A V B
    cx<float>* Gm;               // [b][M][N] G = (A Res) B of the current iteration
    double* rr;                  // [b][L]   |Res_l|^2
    double* gg;                  // [b][nmc] |G(:,chunk)|^2
    int* bad;                    // [0] structure violations seen by the packing kernels, [1] entries of Dt that differ from the unitary 64-point DFT
    int t1_red;                  // 1: T1' is accumulated over chunks by red.global.add into [b][N][L*Nt]; 0: per-chunk partials [b][nmc][N][L*Nt]
    float snap_tol;              // > 0: the pilots were recovered from a dense dictionary (k_recover_psi) and carry fp32 rounding noise: components within
                                 // snap_tol * scale of +-scale are snapped, the Toeplitz check compares with the same tolerance
    int dft;                     // Dt is the unitary DFT grid of wideband_mmwave_channel.m:9-10 with Gt = Nt = 64: rotations run as radix-4 FFTs
};
struct Maps { CUtensorMap E, X, V1, V2, XV, G, SY; };

// ---- pilots -> bf16 image + structure check ------------------------------------------------------------------------------
// grid (ceil(Mext / 16), nE), block 256: thread = (antenna k, 4 columns of e)
__global__ void __launch_bounds__(256) k_pack_psi(In in, int M) {
    const int b = blockIdx.y, Nt = in.Nt, L = in.L, Mext = M + 8;
    const cx<float>* Ps = in.Psi + (long long)b * in.ld_Psi;
    unsigned short* E = in.E + (size_t)b * NKG * Mext * 8;
    float sc = fabsf(Ps[0].re);
    if (sc == 0.f) sc = fabsf(Ps[0].im);
    if (sc == 0.f) sc = 1.f;
    const float tol = in.snap_tol * sc;
    if (blockIdx.x == 0 && threadIdx.x == 0) in.scale[b] = sc;
    const int k = threadIdx.x % NT;
    int nbad = 0;
    for (int q = threadIdx.x / NT; q < 16; q += 256 / NT) {
        const int te = blockIdx.x * 16 + q;
        if (te >= Mext) break;
        const int t = te - (L - 1);                               // column index of e: -(L-1) .. M-1, then zero padding
        cx<float> e = mk<float>(0.f, 0.f);
        if (t >= 0 && t < M) e = Ps[k + (size_t)Nt * t];                                  // tap 0, column t
        else if (t < 0) e = Ps[k + (size_t)Nt * M * (size_t)(-t)];                        // tap -t, column 0
        const unsigned short br = bf16_bits(e.re / sc), bi = bf16_bits(e.im / sc);
        if (tol > 0.f) { if (t < M && (fabsf(bf16_val(br) * sc - e.re) > tol || fabsf(bf16_val(bi) * sc - e.im) > tol || fabsf(bf16_val(br)) != 1.f || fabsf(bf16_val(bi)) != 1.f)) ++nbad; }
        else if (bf16_val(br) * sc != e.re || bf16_val(bi) * sc != e.im) ++nbad;          // not exact in bf16 after the common scaling
        for (int l = 1; l < L; ++l) {                                                     // Toeplitz: tap l reads e(j - l) at column j = t + l
            const int j = t + l;
            if (j >= 0 && j < M && (t >= 0 || l != -t)) {
                const cx<float> v = Ps[k + (size_t)Nt * j + (size_t)Nt * M * l];
                if (tol > 0.f ? (fabsf(v.re - e.re) > 2.f * tol || fabsf(v.im - e.im) > 2.f * tol) : (v.re != e.re || v.im != e.im)) ++nbad;
            }
        }
        // kc = 2k (re), 2k+1 (im): both in group k / 4
        *reinterpret_cast<uint32_t*>(E + ((size_t)(k / 4) * Mext + te) * 8 + 2 * (k % 4)) = (uint32_t)br | ((uint32_t)bi << 16);
    }
    if (nbad) atomicAdd(in.bad, nbad);
}
// Omega -> 16-bit column masks; anything but 0/1 weights is left to the dense kernels.  grid (ceil(M/256), nb)
__global__ void __launch_bounds__(256) k_pack_omega(In in, const float* __restrict__ omega, long long ld_omega, int M) {
    const int b = blockIdx.y, m = blockIdx.x * 256 + threadIdx.x;
    if (m >= M) return;
    const float* o = omega + (long long)b * ld_omega + (size_t)m * N;
    unsigned bits = 0; int nbad = 0;
#pragma unroll
    for (int u = 0; u < N / 4; ++u) {
        const float4 v = *reinterpret_cast<const float4*>(o + 4 * u);
        const float w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { if (w[j] == 1.f) bits |= 1u << (4 * u + j); else if (w[j] != 0.f) ++nbad; }
    }
    in.omask[(size_t)b * M + m] = (unsigned short)bits;
    if (nbad) atomicAdd(in.bad, nbad);
}
// shared pilots: every trial uses the scale of the one image
__global__ void k_spread_scale(float* scale, int nb) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > 0 && t < nb) scale[t] = scale[0];
}

// ---- B = (I (x) Dt') Psi, the dense dictionary (plot_errorVSsnr.m:133-136) for the unstructured route -------------------------
// grid (ceil(M / 64), L, nB), block 256: 64 columns x Gt rows per CTA
template <typename T>
__global__ void __launch_bounds__(256) k_build_b(const cx<T>* __restrict__ Psi, long long ld_Psi, const cx<T>* __restrict__ Dt, long long ld_Dt, cx<T>* __restrict__ B,
                                                 long long ld_B, int Nt, int Gt, int L, int M) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    cx<T>* sD = reinterpret_cast<cx<T>*>(sm_raw);                 // Dt, Nt x Gt
    cx<T>* sP = sD + (size_t)Nt * Gt;                             // Psi_l(:, 64 columns)
    const int b = blockIdx.z, l = blockIdx.y, j0 = blockIdx.x * 64, P = L * Gt;
    const cx<T>* Ps = Psi + (long long)b * ld_Psi + (size_t)Nt * M * l;
    const cx<T>* D = Dt + (long long)b * ld_Dt;
    for (int t = threadIdx.x; t < Nt * Gt; t += 256) sD[t] = D[t];
    for (int t = threadIdx.x; t < Nt * 64; t += 256) { const int k = t % Nt, j = j0 + t / Nt; sP[t] = j < M ? Ps[k + (size_t)Nt * j] : mk<T>(T(0), T(0)); }
    __syncthreads();
    cx<T>* out = B + (long long)b * ld_B;
    for (int t = threadIdx.x; t < Gt * 64; t += 256) {
        const int g = t % Gt, jj = t / Gt;
        if (j0 + jj >= M) continue;
        T re = 0, im = 0;
        for (int k = 0; k < Nt; ++k) { const cx<T> d = sD[k + (size_t)Nt * g], v = sP[k + (size_t)Nt * jj]; cmac<T>(re, im, d.re, -d.im, v.re, v.im); }
        out[(size_t)(l * Gt + g) + (size_t)P * (j0 + jj)] = mk<T>(re, im);
    }
}

// ---- Psi_bar recovered from a dense dictionary: B((l-1)Gt+1 : l Gt, :) = Dt' Psi_bar(:,:,l) with the unitary 64-point DFT grid means
//      Psi_bar(:,:,l) = Dt B_l.  The reference function's own argument list (proposed_algorithm.m:1) carries only B; when B has the
//      structure its drivers give it (plot_errorVSsnr.m:133-136) this puts the call on the Psi-domain tensor-core path.  Whether it has is
//      decided by k_pack_psi on the result (4-QAM components and Toeplitz taps within rounding noise), never assumed.
// grid (ceil(M / 64), L, nB), block 256
__global__ void __launch_bounds__(256) k_recover_psi(const cx<float>* __restrict__ B, long long ld_B, cx<float>* __restrict__ Psi, long long ld_Psi, int L, int M) {
    __shared__ cx<float> sB[NT * 65];                             // B_l(g, 64 columns), padded
    __shared__ cx<float> stw[NT];
    const int b = blockIdx.z, l = blockIdx.y, j0 = blockIdx.x * 64, P = L * NT;
    if (threadIdx.x < NT) { float sn, cs; sincospif(-2.0f * (float)threadIdx.x / NT, &sn, &cs); stw[threadIdx.x] = mk<float>(0.125f * cs, 0.125f * sn); }
    const cx<float>* Bb = B + (long long)b * ld_B;
    for (int t = threadIdx.x; t < NT * 64; t += 256) { const int g = t % NT, jj = t / NT; sB[g + 65 * jj] = j0 + jj < M ? Bb[(size_t)(l * NT + g) + (size_t)P * (j0 + jj)] : mk<float>(0.f, 0.f); }
    __syncthreads();
    cx<float>* out = Psi + (long long)b * ld_Psi + (size_t)NT * M * l;
    for (int t = threadIdx.x; t < NT * 64; t += 256) {
        const int k = t % NT, jj = t / NT;
        if (j0 + jj >= M) continue;
        float re = 0.f, im = 0.f;
        for (int g = 0; g < NT; ++g) { const cx<float> d = stw[(k * g) % NT], v = sB[g + 65 * jj]; cmac<float>(re, im, d.re, d.im, v.re, v.im); }
        out[k + (size_t)NT * (j0 + jj)] = mk<float>(re, im);
    }
}
// Dt(k, g) = exp(-2 pi j k g / 64) / 8 (wideband_mmwave_channel.m:9-10 with Gt = Mt = 64)
__global__ void __launch_bounds__(256) k_make_dft64(cx<float>* Dt) {
    for (int t = threadIdx.x; t < NT * NT; t += 256) {
        const int k = t % NT, g = t / NT;
        double sn, cs; sincospi(-2.0 * (double)((k * g) % NT) / NT, &sn, &cs);
        Dt[t] = mk<float>((float)(0.125 * cs), (float)(0.125 * sn));
    }
}

// ---- small per-(tap, trial) kernels ------------------------------------------------------------------------------------------
// X (N x NT, element (n,k) at X[n + N k]) -> three-term bf16 operand image of pass 1 for tap l.
// image per trial: [tap][ks 0..7][kg 0..1][row group 0..11][8 rows][8 kc], row = 32 split + 2 n + c, kc = 16 ks + 8 kg + i
//   rows (n,re) = [Qr, -Qi], rows (n,im) = [Qi, Qr] along kc = (k,re),(k,im)       -> D[m][(n,c)] = (Q e)^T
// rot = 0: rows [hi | mid | lo].  rot = 2 (odd taps of the G operand): rows [mid | lo | hi], so that with the accumulator shifted by 32
// columns mid and lo of every tap share columns while hi alternates between two column blocks (k_psi_g).
__device__ __forceinline__ void put_q(unsigned char* img, int n, int k, float qr, float qi, int rot) {
    auto put = [&](int row, int kc, float v) {
        unsigned short s[3]; split3(v, s);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int r = 32 * ((u + rot) % 3) + row;
            *reinterpret_cast<unsigned short*>(img + (size_t)(kc / 16) * QKS + ((kc % 16) / 8) * (NRG * 128) + (r / 8) * 128 + (r % 8) * 16 + (kc % 8) * 2) = s[u];
        }
    };
    put(2 * n, 2 * k, qr); put(2 * n, 2 * k + 1, -qi);
    put(2 * n + 1, 2 * k, qi); put(2 * n + 1, 2 * k + 1, qr);
}
constexpr int DLD = NT + 2;      // padded leading dimension of Dt in shared memory: FWD reads (stride 2 * DLD) hit distinct banks, BWD float4 reads stay aligned
constexpr int LDU = N + 2;       // padded leading dimension of the N x 64 work matrices: column pairs land in distinct banks, float4 row pairs stay aligned
struct __align__(16) SmallSmem {
    cx<float> D[DLD * NT];       // Dt(:, 0..Gt-1), column g at D + DLD g; later the operand image / Res staging
    cx<float> A[N * N];          // A, N x G (zero-padded to N x N), element (n,g) at [n + N g]
    cx<float> AH[N * N];         // A', element (g,n) at [g + N n]
    cx<float> U[LDU * NT];       // ping: element (r,c) at [r + LDU c]
    cx<float> V[LDU * NT];       // pong
    cx<float> tw[NT];            // exp(-2 pi j t / 64)
    double red[8];
};
// The two Dt rotations, register-tiled 2 rows x 2 outputs per thread (256 threads = 8 row pairs x 32 output pairs):
//   FWD:  out(n,g) = scale sum_k in(n,k) Dt(k,g)            (T1'_l Dt)
//   BWD:  out(n,k) = scale sum_g in(n,g) conj(Dt(k,g))      ((A S_l) Dt')
// in / out are N x 64 with leading dimension LDU; D is Dt, NT x Gt with leading dimension DLD.
template <bool FWD>
__device__ __forceinline__ void rotate(const cx<float>* __restrict__ D, const cx<float>* __restrict__ in, cx<float>* __restrict__ out, int Gt, float sc) {
    const int n0 = 2 * (threadIdx.x % 8), o0 = 2 * (threadIdx.x / 8);
    const int nout = FWD ? Gt : NT, nin = FWD ? NT : Gt;
    if (o0 >= nout) return;
    const bool two = o0 + 1 < nout;
    float r00 = 0.f, i00 = 0.f, r01 = 0.f, i01 = 0.f, r10 = 0.f, i10 = 0.f, r11 = 0.f, i11 = 0.f;     // [row][output]
#pragma unroll 8
    for (int i = 0; i < nin; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(in + n0 + LDU * i);                           // rows n0, n0+1 of column i
        float d0r, d0i, d1r, d1i;
        if (FWD) {
            const cx<float> d0 = D[i + DLD * o0], d1 = D[i + DLD * (two ? o0 + 1 : o0)];
            d0r = d0.re; d0i = d0.im; d1r = d1.re; d1i = d1.im;
        } else {
            const float4 d = *reinterpret_cast<const float4*>(D + o0 + DLD * i);                         // Dt(o0, i), Dt(o0+1, i)  (DLD is even)
            d0r = d.x; d0i = -d.y; d1r = d.z; d1i = -d.w;
        }
        cmac<float>(r00, i00, a.x, a.y, d0r, d0i); cmac<float>(r10, i10, a.z, a.w, d0r, d0i);
        cmac<float>(r01, i01, a.x, a.y, d1r, d1i); cmac<float>(r11, i11, a.z, a.w, d1r, d1i);
    }
    *reinterpret_cast<float4*>(out + n0 + LDU * o0) = make_float4(sc * r00, sc * i00, sc * r10, sc * i10);
    if (two) *reinterpret_cast<float4*>(out + n0 + LDU * (o0 + 1)) = make_float4(sc * r01, sc * i01, sc * r11, sc * i11);
}
// out (N x cols) = Mx * in, Mx an N x N matrix with element (r,i) at [r + N i] (sm.A, or sm.AH for A'; unused rows / columns are zero),
// in / out with leading dimension LDU.  2 x 2 register tile, two inner indices per step, 16-byte shared-memory accesses only.
__device__ __forceinline__ void apply_a(const cx<float>* __restrict__ Mx, const cx<float>* __restrict__ in, cx<float>* __restrict__ out, int cols) {
    const int r0 = 2 * (threadIdx.x % 8), c0 = 2 * (threadIdx.x / 8);
    if (c0 >= cols) return;
    const bool two = c0 + 1 < cols;
    const int c1 = two ? c0 + 1 : c0;
    float r00 = 0.f, i00 = 0.f, r01 = 0.f, i01 = 0.f, r10 = 0.f, i10 = 0.f, r11 = 0.f, i11 = 0.f;     // [row][column]
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        const float4 x0 = *reinterpret_cast<const float4*>(in + i + LDU * c0), x1 = *reinterpret_cast<const float4*>(in + i + LDU * c1);   // rows i, i+1
        const float4 a = *reinterpret_cast<const float4*>(Mx + r0 + N * i), bq = *reinterpret_cast<const float4*>(Mx + r0 + N * (i + 1));  // rows r0, r0+1
        cmac<float>(r00, i00, a.x, a.y, x0.x, x0.y); cmac<float>(r10, i10, a.z, a.w, x0.x, x0.y);
        cmac<float>(r01, i01, a.x, a.y, x1.x, x1.y); cmac<float>(r11, i11, a.z, a.w, x1.x, x1.y);
        cmac<float>(r00, i00, bq.x, bq.y, x0.z, x0.w); cmac<float>(r10, i10, bq.z, bq.w, x0.z, x0.w);
        cmac<float>(r01, i01, bq.x, bq.y, x1.z, x1.w); cmac<float>(r11, i11, bq.z, bq.w, x1.z, x1.w);
    }
    *reinterpret_cast<float4*>(out + r0 + LDU * c0) = make_float4(r00, i00, r10, i10);
    if (two) *reinterpret_cast<float4*>(out + r0 + LDU * c1) = make_float4(r01, i01, r11, i11);
}

// Dt == exp(-2 pi j k g / 64) / 8 ?  (Dr/Dt of wideband_mmwave_channel.m:9-10 with Gt = Mt = 64).  grid nDt, block 256
__global__ void __launch_bounds__(256) k_check_dt(In in) {
    const cx<float>* D = in.Dt + (long long)blockIdx.x * in.ld_Dt;
    int nbad = 0;
    for (int t = threadIdx.x; t < NT * NT; t += 256) {
        const int k = t % NT, g = t / NT;
        double sn, cs;
        sincospi(-2.0 * (double)((k * g) % NT) / NT, &sn, &cs);
        const cx<float> d = D[t];
        if (fabs((double)d.re - 0.125 * cs) > 1e-7 || fabs((double)d.im - 0.125 * sn) > 1e-7) ++nbad;
    }
    if (nbad) atomicAdd(in.bad + 1, nbad);
}
// 64-point FFT of the N rows of `in` (element (n,k) at [n + N k]) into `out`, scaled; INV: exp(+...) kernel.  256 threads, radix-4 DIF:
// one butterfly per thread and stage, digit-reversed positions resolved by the last stage's stores.  `in` is used as scratch.
struct CtaSync { __device__ __forceinline__ void operator()() const { __syncthreads(); } };
template <bool INV, class Sync = CtaSync>
__device__ __forceinline__ void fft64(cx<float>* __restrict__ in, cx<float>* __restrict__ out, const cx<float>* __restrict__ tw, float scale, Sync sync = Sync()) {
    const int n = threadIdx.x % N, j = threadIdx.x / N;          // j = 0..15
    auto twid = [&](int t) { cx<float> w = tw[t]; if (INV) w.im = -w.im; return w; };
    auto bfly = [&](cx<float> x0, cx<float> x1, cx<float> x2, cx<float> x3, cx<float> (&a)[4]) {
        const float s02r = x0.re + x2.re, s02i = x0.im + x2.im, d02r = x0.re - x2.re, d02i = x0.im - x2.im;
        const float s13r = x1.re + x3.re, s13i = x1.im + x3.im, d13r = x1.re - x3.re, d13i = x1.im - x3.im;
        a[0] = mk<float>(s02r + s13r, s02i + s13i);
        a[2] = mk<float>(s02r - s13r, s02i - s13i);
        // forward: a1 = d02 - j d13, a3 = d02 + j d13 ; inverse: the other way round      (-j (r + j i) = i - j r)
        const cx<float> m = mk<float>(d02r + d13i, d02i - d13r), q = mk<float>(d02r - d13i, d02i + d13r);
        a[1] = INV ? q : m; a[3] = INV ? m : q;
    };
    cx<float> a[4];
    // stage 0: length 64, i = j
    bfly(in[n + LDU * j], in[n + LDU * (j + 16)], in[n + LDU * (j + 32)], in[n + LDU * (j + 48)], a);
    sync();
#pragma unroll
    for (int q = 0; q < 4; ++q) in[n + LDU * (j + 16 * q)] = q ? a[q] * twid(q * j) : a[q];
    sync();
    {   // stage 1: length 16, group j / 4, i = j % 4
        const int base = 16 * (j / 4), i = j % 4;
        cx<float> x0 = in[n + LDU * (base + i)], x1 = in[n + LDU * (base + i + 4)], x2 = in[n + LDU * (base + i + 8)], x3 = in[n + LDU * (base + i + 12)];
        bfly(x0, x1, x2, x3, a);
        sync();
#pragma unroll
        for (int q = 0; q < 4; ++q) in[n + LDU * (base + i + 4 * q)] = q ? a[q] * twid(4 * q * i) : a[q];
    }
    sync();
    {   // stage 2: length 4, group j; position 4 j + q holds output index rev4(4 j + q) = 16 q + 4 (j % 4) + j / 4
        bfly(in[n + LDU * (4 * j)], in[n + LDU * (4 * j + 1)], in[n + LDU * (4 * j + 2)], in[n + LDU * (4 * j + 3)], a);
#pragma unroll
        for (int q = 0; q < 4; ++q) out[n + LDU * (16 * q + 4 * (j % 4) + j / 4)] = mk<float>(scale * a[q].re, scale * a[q].im);
    }
}
__device__ __forceinline__ void load_small(SmallSmem& sm, const In& in, const AdmmP<float>& p, int b) {
    const cx<float>* D = in.Dt + (long long)b * in.ld_Dt;
    const cx<float>* A = p.A + (long long)b * p.ld_A;
    if (in.dft) { if (threadIdx.x < NT) { float sn, cs; sincospif(-2.0f * (float)threadIdx.x / NT, &sn, &cs); sm.tw[threadIdx.x] = mk<float>(cs, sn); } }
    else for (int t = threadIdx.x; t < NT * in.Gt; t += 256) sm.D[(t % NT) + DLD * (t / NT)] = D[t];
    for (int t = threadIdx.x; t < N * N; t += 256) {
        const cx<float> a = t < N * p.G ? A[t] : mk<float>(0.f, 0.f);
        sm.A[t] = a; sm.AH[(t / N) + N * (t % N)] = mk<float>(a.re, -a.im);
    }
}

// sm.U (N x NT) -> operand image of one tap: assembled in shared memory (over Dt, which is no longer needed), written out coalesced
__device__ __forceinline__ void write_image(SmallSmem& sm, unsigned char* __restrict__ gimg, int rot) {
    static_assert(sizeof(sm.D) >= QTAP, "image does not fit over Dt");
    unsigned char* img = reinterpret_cast<unsigned char*>(sm.D);
    const int n = threadIdx.x % N;
    for (int k = threadIdx.x / N; k < NT; k += 256 / N) { const cx<float> q = sm.U[n + LDU * k]; put_q(img, n, k, q.re, q.im, rot); }
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(img);
    uint4* dst = reinterpret_cast<uint4*>(gimg);
    for (int t = threadIdx.x; t < QTAP / 16; t += 256) dst[t] = src[t];
}

// Res_l = A' (scale T1'_l Dt) ; |Res_l|^2 ; operand image of G_l = (A Res_l) Dt'           (proposed_algorithm.m:47)
// grid (L, nb), block 256
__global__ void __launch_bounds__(256, 4) k_psi_res(AdmmP<float> p, In in) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    SmallSmem& sm = *reinterpret_cast<SmallSmem*>(sm_raw);
    const int b = blockIdx.y, l = blockIdx.x, Gt = in.Gt, Pp = in.L * NT, G = p.G;
    const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;
    JSTSP_STAMP(p, 6, cta_id, 0);
    load_small(sm, in, p, b);
    if (in.t1_red) {
        cx<float>* src = in.T1p + (size_t)b * N * Pp;
        for (int t = threadIdx.x; t < N * NT / 2; t += 256) {    // two adjacent k per thread: 16-byte accesses; the accumulator is handed back zeroed
            const int n = t / (NT / 2), k = 2 * (t % (NT / 2));
            float4* q = reinterpret_cast<float4*>(src + (size_t)n * Pp + l * NT + k);
            const float4 v = *q;
            *q = make_float4(0.f, 0.f, 0.f, 0.f);
            sm.U[n + LDU * k] = mk<float>(v.x, v.y); sm.U[n + LDU * (k + 1)] = mk<float>(v.z, v.w);
        }
    } else {
        // sum of the chunk partials: the 8 loads of a position are issued before the first add
        const cx<float>* src = in.T1p + (size_t)b * p.nmc * N * Pp + l * NT;
        for (int t = threadIdx.x; t < N * NT / 2; t += 256) {    // two adjacent k per thread: 16-byte coalesced loads
            const int n = t / (NT / 2), k = 2 * (t % (NT / 2));
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int cb = 0; cb < p.nmc; cb += 8) {
                float4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    v[c] = cb + c < p.nmc ? *reinterpret_cast<const float4*>(src + (size_t)(cb + c) * N * Pp + (size_t)n * Pp + k) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < 8; ++c) { acc.x += v[c].x; acc.y += v[c].y; acc.z += v[c].z; acc.w += v[c].w; }
            }
            sm.U[n + LDU * k] = mk<float>(acc.x, acc.y); sm.U[n + LDU * (k + 1)] = mk<float>(acc.z, acc.w);
        }
    }
    __syncthreads();
    const float sc = in.scale[b];
    cx<float>* Res = p.Res + (size_t)b * G * p.P + (size_t)G * Gt * l;
    double rr = 0.0;
    JSTSP_STAMP(p, 6, cta_id, 1);
    cx<float>* R = sm.D;                                         // Res_l staging (leading dimension LDU); Dt is dead in the FFT route, unused here otherwise
    if (in.dft) {
        // Dt is unitary: Res_l = (A' T1'_l) Dt (one FFT) and (A Res_l) Dt' = A A' T1'_l (no rotation at all)
        apply_a(sm.AH, sm.U, sm.V, NT);                          // A' T1'_l                        G x NT (rows >= G zero)
        __syncthreads();
        apply_a(sm.A, sm.V, sm.U, NT);                           // A A' T1'_l                      N x NT
        __syncthreads();
        for (int t = threadIdx.x; t < N * NT; t += 256) { const int i = (t % N) + LDU * (t / N); cx<float> v = sm.U[i]; sm.U[i] = mk<float>(sc * sc * v.re, sc * sc * v.im); }
        JSTSP_STAMP(p, 6, cta_id, 2);
        fft64<false>(sm.V, R, sm.tw, 0.125f * sc);               // (write_image reads sm.U after its own barrier)
        __syncthreads();
        JSTSP_STAMP(p, 6, cta_id, 3);
    } else {
        rotate<true>(sm.D, sm.U, sm.V, Gt, sc);                  // T1_l = scale T1'_l Dt           N x Gt
        __syncthreads();
        apply_a(sm.AH, sm.V, sm.U, Gt);                          // Res_l = A' T1_l                 G x Gt (rows >= G zero)
        __syncthreads();
        R = sm.U;
    }
    for (int t = threadIdx.x; t < N * Gt; t += 256) {
        const int r = t % N, c = t / N;
        const cx<float> v = R[r + LDU * c];
        if (r < G) { Res[r + (size_t)G * c] = v; rr += (double)v.re * v.re + (double)v.im * v.im; }
    }
    for (int o = 16; o > 0; o >>= 1) rr += __shfl_down_sync(0xffffffffu, rr, o);
    if (threadIdx.x % 32 == 0) sm.red[threadIdx.x / 32] = rr;
    if (!in.dft) apply_a(sm.A, sm.U, sm.V, Gt);                  // A Res_l                         N x Gt
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0; for (int w = 0; w < 8; ++w) s += sm.red[w]; in.rr[(size_t)b * in.L + l] = s; }
    if (!in.dft) { rotate<false>(sm.D, sm.V, sm.U, Gt, sc); __syncthreads(); }     // scale (A Res_l) Dt'   N x NT
    JSTSP_STAMP(p, 6, cta_id, 4);
    write_image(sm, in.QopG + ((size_t)b * in.L + l) * QTAP, (l & 1) ? 2 : 0);
    JSTSP_STAMP(p, 6, cta_id, 5);
}

// alpha = res'res / (res' R res) (proposed_algorithm.m:48) from the per-tap / per-chunk partial sums; called by the first warp, same
// summation order in every CTA of a trial (V and XV must be stepped by the same number)
__device__ __forceinline__ float psi_alpha(const In& in, int b, int L, int nmc) {
    double a = 0, c = 0;
    for (int i = threadIdx.x; i < L; i += 32) a += in.rr[(size_t)b * L + i];
    for (int i = threadIdx.x; i < nmc; i += 32) c += in.gg[(size_t)b * nmc + i];
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); c += __shfl_down_sync(0xffffffffu, c, o); }
    return (float)(a / c);
}
// alpha = |Res|^2 / |G|^2 ; V += alpha Res ; S = soft(V) [masked] ; XV += alpha G ; operand image of Xs = (A S) B
// (proposed_algorithm.m:48-58, proposed_algorithm_angles.m:68).  grid (L, nb), block 256
__global__ void __launch_bounds__(256, 4) k_psi_step(AdmmP<float> p, In in, int make_q, int do_xv) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    SmallSmem& sm = *reinterpret_cast<SmallSmem*>(sm_raw);
    __shared__ float s_alpha;
    const int b = blockIdx.y, l = blockIdx.x, Gt = in.Gt, G = p.G, L = in.L;
    if (make_q) load_small(sm, in, p, b);
    if (threadIdx.x < 32) { const float al = psi_alpha(in, b, L, p.nmc); if (threadIdx.x == 0) s_alpha = al; }   // alpha = res'res / (res' R res)  (.m:48)
    __syncthreads();
    const float alpha = s_alpha;
    const float thr = (float)(p.tauS[b] / p.rho[b]);
    const size_t off = (size_t)b * G * p.P + (size_t)G * Gt * l;
    cx<float>* __restrict__ V = p.V + off;
    cx<float>* __restrict__ S = p.S + off;
    const cx<float>* __restrict__ Res = p.Res + off;
    const unsigned char* mask = p.angles ? p.smask + off : nullptr;
    for (int t = threadIdx.x; t < G * Gt; t += 256) {
        const cx<float> r = Res[t];
        cx<float> v = V[t];
        v = mk<float>(v.re + alpha * r.re, v.im + alpha * r.im);
        cx<float> s = mk<float>(soft1<float>(v.re, thr), soft1<float>(v.im, thr));
        if (mask && !mask[t]) s = mk<float>(0.f, 0.f);
        V[t] = v; S[t] = s;
        if (make_q) sm.U[(t % G) + LDU * (t / G)] = s;
    }
    if (do_xv) {   // XV += alpha G on this tap's share of the columns, four independent 16-byte accesses in flight per thread
        const int M = p.M, per = (M + L - 1) / L, m0 = l * per, m1 = (m0 + per) < M ? (m0 + per) : M;
        float4* xv = reinterpret_cast<float4*>(in.XV + (size_t)b * N * M);
        const float4* g = reinterpret_cast<const float4*>(in.Gm + (size_t)b * N * M);
        const int t1 = m1 * (N / 2);
        for (int t = m0 * (N / 2) + threadIdx.x; t < t1; t += 4 * 256) {
            float4 x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) if (t + u * 256 < t1) { x[u] = xv[t + u * 256]; y[u] = g[t + u * 256]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t + u * 256 < t1) {
                    x[u].x += alpha * y[u].x; x[u].y += alpha * y[u].y; x[u].z += alpha * y[u].z; x[u].w += alpha * y[u].w;
                    xv[t + u * 256] = x[u];
                }
        }
    }
    if (!make_q) return;
    if (G < N) for (int t = threadIdx.x; t < N * Gt; t += 256) if (t % N >= G) sm.U[(t % N) + LDU * (t / N)] = mk<float>(0.f, 0.f);
    __syncthreads();
    apply_a(sm.A, sm.U, sm.V, Gt);                     // A S_l                           N x Gt   (.m:58, left factor)
    __syncthreads();
    if (in.dft) fft64<true>(sm.V, sm.U, sm.tw, 0.125f * in.scale[b]);
    else rotate<false>(sm.D, sm.V, sm.U, Gt, in.scale[b]);       // scale (A S_l) Dt'               N x NT
    __syncthreads();
    write_image(sm, in.QopS + ((size_t)b * L + l) * QTAP, 0);
}

// ---- shared pieces of the two tensor-core kernels ---------------------------------------------------------------------------
// producer: stream the operand image of `taps` taps in half-tap slots
template <int NSLOT>
__device__ __forceinline__ void stream_q(const unsigned char* img, unsigned char* slots, uint64_t* q_full, uint64_t* q_empty, int taps) {
    for (int i = 0; i < 2 * taps; ++i) {
        const int slot = i % NSLOT;
        if (i >= NSLOT) mbar_wait(&q_empty[slot], ((i / NSLOT) - 1) & 1);
        mbar_expect_tx(&q_full[slot], QSLOT);
        tma_bulk_g2s(slots + slot * QSLOT, img + (size_t)i * QSLOT, QSLOT, &q_full[slot]);
    }
}
// MMA thread: D[l & 1][m][(split,n,c)] += e(chunk - l)^T Q'_l^T for l < taps
template <int NSLOT>
__device__ __forceinline__ void issue_pass1(uint32_t tile_a, uint32_t slots_a, const uint32_t (&D)[2], uint64_t* q_full, uint64_t* q_empty, int taps, int L) {
    constexpr uint32_t id1 = instr_desc_bf16(128, NS, 0);
    for (int i = 0; i < 2 * taps; ++i) {
        const int slot = i % NSLOT, l = i >> 1, hf = i & 1;
        mbar_wait(&q_full[slot], (i / NSLOT) & 1);
        tc::tc_fence_after();
        const uint32_t a0 = tile_a + (uint32_t)(L - 1 - l) * 16, b0 = slots_a + slot * QSLOT;
#pragma unroll
        for (int j = 0; j < KC / 32; ++j) {                  // K-major A: kc groups RS apart, 8-column groups 128 B apart
            const int ks = hf * (KC / 32) + j;
            umma_bf16(D[l & 1], tc::smem_desc(a0 + ks * 2 * RS, RS, 128, 0), tc::smem_desc(b0 + j * QKS, NRG * 128, 128, 0), id1, (l >= 2 || ks) ? 1u : 0u);
        }
        tc::umma_commit(&q_empty[slot]);
    }
}
// worker: this thread's 8 rows of column m of the pass-1 result; smallest terms first: lo, mid, hi of both accumulators
__device__ __forceinline__ void read_pass1(const uint32_t (&D)[2], uint32_t lane_base, int n0, int taps, float (&xr)[8], float (&xi)[8]) {
    float hi[16], mid[16], lo[16];
    tc::tmem_ld16x3(D[0] + lane_base + 2 * n0, D[0] + lane_base + 32 + 2 * n0, D[0] + lane_base + 64 + 2 * n0, hi, mid, lo);
#pragma unroll
    for (int r = 0; r < 8; ++r) { xr[r] = lo[2 * r]; xi[r] = lo[2 * r + 1]; }
    if (taps > 1) {
        float hi1[16], mid1[16], lo1[16];
        tc::tmem_ld16x3(D[1] + lane_base + 2 * n0, D[1] + lane_base + 32 + 2 * n0, D[1] + lane_base + 64 + 2 * n0, hi1, mid1, lo1);
#pragma unroll
        for (int j = 0; j < 16; ++j) { lo[j] = lo1[j]; mid[j] += mid1[j]; hi[j] += hi1[j]; }
#pragma unroll
        for (int r = 0; r < 8; ++r) { xr[r] += lo[2 * r]; xi[r] += lo[2 * r + 1]; }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) { xr[r] += mid[2 * r]; xi[r] += mid[2 * r + 1]; }
#pragma unroll
    for (int r = 0; r < 8; ++r) { xr[r] += hi[2 * r]; xi[r] += hi[2 * r + 1]; }
}

// ---- G = (A Res) B on the chunk, |G|^2 partial ------------------------------------------------------------------------------
// grid (M / 128, nb), 320 threads: warps 0-7 workers, warp 8 TMA producer, warp 9 MMA issuer; 74 KiB and 128 TMEM columns per CTA,
// three CTAs per SM.  Accumulator columns: [0,32) hi of even taps, [32,64) mid, [64,96) lo, [96,128) hi of odd taps - even taps
// accumulate into [0,96) with operand rows [hi|mid|lo], odd taps into [32,128) with rows [mid|lo|hi] (k_psi_res writes them so),
// which keeps the large hi chains short (the tensor core truncates when it adds) in half the columns.
__global__ void __launch_bounds__(THREADS, 3) k_psi_g(AdmmP<float> p, const __grid_constant__ Maps maps, In in) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tile = smem;
    unsigned char* slots = smem + G_OFF_OP;
    unsigned char* outt = smem + G_OFF_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_OFF_BAR);
    uint64_t *tile_full = bars, *q_full = bars + 1, *q_empty = bars + 1 + G_NSLOT, *d1_full = bars + 1 + 2 * G_NSLOT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * G_NSLOT);
    double* red = reinterpret_cast<double*>(bars + 4 + 2 * G_NSLOT);
    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int b = blockIdx.y, chunk = blockIdx.x, c0 = chunk * MC, L = in.L;
    if (tid == 0) {
        mbar_init(tile_full, 1);
        for (int s = 0; s < G_NSLOT; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
        mbar_init(d1_full, 1);
        mbar_fence_init();
    }
    if (tid < 256) reinterpret_cast<uint4*>(outt)[tid] = make_uint4(0u, 0u, 0u, 0u);      // 4 KiB of zeros: the operand of the accumulator-clearing MMA
    tc::fence_async_smem();
    if (warp == NWW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)G_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = *tmem_slot;
    if (warp == NWW) {
        if (lane == 0) {
            mbar_expect_tx(tile_full, TILE);
            tc::tma_4d(tile, &maps.E, 0, c0, 0, in.ld_Psi ? b : 0, tile_full);
            stream_q<G_NSLOT>(in.QopG + (size_t)b * L * QTAP, slots, q_full, q_empty, L);
        }
        __syncwarp();
    } else if (warp == NWW + 1) {
        if (lane == 0) {
            constexpr uint32_t id1 = instr_desc_bf16(128, NS, 0), idz = instr_desc_bf16(128, 128, 0);
            const uint32_t tile_a = smem_u32(tile), slots_a = smem_u32(slots);
            mbar_wait(tile_full, 0);
            tc::tc_fence_after();
            // clear all 128 columns: (any 128 x 16 tile) x (128 x 16 zeros)
            umma_bf16(tm, tc::smem_desc(tile_a, RS, 128, 0), tc::smem_desc(smem_u32(outt), 2048, 128, 0), idz, 0u);
            for (int i = 0; i < 2 * L; ++i) {
                const int slot = i % G_NSLOT, l = i >> 1, hf = i & 1;
                mbar_wait(&q_full[slot], (i / G_NSLOT) & 1);
                tc::tc_fence_after();
                const uint32_t a0 = tile_a + (uint32_t)(L - 1 - l) * 16, b0 = slots_a + slot * QSLOT;
#pragma unroll
                for (int j = 0; j < KC / 32; ++j) {
                    const int ks = hf * (KC / 32) + j;
                    umma_bf16(tm + 32 * (l & 1), tc::smem_desc(a0 + ks * 2 * RS, RS, 128, 0), tc::smem_desc(b0 + j * QKS, NRG * 128, 128, 0), id1, 1u);
                }
                tc::umma_commit(&q_empty[slot]);
            }
            tc::umma_commit(d1_full);
        }
        __syncwarp();
    } else {
        const int quad = warp % 4, half = warp / 4, m = quad * 32 + lane, n0 = half * 8;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        mbar_wait(d1_full, 0);
        tc::tc_fence_after();
        float gr[8] = {}, gi[8] = {}, a[16];
        // smallest terms first: lo [64,96), mid [32,64), then the two hi blocks [0,32) and [96,128)
#pragma unroll
        for (int blk = 0; blk < 4; ++blk) {
            const int col = blk == 0 ? 64 : blk == 1 ? 32 : blk == 2 ? 0 : 96;
            tc::tmem_ld16(tm + lane_base + col + 2 * n0, a);
#pragma unroll
            for (int r = 0; r < 8; ++r) { gr[r] += a[2 * r]; gi[r] += a[2 * r + 1]; }
        }
        tc::tc_fence_before();
        cx<float> g[8];
        double ss = 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) { g[r] = mk<float>(gr[r], gi[r]); ss += (double)gr[r] * gr[r] + (double)gi[r] * gi[r]; }
        tile_write8(outt, m, half, g);                           // (the zero block has been consumed: d1_full covers the clearing MMA)
        tc::fence_async_smem();
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
        if (lane == 0) red[warp] = ss;
        tc::worker_sync();
        if (tid == 0) {
            tma_store_3d(&maps.G, outt, 0, c0, b);
            bulk_commit();
            double s = 0; for (int w = 0; w < NWW; ++w) s += red[w];
            in.gg[(size_t)b * p.nmc + chunk] = s;
            bulk_wait_read();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == NWW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)G_TMEM_COLS));
}

// ---- the fused iteration kernel ------------------------------------------------------------------------------------------
// grid (M / 128, nb), 320 threads: warps 0-7 workers, warp 8 TMA producer, warp 9 MMA issuer; two CTAs per SM.
__global__ void __launch_bounds__(THREADS, 2) k_fused_psi(AdmmP<float> p, const __grid_constant__ Maps maps, In in) {
    constexpr int NH = N / 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* tile = smem;
    unsigned char* opnd = smem + OFF_OP;
    unsigned char* S0 = smem + OFF_S0;
    unsigned char* S1 = smem + OFF_S1;
    float* Zre = reinterpret_cast<float*>(smem + OFF_Z);
    float* Zim = Zre + N * ZP;
    cx<float>* Wsm = reinterpret_cast<cx<float>*>(smem + OFF_W);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t *tile_full = bars, *q_full = bars + 1, *q_empty = bars + 3, *d1_full = bars + 5, *kop_ready = bars + 6, *d2_full = bars + 7, *d2_empty = bars + 9,
             *ld_full = bars + 11;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int b = blockIdx.y, chunk = blockIdx.x, c0 = chunk * MC;
    const int M = p.M, L = in.L, Pp = L * NT;
    const int S1t = p.iter > 0 ? L : 0;                      // iteration 0: Xs = C = V2 = 0, nothing to multiply
    const int sy_b = p.ld_subY ? b : 0;

    if (tid == 0) {
        mbar_init(tile_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&d2_full[s], 1); mbar_init(&d2_empty[s], NWW); }
        for (int s = 0; s < 3; ++s) mbar_init(&ld_full[s], 1);
        mbar_init(d1_full, 1); mbar_init(kop_ready, NWW);
        mbar_fence_init();
    }
    if (warp == NWW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t D[2] = {tm, tm + 128};                    // even / odd taps: short accumulation chains (the tensor core truncates when it adds)

    if (warp == NWW) {
        // ===== TMA producer =====
        if (lane == 0) {
            // the pilot tile first: pass 1 (and with it every later phase) waits for it, the workers have slack until X / V1 land
            mbar_expect_tx(tile_full, TILE);
            tc::tma_4d(tile, &maps.E, 0, c0, 0, in.ld_Psi ? b : 0, tile_full);
            mbar_expect_tx(&ld_full[0], 2 * SLOT);
            tc::tma_3d(S0, &maps.X, 0, c0, b, &ld_full[0]);
            tc::tma_3d(S1, &maps.V1, 0, c0, b, &ld_full[0]);
            if (S1t == 0) {                                  // no pass 1: the operand region is free for the XV tile right away
                mbar_expect_tx(&ld_full[2], SLOT);
                tc::tma_3d(opnd, &maps.XV, 0, c0, b, &ld_full[2]);
            }
            stream_q<2>(in.QopS + (size_t)b * L * QTAP, opnd, q_full, q_empty, S1t);
        }
        __syncwarp();
    } else if (warp == NWW + 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t id2 = instr_desc_bf16(128, NS, 1);
            const uint32_t tile_a = smem_u32(tile), op_a = smem_u32(opnd);
            mbar_wait(tile_full, 0);
            issue_pass1<2>(tile_a, op_a, D, q_full, q_empty, S1t, L);
            if (S1t > 0) tc::umma_commit(d1_full);
            mbar_wait(kop_ready, 0);                         // the workers have drained D and written the K operand
            tc::tc_fence_after();
            for (int l = 0; l < L; ++l) {                    // pass 2: D[l & 1][kc][(split,n,c)] = e(chunk - l) K'^T
                if (l >= 2) { mbar_wait(&d2_empty[l & 1], ((l >> 1) - 1) & 1); tc::tc_fence_after(); }
                const uint32_t a0 = tile_a + (uint32_t)(L - 1 - l) * 16;
#pragma unroll
                for (int ks = 0; ks < MC / 16; ++ks)         // MN-major A: LBO = 8-column (K) group stride, SBO = kc (MN) group stride
                    umma_bf16(D[l & 1], tc::smem_desc(a0 + ks * 256, 128, RS, 0), tc::smem_desc(op_a + ks * 2 * KOP_LBO, KOP_LBO, 128, 0), id2, ks ? 1u : 0u);
                tc::umma_commit(&d2_full[l & 1]);
            }
        }
        __syncwarp();
    } else {
        // ===== workers =====
        const int quad = warp % 4, half = warp / 4;
        const int m = quad * 32 + lane;                      // column of the chunk (pass 1) / kc row of the tile (pass 2)
        const int n0 = half * NH;
        const float rho = (float)p.rho[b];
        const float irho = 1.0f / rho, kap = rho / (rho + 1.0f);
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;
        JSTSP_STAMP(p, 3, cta_id, 0);
        Wsm[tid] = p.W[(size_t)b * N * N + tid];             // WORKERS == N * N
        const unsigned ombits = (unsigned)in.omask[(size_t)b * M + c0 + m] >> n0;

        // ---- round 1: X, V1 -> Z = X - V1/rho (SVT input, .m:35) ----
        cx<float> xo[NH], v1[NH];
        mbar_wait(&ld_full[0], 0);
        tile_read8(S0, m, half, xo); tile_read8(S1, m, half, v1);
#pragma unroll
        for (int r = 0; r < NH; ++r) { Zre[(n0 + r) * ZP + m] = xo[r].re - irho * v1[r].re; Zim[(n0 + r) * ZP + m] = xo[r].im - irho * v1[r].im; }
        tc::fence_async_smem();                              // generic-proxy reads of S0 / S1 before the async-proxy refill (cross-proxy WAR)
        tc::worker_sync();                                   // S0 / S1 consumed, Z and W complete
        if (tid == 0) {
            mbar_expect_tx(&ld_full[1], 2 * SLOT);
            tc::tma_3d(S0, &maps.V2, 0, c0, b, &ld_full[1]);
            tc::tma_3d(S1, &maps.SY, 0, c0, sy_b, &ld_full[1]);
        }
        // Y = W Z  (W = U diag(max(0,1-tau/sigma)) U^H from k_svt_weights); u = V1 + rho Y
        {
            float y_r[NH], y_i[NH];
#pragma unroll
            for (int r = 0; r < NH; ++r) { y_r[r] = 0.f; y_i[r] = 0.f; }
#pragma unroll 4
            for (int k = 0; k < N; ++k) {
                const float zr = Zre[k * ZP + m], zi = Zim[k * ZP + m];
                cx<float> w[NH];
#pragma unroll
                for (int hf = 0; hf < NH / 4; ++hf) {
                    cx<float> t[4];
                    ld4c<float>(Wsm + N * k + n0 + 4 * hf, t);
#pragma unroll
                    for (int u = 0; u < 4; ++u) w[4 * hf + u] = t[u];
                }
#pragma unroll
                for (int r = 0; r < NH; ++r) cmac<float>(y_r[r], y_i[r], w[r].re, w[r].im, zr, zi);
            }
            if ((p.iter == p.imax - 1) && p.Yout != nullptr) {
#pragma unroll
                for (int hf = 0; hf < NH / 4; ++hf) {
                    cx<float> t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) t[u] = mk<float>(y_r[4 * hf + u], y_i[4 * hf + u]);
                    st4c<float>(p.Yout + (long long)b * p.ld_Y + (size_t)(c0 + m) * N + n0 + 4 * hf, t);
                }
            }
#pragma unroll
            for (int r = 0; r < NH; ++r) { v1[r].re += rho * y_r[r]; v1[r].im += rho * y_i[r]; }        // u
        }
        // ---- pass 1 result: w = X - Xs ----
        float xs_r[NH], xs_i[NH];
        if (S1t > 0) {
            mbar_wait(d1_full, 0);
            tc::tc_fence_after();
            read_pass1(D, lane_base, n0, S1t, xs_r, xs_i);
            tc::tc_fence_before();
            if (tid == 0) {                                  // pass 1 is complete: its operand slots now receive the XV tile
                mbar_expect_tx(&ld_full[2], SLOT);
                tc::tma_3d(opnd, &maps.XV, 0, c0, b, &ld_full[2]);
            }
        } else {
#pragma unroll
            for (int r = 0; r < NH; ++r) { xs_r[r] = 0.f; xs_i[r] = 0.f; }
        }
        JSTSP_STAMP(p, 3, cta_id, 1);
        // ---- round 2: V2, subY -> C, V2, X, V1, K ----
        cx<float> v2[NH], kt[NH];
        {
            cx<float> sy[NH];
            mbar_wait(&ld_full[1], 0);
            tile_read8(S0, m, half, v2); tile_read8(S1, m, half, sy);
#pragma unroll
            for (int r = 0; r < NH; ++r) {
                // C = rho/(rho+1) (X - Xs - V2/rho) ; V2 += rho (C - X + Xs)      (.m:61,65 of the previous iteration; all zero at i = 1)
                const float wr = xo[r].re - xs_r[r], wi = xo[r].im - xs_i[r];
                const float cr = kap * (wr - irho * v2[r].re), ci = kap * (wi - irho * v2[r].im);
                v2[r].re += rho * (cr - wr); v2[r].im += rho * (ci - wi);
                const float d = 1.0f / (((ombits >> r) & 1u ? 1.0f : 0.0f) + 2.0f * rho);                                     // iK1 (.m:20)
                const float xr = (v1[r].re + sy[r].re + v2[r].re + rho * cr + rho * xs_r[r]) * d;                             // .m:38-40
                const float xi = (v1[r].im + sy[r].im + v2[r].im + rho * ci + rho * xs_i[r]) * d;
                v1[r].re -= rho * xr; v1[r].im -= rho * xi;                                                                   // .m:64: V1 + rho (Y - X)
                kt[r] = mk<float>(xr - irho * v2[r].re - cr, xi - irho * v2[r].im - ci);                                      // .m:43
                xo[r] = mk<float>(xr, xi);
            }
        }
        // X and V1 go back through the tiles they came from (thread-private segments); the next SVT input replaces Z later
        tile_write8(S0, m, half, xo); tile_write8(S1, m, half, v1);
#pragma unroll
        for (int r = 0; r < NH; ++r) { xo[r].re -= irho * v1[r].re; xo[r].im -= irho * v1[r].im; }      // zn = X - V1/rho
        // ---- round 3: XV = A V B -> K - XV, the operand of pass 2 (res = K2'(k - K2 v), .m:47) ----
        mbar_wait(&ld_full[2], 0);
        {
            cx<float> xv[NH];
            tile_read8(opnd, m, half, xv);
#pragma unroll
            for (int r = 0; r < NH; ++r) { kt[r].re -= xv[r].re; kt[r].im -= xv[r].im; }
        }
        tc::worker_sync();                                   // every thread has read its XV segment and finished W Z
        {   // pass-2 small operand: rows (split, n, c), k = m, K-major SWIZZLE_NONE with a padded k-group stride
            unsigned char* kb = opnd + (size_t)(m / 8) * KOP_LBO + (m % 8) * 2;
#pragma unroll
            for (int r = 0; r < NH; ++r) {
                const float v[2] = {kt[r].re, kt[r].im};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    unsigned short s[3]; split3(v[c], s);
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int row = 32 * u + 2 * (n0 + r) + c;
                        *reinterpret_cast<unsigned short*>(kb + (row / 8) * 128 + (row % 8) * 16) = s[u];
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < NH; ++r) { Zre[(n0 + r) * ZP + m] = xo[r].re; Zim[(n0 + r) * ZP + m] = xo[r].im; }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(kop_ready);
        tc::worker_sync();                                   // X / V1 tiles complete and fenced
        if (tid == 0) {
            tma_store_3d(&maps.X, S0, 0, c0, b);
            tma_store_3d(&maps.V1, S1, 0, c0, b);
            bulk_commit();
        }
        JSTSP_STAMP(p, 3, cta_id, 2);

        // ---- pass 2 epilogues: T1'_l partial, row-major [N][L * Nt] complex ----
        float* __restrict__ T1f = reinterpret_cast<float*>(in.T1p + (in.t1_red ? (size_t)b : (size_t)b * p.nmc + chunk) * (size_t)N * Pp);
        for (int l = 0; l < L; ++l) {
            mbar_wait(&d2_full[l & 1], (l >> 1) & 1);
            tc::tc_fence_after();
            float acc[16], a1[16], a2[16];
            tc::tmem_ld16x3(D[l & 1] + lane_base + 64 + 2 * n0, D[l & 1] + lane_base + 32 + 2 * n0, D[l & 1] + lane_base + 2 * n0, acc, a1, a2);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[j] = (acc[j] + a1[j]) + a2[j];          // lo + mid, then hi
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&d2_empty[l & 1]);
            // lane = kc: even lanes hold the e_re rows, odd lanes the e_im rows of the same antenna
            //   T1'r = D[(k,re),(n,re)] + D[(k,im),(n,im)] ; T1'i = D[(k,re),(n,im)] - D[(k,im),(n,re)]
#pragma unroll
            for (int r = 0; r < NH; ++r) {
                const float vre = acc[2 * r], vim = acc[2 * r + 1];
                const float other = __shfl_xor_sync(0xffffffffu, vim, 1);
                const float out = (lane & 1) ? (other - vre) : (vre + other);
                if (in.t1_red) atomicAdd(T1f + (size_t)(n0 + r) * 2 * Pp + KC * l + m, out);      // zeroed again by k_psi_res
                else T1f[(size_t)(n0 + r) * 2 * Pp + KC * l + m] = out;
            }
        }
        JSTSP_STAMP(p, 3, cta_id, 3);
        // ---- V2 goes out through S0 once the X store has read it ----
        if (tid == 0) bulk_wait_read();
        tc::worker_sync();
        tile_write8(S0, m, half, v2);
        tc::fence_async_smem();
        tc::worker_sync();
        if (tid == 0) { tma_store_3d(&maps.V2, S0, 0, c0, b); bulk_commit(); }

        // ---- partial Gram of the next SVT input (all MMAs have completed: the pilot tile is scratch now) ----
        {
            float* scratch = reinterpret_cast<float*>(tile);     // [slice][N*N][2]
            constexpr int NB4 = N / 4, COMBOS = NB4 * NB4, SLICES = WORKERS / COMBOS, CPS = MC / SLICES;
            static_assert(SLICES * N * N * 2 * 4 <= TILE, "Gram scratch does not fit the pilot tile");
            const int combo = tid % COMBOS, slice = tid / COMBOS, ib = combo % NB4, jb = combo / NB4;
            float ar[4][4] = {}, ai[4][4] = {};
            for (int c = slice * CPS; c < (slice + 1) * CPS; ++c) {
                float xr[4], xi[4], yr[4], yi[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { xr[u] = Zre[(ib * 4 + u) * ZP + c]; xi[u] = Zim[(ib * 4 + u) * ZP + c]; yr[u] = Zre[(jb * 4 + u) * ZP + c]; yi[u] = Zim[(jb * 4 + u) * ZP + c]; }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) cmac<float>(ar[u][v], ai[u][v], xr[u], xi[u], yr[v], -yi[v]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int i = ib * 4 + u, j = jb * 4 + v;
                    scratch[((size_t)slice * N * N + i + N * j) * 2] = ar[u][v];
                    scratch[((size_t)slice * N * N + i + N * j) * 2 + 1] = ai[u][v];
                }
            tc::worker_sync();
            double* out = p.gram + ((size_t)b * p.nmc + chunk) * 2 * N * N;
            for (int t = tid; t < N * N; t += WORKERS) {
                double re = 0.0, im = 0.0;
#pragma unroll 4
                for (int s = 0; s < SLICES; ++s) { re += (double)scratch[((size_t)s * N * N + t) * 2]; im += (double)scratch[((size_t)s * N * N + t) * 2 + 1]; }
                out[2 * t] = re; out[2 * t + 1] = im;
            }
        }
        if (tid == 0) bulk_wait_read();                          // shared memory must outlive the store's reads; the writes complete with the grid
        JSTSP_STAMP(p, 3, cta_id, 4);
        if (p.dbg && p.dbg_kernel == 3 && threadIdx.x == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.dbg[(size_t)cta_id * 8 + 7] = sm; }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == NWW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)TMEM_COLS));
}

// ---- host side: tensor maps --------------------------------------------------------------------------------------------------
// bf16 pilot image: dims {8, Mext, NKG, nE}, box {8, ROWS, NKG, 1}, no swizzle
inline bool make_map_e(const unsigned short* E, int nE, int M, CUtensorMap* map) {
    auto enc = tc::encode_fn();
    if (!enc) return false;
    const cuuint64_t Mext = (cuuint64_t)M + 8;
    cuuint64_t dims[4] = {8, Mext, (cuuint64_t)NKG, (cuuint64_t)nE};
    cuuint64_t strides[3] = {16, Mext * 16, Mext * 16 * NKG};
    cuuint32_t box[4] = {8, (cuuint32_t)ROWS, (cuuint32_t)NKG, 1}, es[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<unsigned short*>(E), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// N x M complex fp32 state array (column-major, one column = 128 bytes): dims {32 floats, M, nb}, box {32, 128, 1}, SWIZZLE_128B
inline bool make_map_state(const cx<float>* base, long long ld, int nb, int M, CUtensorMap* map) {
    auto enc = tc::encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {2 * N, (cuuint64_t)M, (cuuint64_t)(ld ? nb : 1)};
    cuuint64_t strides[2] = {2 * N * 4, (cuuint64_t)(ld ? ld : (long long)N * M) * 8};
    cuuint32_t box[3] = {2 * N, (cuuint32_t)MC, 1}, es[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<cx<float>*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace psi
}  // namespace jstsp
