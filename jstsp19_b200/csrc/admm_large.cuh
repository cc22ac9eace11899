// admm_large.cuh - large-array route of the proposed ADMM estimator (proposed_algorithm.m:35-65, 'approximate'), fp32.
//
// BASELINE config 4 (plot_errorVSdelays.m:45-49 shape at Nt = 256, Nr = 64, 128 frames, L = 8) has subY 64 x 32768 and a
// dictionary B of 2048 x 32768 - 512 MiB per trial in fp32.  It is never formed: with B_l = Dt' Psi_l and
// Psi_l(k, j) = e_k(j - l) (Hermitian Toeplitz 4-QAM pilots, plot_errorVSsnr.m:63-67) the three big products of an iteration
//     Xs  = sum_l (A S_l Dt') e(:, j - l)          proposed_algorithm.m:58
//     T1' = (K - XV) e(:, j - l)^H                 proposed_algorithm.m:47   (Res_l = A' T1'_l Dt)
//     G   = sum_l (A Res_l Dt') e(:, j - l)        |G|^2 = res' R res of the line search, proposed_algorithm.m:48
// run on the tensor cores out of ONE bf16 image of the pilot signs (32 MiB, L2 resident), the same restatement as admm_psi.cuh
// with the geometry turned around for 64 rows:
//   pass 1  (k_lg_mma<0>)  CTA = 128 columns of X.  A operand = the (128 + L - 1)-column window of e, resident in shared memory,
//           K-major, a delay tap is a 16-byte start offset; B operand = the image of Q_l = scale (A S_l) Dt' split into three bf16
//           terms that accumulate into the SAME 128 x 2N TMEM accumulator (3 MMAs of M128 N128 K16), streamed by bulk copies.
//   pass 2  (k_lg_mma<1>)  CTA = (tap l, 2048-column slice).  A operand = e window MN-major (lanes = (antenna, re|im), four
//           128-lane tiles = the full 512 TMEM columns), B operand = the three-term image of K - XV; raw fp32 partials per slice are
//           combined (complex products across lane / column pairs) and reduced in a fixed order by k_lg_t1c.
// The pilot operand is exact, the split operand carries 24 mantissa bits, accumulation is fp32: fp32-grade results.
// Everything else - SVT weights (the existing Jacobi kernel on the 64 x 64 Gram matrix), W Z, the Dt rotations and the products
// with A (batched complex GEMMs), the element-wise state update - runs in CUDA-core kernels around the two passes.
// Reductions are ordered (per-CTA partials summed in a fixed order), so results are repeatable bit for bit.
#pragma once

namespace jstsp {
namespace lg {

using psi::bf16_bits; using psi::bf16_val; using psi::split3; using psi::instr_desc_bf16; using psi::umma_bf16;

constexpr int WIN = 136;            // columns of e resident per 128-column tile (128 + L - 1 <= 135)
constexpr int NST = 3;              // stages of the operand rings
constexpr int SC2 = 32;             // columns per pass-2 stage
constexpr int MAXL = 8;
constexpr int CT = 64;              // column tile of the state kernel
constexpr int CPC = 512;            // columns per CTA of the state kernel (one Gram partial each)
constexpr int MMA_THREADS = 192;    // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- pilots -> bf16 sign image ------------------------------------------------------------------------------------------
// E[g = k / 4][te][2 (k % 4) + c] = e_k(te - (L - 1)) / scale,  e_k(t) = s_k(t) (t >= 0), conj(s_k(-t)) (t < 0), 0 beyond M.
// grid (ceil(Mext / 32), nE), block 256
__global__ void __launch_bounds__(256) k_lg_pack_e(const cx<float>* __restrict__ pil, long long ld_pil, unsigned short* __restrict__ E, float* __restrict__ scale, int* bad,
                                                   int Nt, int M, int Mext, int L) {
    const int b = blockIdx.y;
    const cx<float>* s = pil + (long long)b * ld_pil;
    unsigned short* Eb = E + (size_t)b * (Nt / 4) * Mext * 8;
    float sc = fabsf(s[0].re);
    if (sc == 0.f) sc = fabsf(s[0].im);
    if (sc == 0.f) sc = 1.f;
    if (blockIdx.x == 0 && threadIdx.x == 0) scale[b] = sc;
    int nbad = 0;
    for (int idx = threadIdx.x; idx < 32 * Nt; idx += 256) {
        const int k = idx % Nt, te = blockIdx.x * 32 + idx / Nt;
        if (te >= Mext) break;
        const int t = te - (L - 1);
        cx<float> e = mk<float>(0.f, 0.f);
        if (t >= 0 && t < M) e = s[k + (size_t)Nt * t];
        else if (t < 0) e = conj(s[k + (size_t)Nt * (-t)]);
        const unsigned short br = bf16_bits(e.re / sc), bi = bf16_bits(e.im / sc);
        if (t < M && (fabsf(bf16_val(br)) != 1.f || fabsf(bf16_val(bi)) != 1.f || bf16_val(br) * sc != e.re || bf16_val(bi) * sc != e.im)) ++nbad;
        *reinterpret_cast<uint32_t*>(Eb + ((size_t)(k / 4) * Mext + te) * 8 + 2 * (k % 4)) = (uint32_t)br | ((uint32_t)bi << 16);
    }
    if (nbad) atomicAdd(bad, nbad);
}
__global__ void k_lg_spread_scale(float* scale, int nb) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > 0 && t < nb) scale[t] = scale[0];
}

// ---- the two tensor-core passes ---------------------------------------------------------------------------------------------
struct MmaArgs {
    const unsigned char* img; long long ld_img;   // operand image (bytes per trial)
    float* out; long long ld_out;                 // MODE 0: [M][2N] per trial; MODE 1: [ksplit][L][2Nt][2N] raw partials per trial
    double* gg;                                   // MODE 0: [b][tiles] |tile|^2 (or null)
    int N2, L, nkg, e_shared, cps, M;             // nkg = 2 Nt / 8; cps = columns per pass-2 slice
};
__host__ __device__ inline int lg_stage_bytes(int N2) { return 3 * 4 * N2 * 16; }                 // three splits x four 8-wide K groups x 2N rows
inline size_t lg_smem(int mode, int N2, int nkg) {
    const size_t st = mode == 0 ? (size_t)lg_stage_bytes(N2) : (size_t)lg_stage_bytes(N2) + (size_t)nkg * SC2 * 16;
    const size_t win = mode == 0 ? (((size_t)nkg * WIN * 16 + 1023) / 1024) * 1024 : 0;
    return win + NST * st + 256;
}

template <int MODE>
__global__ void __launch_bounds__(MMA_THREADS, 1) k_lg_mma(const __grid_constant__ CUtensorMap mapE, MmaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int N2 = a.N2, nkg = a.nkg, L = a.L;
    const int SB = lg_stage_bytes(N2);                          // split-operand bytes per stage
    const int EB = MODE == 1 ? nkg * SC2 * 16 : 0;              // pass 2: pilot bytes per stage
    const int STG = SB + EB;
    const int win_bytes = MODE == 0 ? ((nkg * WIN * 16 + 1023) / 1024) * 1024 : 0;
    unsigned char* ring = smem + win_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + NST * STG);
    uint64_t *full = bars, *empty = bars + NST, *e_full = bars + 2 * NST, *acc_full = bars + 2 * NST + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2);
    const int ntile = MODE == 1 ? nkg / 16 : 1;
    const int ncols = ntile * N2;
    const uint32_t tcols = ncols <= 32 ? 32 : ncols <= 64 ? 64 : ncols <= 128 ? 128 : ncols <= 256 ? 256 : 512;

    const int b = MODE == 0 ? blockIdx.y : blockIdx.z;
    const int l2 = MODE == 1 ? blockIdx.y : 0;
    const int m0 = MODE == 0 ? blockIdx.x * 128 : blockIdx.x * a.cps;
    const int nit = MODE == 0 ? L * (nkg / 4) : a.cps / SC2;
    const int eg0 = a.e_shared ? 0 : b * nkg;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(e_full, 1); mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tcols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            if (MODE == 0) { mbar_expect_tx(e_full, nkg * WIN * 16); tc::tma_3d(smem, &mapE, 0, m0, eg0, e_full); }
            const unsigned char* src = a.img + (long long)b * a.ld_img + (MODE == 0 ? 0ll : (long long)(m0 / SC2) * SB);
            for (int it = 0; it < nit; ++it) {
                const int s = it % NST;
                if (it >= NST) mbar_wait(&empty[s], ((it / NST) - 1) & 1);
                unsigned char* dst = ring + s * STG;
                mbar_expect_tx(&full[s], STG);
                if (MODE == 1) tc::tma_3d(dst + SB, &mapE, 0, m0 + it * SC2 + (L - 1 - l2), eg0, &full[s]);
                tma_bulk_g2s(dst, src + (long long)it * SB, SB, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            if (MODE == 0) {
                const uint32_t id = instr_desc_bf16(128, N2, 0);
                const uint32_t RS = WIN * 16;
                mbar_wait(e_full, 0);
                for (int it = 0; it < nit; ++it) {
                    const int s = it % NST, l = it / (nkg / 4), jb = it % (nkg / 4);
                    mbar_wait(&full[s], (it / NST) & 1);
                    tc::tc_fence_after();
                    const uint32_t a0 = smem_u32(smem) + (L - 1 - l) * 16 + jb * 4 * RS, b0 = smem_u32(ring + s * STG);
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                        for (int sp = 0; sp < 3; ++sp)
                            umma_bf16(tm, tc::smem_desc(a0 + ks * 2 * RS, RS, 128, 0), tc::smem_desc(b0 + sp * 4 * N2 * 16 + ks * 2 * N2 * 16, N2 * 16, 128, 0), id,
                                      (it | ks | sp) ? 1u : 0u);
                    tc::umma_commit(&empty[s]);
                }
            } else {
                const uint32_t id = instr_desc_bf16(128, N2, 1);
                for (int it = 0; it < nit; ++it) {
                    const int s = it % NST;
                    mbar_wait(&full[s], (it / NST) & 1);
                    tc::tc_fence_after();
                    const uint32_t b0 = smem_u32(ring + s * STG), a0 = b0 + SB;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        for (int t = 0; t < ntile; ++t)
#pragma unroll
                            for (int sp = 0; sp < 3; ++sp)
                                umma_bf16(tm + t * N2, tc::smem_desc(a0 + t * 16 * (SC2 * 16) + ks * 256, 128, SC2 * 16, 0),
                                          tc::smem_desc(b0 + sp * 4 * N2 * 16 + ks * 2 * N2 * 16, N2 * 16, 128, 0), id, (it | ks | sp) ? 1u : 0u);
                    tc::umma_commit(&empty[s]);
                }
            }
            tc::umma_commit(acc_full);
        }
    } else {
        const int q = warp % 4;                                    // TMEM lane quarter this warp may read
        mbar_wait(acc_full, 0);
        tc::tc_fence_after();
        const uint32_t tl = tm + ((uint32_t)(32 * q) << 16);
        if (MODE == 0) {
            const int m = m0 + 32 * q + lane;
            float* o = a.out + (long long)b * a.ld_out + (size_t)m * N2;
            float ss = 0.f;
            for (int c = 0; c < N2; c += 16) {
                float v[16];
                tc::tmem_ld16(tl + c, v);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    *reinterpret_cast<float4*>(o + c + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    ss = fmaf(v[4 * u], v[4 * u], ss); ss = fmaf(v[4 * u + 1], v[4 * u + 1], ss); ss = fmaf(v[4 * u + 2], v[4 * u + 2], ss); ss = fmaf(v[4 * u + 3], v[4 * u + 3], ss);
                }
            }
            if (a.gg) {
                double d = (double)ss;
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o2);
                double* red = reinterpret_cast<double*>(bars + 2 * NST + 4);
                if (lane == 0) red[q] = d;
                asm volatile("bar.sync 2, 128;" ::: "memory");
                if (warp == 2 && lane == 0) a.gg[(size_t)b * gridDim.x + blockIdx.x] = red[0] + red[1] + red[2] + red[3];
            }
        } else {
            float* o = a.out + (long long)b * a.ld_out + (((size_t)blockIdx.x * L + l2) * (size_t)(nkg * 8)) * N2;
            for (int t = 0; t < ntile; ++t) {
                float* row = o + (size_t)(t * 128 + 32 * q + lane) * N2;
                for (int c = 0; c < N2; c += 16) {
                    float v[16];
                    tc::tmem_ld16(tl + t * N2 + c, v);
#pragma unroll
                    for (int u = 0; u < 4; ++u) *reinterpret_cast<float4*>(row + c + 4 * u) = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, tcols);
}

// raw pass-2 partials -> T1'[b][l][a][n] = scale * sum_slices ( P[(a,re)][(n,re)] + P[(a,im)][(n,im)],  P[(a,re)][(n,im)] - P[(a,im)][(n,re)] )
// grid (L * Nt / 4, nb), block (N, 4)
__global__ void k_lg_t1c(const float* __restrict__ part, long long ld_part, cx<float>* __restrict__ T1, const float* __restrict__ scale, int N, int Nt, int L, int KS) {
    const int b = blockIdx.y, n = threadIdx.x, la = blockIdx.x * 4 + threadIdx.y;      // la = l * Nt + a
    const int N2 = 2 * N;
    const float* p = part + (long long)b * ld_part + (size_t)la * 2 * N2 + 2 * n;
    const size_t slice = (size_t)L * Nt * 2 * N2;
    float re = 0.f, im = 0.f;
    for (int k = 0; k < KS; ++k) {
        const float2 r0 = *reinterpret_cast<const float2*>(p + k * slice), r1 = *reinterpret_cast<const float2*>(p + k * slice + N2);
        re += r0.x + r1.y; im += r0.y - r1.x;
    }
    const float sc = scale[b];
    T1[(size_t)b * L * Nt * N + (size_t)la * N + n] = mk<float>(sc * re, sc * im);
}

// ---- batched complex GEMM on the CUDA cores: C (m x n) = op(A) (m x k) op(B) (k x n), column-major ------------------------------------
// batch index z = z1 * nz2 + z2; operand offsets z1 * s?1 + z2 * s?2.  64 x 64 outputs per CTA, 4 x 4 per thread.
enum { OPN = 0, OPH = 1 };
struct GemmArgs {
    int m, n, k, nz2;
    const cx<float>* A; long long sA1, sA2; int ldA, opA;
    const cx<float>* B; long long sB1, sB2; int ldB, opB;
    cx<float>* C; long long sC1, sC2; int ldC;
};
__global__ void __launch_bounds__(256) k_cgemm64(GemmArgs g) {
    __shared__ __align__(16) cx<float> As[16][64], Bs[16][64];
    const int z1 = blockIdx.z / g.nz2, z2 = blockIdx.z % g.nz2;
    const cx<float>* A = g.A + z1 * g.sA1 + z2 * g.sA2;
    const cx<float>* B = g.B + z1 * g.sB1 + z2 * g.sB2;
    cx<float>* C = g.C + z1 * g.sC1 + z2 * g.sC2;
    const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64, tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float cr[4][4] = {}, ci[4][4] = {};
    for (int k0 = 0; k0 < g.k; k0 += 16) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = threadIdx.x + 256 * u;
            {   // As[kk][i]
                int i, kk;
                if (g.opA == OPN) { i = e % 64; kk = e / 64; } else { kk = e % 16; i = e / 16; }
                cx<float> v = mk<float>(0.f, 0.f);
                if (i0 + i < g.m && k0 + kk < g.k) v = g.opA == OPN ? A[(i0 + i) + (size_t)g.ldA * (k0 + kk)] : conj(A[(k0 + kk) + (size_t)g.ldA * (i0 + i)]);
                As[kk][i] = v;
            }
            {   // Bs[kk][j]
                int j, kk;
                if (g.opB == OPN) { kk = e % 16; j = e / 16; } else { j = e % 64; kk = e / 64; }
                cx<float> v = mk<float>(0.f, 0.f);
                if (j0 + j < g.n && k0 + kk < g.k) v = g.opB == OPN ? B[(k0 + kk) + (size_t)g.ldB * (j0 + j)] : conj(B[(j0 + j) + (size_t)g.ldB * (k0 + kk)]);
                Bs[kk][j] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            cx<float> av[4], bv[4];
            *reinterpret_cast<float4*>(&av[0]) = *reinterpret_cast<const float4*>(&As[kk][4 * ty]);
            *reinterpret_cast<float4*>(&av[2]) = *reinterpret_cast<const float4*>(&As[kk][4 * ty + 2]);
            *reinterpret_cast<float4*>(&bv[0]) = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx]);
            *reinterpret_cast<float4*>(&bv[2]) = *reinterpret_cast<const float4*>(&Bs[kk][4 * tx + 2]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) cmac<float>(cr[i][j], ci[i][j], av[i].re, av[i].im, bv[j].re, bv[j].im);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = i0 + 4 * ty + i, c = j0 + 4 * tx + j;
            if (r < g.m && c < g.n) C[r + (size_t)g.ldC * c] = mk<float>(cr[i][j], ci[i][j]);
        }
}

// Q[b][l][a][n] (N x Nt per tap) -> pass-1 operand image [l][kc / 32][split][(kc % 32) / 8][row = 2 n + c][kc % 8], kc = 2 a + (re|im):
// rows (n,re) = scale [Qr, -Qi], rows (n,im) = scale [Qi, Qr].  One thread per (tap, 8-wide kc group, row).
__global__ void __launch_bounds__(256) k_lg_qimg(const cx<float>* __restrict__ Q, unsigned char* __restrict__ img, long long ld_img, const float* __restrict__ scale, int N, int Nt, int L) {
    const int b = blockIdx.y, N2 = 2 * N, nkg = Nt / 4;
    const int u = blockIdx.x * 256 + threadIdx.x;
    if (u >= L * nkg * N2) return;
    const int r = u % N2, g = (u / N2) % nkg, l = u / (N2 * nkg);
    const int n = r / 2, c = r % 2;
    const float sc = scale[b];
    const cx<float>* q = Q + ((size_t)b * L + l) * (size_t)N * Nt + n;
    unsigned short s[8][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const cx<float> v = q[(size_t)N * (4 * g + i)];
        split3(sc * (c ? v.im : v.re), s[2 * i]);
        split3(sc * (c ? v.re : -v.im), s[2 * i + 1]);
    }
    unsigned char* base = img + (long long)b * ld_img + ((size_t)(l * (nkg / 4) + g / 4) * 3) * (4 * N2 * 16) + ((size_t)(g % 4) * N2 + r) * 16;
#pragma unroll
    for (int sp = 0; sp < 3; ++sp) {
        uint4 w;
        w.x = (uint32_t)s[0][sp] | ((uint32_t)s[1][sp] << 16); w.y = (uint32_t)s[2][sp] | ((uint32_t)s[3][sp] << 16);
        w.z = (uint32_t)s[4][sp] | ((uint32_t)s[5][sp] << 16); w.w = (uint32_t)s[6][sp] | ((uint32_t)s[7][sp] << 16);
        *reinterpret_cast<uint4*>(base + (size_t)sp * (4 * N2 * 16)) = w;
    }
}

// ---- element-wise state update (the closing step of the previous iteration - XV += alpha G, C, V2: proposed_algorithm.m:60,62 - folded in; C is never
//      stored), W Z, the pass-2 operand and the next Gram matrix (proposed_algorithm.m:37-47,61) -------------------------------------------------
struct StateArgs {
    int N, M, last;
    const cx<float>* subY; long long ld_subY;
    const float* omega; long long ld_omega;
    const double* rho;
    const cx<float>* W;                    // [b][N*N]
    cx<float> *X, *V1, *Y;                 // [b][M][N]; Y only written when last
    long long ld_Y;
    cx<float> *V2, *XV;                     // updated in place: the closing step of the previous iteration (proposed_algorithm.m:60,62 and XV += alpha G) is folded in here
    const cx<float> *Xs, *Gm; const float* alpha;
    unsigned char* dimg; long long ld_dimg;    // [M / 32][split][m group][2N][8]
    double* gram;                              // [b][M / CPC][N*N][2]
};
// grid (M / CPC, nb), block 256: thread (ty, tx) owns rows 4 ty.., columns 4 tx.. of each 64-column tile
__global__ void __launch_bounds__(256, 2) k_lg_state(StateArgs s) {
    extern __shared__ __align__(16) unsigned char smem[];
    cx<float>* Ws = reinterpret_cast<cx<float>*>(smem);                 // [k][n] = W(n, k), 64 x 64 zero padded
    cx<float>* Zs = Ws + 64 * 64;                                       // [k][m] Z tile, then [i][m] Z' tile
    float* Ds = reinterpret_cast<float*>(Zs + 64 * 64);                 // [2N rows][65]
    const int b = blockIdx.y, N = s.N, N2 = 2 * N, tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const float rho = (float)s.rho[b], irho = 1.f / rho, kap = rho / (rho + 1.f), al = s.alpha[b];
    const size_t off = (size_t)b * s.M * N;
    // operands of the two FMA loops are held PLANAR (re plane | im plane) so that two neighbouring rows / columns are one 64-bit operand of the packed fp32x2 FMA
    float* Wre = reinterpret_cast<float*>(Ws); float* Wim = Wre + 64 * 64;      // [k][n]
    float* Zre = reinterpret_cast<float*>(Zs); float* Zim = Zre + 64 * 64;      // [k][m], then Z' as [m][n]
    auto pk = [](float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; };
    auto f2 = [](uint64_t& d, uint64_t a, uint64_t bb) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(bb)); };
    auto up = [](uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); };
    for (int t = threadIdx.x; t < 64 * 64; t += 256) { const int n = t % 64, k = t / 64; const cx<float> w = (n < N && k < N) ? s.W[(size_t)b * N * N + n + N * k] : mk<float>(0.f, 0.f); Wre[t] = w.re; Wim[t] = w.im; }
    double gr[4][4] = {}, gi[4][4] = {};
    for (int tile = 0; tile < CPC / CT; ++tile) {
        const int m0 = blockIdx.x * CPC + tile * CT;
        __syncthreads();
        for (int t = threadIdx.x; t < 64 * 64; t += 256) {
            const int k = t % 64, m = t / 64;
            cx<float> z = mk<float>(0.f, 0.f);
            if (k < N) { const size_t i = off + (size_t)(m0 + m) * N + k; const cx<float> x = s.X[i], v = s.V1[i]; z = mk<float>(x.re - v.re * irho, x.im - v.im * irho); }
            Zre[k * 64 + m] = z.re; Zim[k * 64 + m] = z.im;
        }
        __syncthreads();
        float yr[4][4], yi[4][4];
        {   // Y = W Z: accumulators pair two columns; per k and row i: YR += w.re zr ; YR += (-w.im) zi ; YI += w.re zi ; YI += w.im zr  (cmac()'s operations and order)
            uint64_t YR[4][2] = {}, YI[4][2] = {};
            for (int k = 0; k < N; ++k) {
                const float4 wr = *reinterpret_cast<const float4*>(Wre + k * 64 + 4 * ty), wi = *reinterpret_cast<const float4*>(Wim + k * 64 + 4 * ty);
                const ulonglong2 zr = *reinterpret_cast<const ulonglong2*>(Zre + k * 64 + 4 * tx), zi = *reinterpret_cast<const ulonglong2*>(Zim + k * 64 + 4 * tx);
                const float wrs[4] = {wr.x, wr.y, wr.z, wr.w}, wis[4] = {wi.x, wi.y, wi.z, wi.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint64_t WR = pk(wrs[i], wrs[i]), WI = pk(wis[i], wis[i]), WN = pk(-wis[i], -wis[i]);
                    f2(YR[i][0], WR, zr.x); f2(YR[i][0], WN, zi.x); f2(YI[i][0], WR, zi.x); f2(YI[i][0], WI, zr.x);
                    f2(YR[i][1], WR, zr.y); f2(YR[i][1], WN, zi.y); f2(YI[i][1], WR, zi.y); f2(YI[i][1], WI, zr.y);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { up(YR[i][0], yr[i][0], yr[i][1]); up(YR[i][1], yr[i][2], yr[i][3]); up(YI[i][0], yi[i][0], yi[i][1]); up(YI[i][1], yi[i][2], yi[i][3]); }
        }
        __syncthreads();                                               // Zs is rewritten with Z' ([m][n]) below
        if (4 * ty < N) {
            auto ld4 = [](const cx<float>* p, cx<float> (&v)[4]) {
                *reinterpret_cast<float4*>(&v[0]) = *reinterpret_cast<const float4*>(p);
                *reinterpret_cast<float4*>(&v[2]) = *reinterpret_cast<const float4*>(p + 2);
            };
            auto st4 = [](cx<float>* p, const cx<float> (&v)[4]) {
                *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(&v[0]);
                *reinterpret_cast<float4*>(p + 2) = *reinterpret_cast<const float4*>(&v[2]);
            };
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + 4 * tx + j, n0 = 4 * ty;
                const size_t idx = off + (size_t)m * N + n0;
                cx<float> v1[4], v2[4], c[4], xs[4], xv[4], sy[4], x[4], y[4], zn[4], gm[4], xp[4];
                ld4(s.V1 + idx, v1); ld4(s.V2 + idx, v2); ld4(s.Xs + idx, xs); ld4(s.XV + idx, xv); ld4(s.Gm + idx, gm); ld4(s.X + idx, xp);
#pragma unroll
                for (int i = 0; i < 4; ++i) {             // close the previous iteration: XV += alpha G; C = rho/(rho+1) (X - Xs - V2/rho); V2 += rho (C - X + Xs)   (.m:60,62)
                    xv[i] = mk<float>(xv[i].re + al * gm[i].re, xv[i].im + al * gm[i].im);
                    c[i] = mk<float>(kap * (xp[i].re - xs[i].re - v2[i].re * irho), kap * (xp[i].im - xs[i].im - v2[i].im * irho));
                    v2[i] = mk<float>(v2[i].re + rho * (c[i].re - xp[i].re + xs[i].re), v2[i].im + rho * (c[i].im - xp[i].im + xs[i].im));
                }
                ld4(s.subY + (long long)b * s.ld_subY + (size_t)m * N + n0, sy);
                const float4 om = *reinterpret_cast<const float4*>(s.omega + (long long)b * s.ld_omega + (size_t)m * N + n0);
                const float omv[4] = {om.x, om.y, om.z, om.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float d = 1.f / (omv[i] + 2.f * rho);                                                           // .m:14,40
                    y[i] = mk<float>(yr[i][j], yi[i][j]);
                    x[i] = mk<float>((v1[i].re + rho * y[i].re + sy[i].re + v2[i].re + rho * c[i].re + rho * xs[i].re) * d,
                                     (v1[i].im + rho * y[i].im + sy[i].im + v2[i].im + rho * c[i].im + rho * xs[i].im) * d);  // .m:40
                    v1[i] = mk<float>(v1[i].re + rho * (y[i].re - x[i].re), v1[i].im + rho * (y[i].im - x[i].im));         // .m:61
                    // K - XV, K = X - V2/rho - C   (.m:43-44)
                    Ds[(2 * (n0 + i)) * 65 + 4 * tx + j] = x[i].re - v2[i].re * irho - c[i].re - xv[i].re;
                    Ds[(2 * (n0 + i) + 1) * 65 + 4 * tx + j] = x[i].im - v2[i].im * irho - c[i].im - xv[i].im;
                    zn[i] = mk<float>(x[i].re - v1[i].re * irho, x[i].im - v1[i].im * irho);      // next SVT input
                }
                st4(s.X + idx, x); st4(s.V1 + idx, v1); st4(s.V2 + idx, v2); st4(s.XV + idx, xv);
                if (s.last) st4(s.Y + (long long)b * s.ld_Y + (size_t)m * N + n0, y);
                *reinterpret_cast<float4*>(Zre + (4 * tx + j) * 64 + n0) = make_float4(zn[0].re, zn[1].re, zn[2].re, zn[3].re);
                *reinterpret_cast<float4*>(Zim + (4 * tx + j) * 64 + n0) = make_float4(zn[0].im, zn[1].im, zn[2].im, zn[3].im);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { *reinterpret_cast<float4*>(Zre + (4 * tx + j) * 64 + 4 * ty) = make_float4(0.f, 0.f, 0.f, 0.f); *reinterpret_cast<float4*>(Zim + (4 * tx + j) * 64 + 4 * ty) = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
        __syncthreads();
        // pass-2 operand: unit = (32-column block, 8-column group, row); three bf16 terms
        for (int u = threadIdx.x; u < 2 * 4 * N2; u += 256) {
            const int r = u % N2, mg = (u / N2) % 4, mb = u / (4 * N2);
            const float* src = Ds + r * 65 + mb * 32 + mg * 8;
            unsigned short sp[8][3];
#pragma unroll
            for (int i = 0; i < 8; ++i) split3(src[i], sp[i]);
            unsigned char* base = s.dimg + (long long)b * s.ld_dimg + (size_t)(m0 / 32 + mb) * (3 * 4 * N2 * 16) + ((size_t)mg * N2 + r) * 16;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                uint4 w;
                w.x = (uint32_t)sp[0][q] | ((uint32_t)sp[1][q] << 16); w.y = (uint32_t)sp[2][q] | ((uint32_t)sp[3][q] << 16);
                w.z = (uint32_t)sp[4][q] | ((uint32_t)sp[5][q] << 16); w.w = (uint32_t)sp[6][q] | ((uint32_t)sp[7][q] << 16);
                *reinterpret_cast<uint4*>(base + (size_t)q * (4 * N2 * 16)) = w;
            }
        }
        // Gram of the next SVT input over this tile: fp32 inside the tile, fp64 across tiles
        float tr[4][4], ti[4][4];
        {   // G += Z' Z'^H over the tile's columns: accumulators pair two rows i; per column and j: TR += a.re b.re ; TR += a.im b.im ; TI += a.re (-b.im) ; TI += a.im b.re
            uint64_t TR[2][4] = {}, TI[2][4] = {};
            for (int m = 0; m < 64; ++m) {
                const ulonglong2 ar = *reinterpret_cast<const ulonglong2*>(Zre + m * 64 + 4 * ty), ai = *reinterpret_cast<const ulonglong2*>(Zim + m * 64 + 4 * ty);
                const float4 br = *reinterpret_cast<const float4*>(Zre + m * 64 + 4 * tx), bi = *reinterpret_cast<const float4*>(Zim + m * 64 + 4 * tx);
                const float brs[4] = {br.x, br.y, br.z, br.w}, bis[4] = {bi.x, bi.y, bi.z, bi.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t BR = pk(brs[j], brs[j]), BI = pk(bis[j], bis[j]), BN = pk(-bis[j], -bis[j]);
                    f2(TR[0][j], ar.x, BR); f2(TR[0][j], ai.x, BI); f2(TI[0][j], ar.x, BN); f2(TI[0][j], ai.x, BR);
                    f2(TR[1][j], ar.y, BR); f2(TR[1][j], ai.y, BI); f2(TI[1][j], ar.y, BN); f2(TI[1][j], ai.y, BR);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { up(TR[0][j], tr[0][j], tr[1][j]); up(TR[1][j], tr[2][j], tr[3][j]); up(TI[0][j], ti[0][j], ti[1][j]); up(TI[1][j], ti[2][j], ti[3][j]); }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { gr[i][j] += (double)tr[i][j]; gi[i][j] += (double)ti[i][j]; }
    }
    double* g = s.gram + ((size_t)b * gridDim.x + blockIdx.x) * 2 * N * N;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = 4 * ty + i, c = 4 * tx + j;
            if (r < N && c < N) { g[2 * (r + N * c)] = gr[i][j]; g[2 * (r + N * c) + 1] = gi[i][j]; }
        }
}

// partial sums of |x|^2: grid (64, nb)
__global__ void __launch_bounds__(256) k_lg_sumsq(const cx<float>* __restrict__ x, size_t per, double* __restrict__ part) {
    const int b = blockIdx.y;
    const cx<float>* p = x + (size_t)b * per;
    double acc = 0.0;
    for (size_t t = blockIdx.x * 256 + threadIdx.x; t < per; t += (size_t)gridDim.x * 256) { const cx<float> v = p[t]; acc += (double)v.re * v.re + (double)v.im * v.im; }
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0.0; for (int k = 0; k < 8; ++k) t += red[k]; part[(size_t)b * gridDim.x + blockIdx.x] = t; }
}
// alpha = res'res / (res' R res) = |Res|^2 / |G|^2 (proposed_algorithm.m:48); 0/0 -> NaN like MATLAB.  grid nb, one thread
__global__ void k_lg_alpha(const double* __restrict__ rr, int nrr, const double* __restrict__ gg, int ngg, float* __restrict__ alpha) {
    const int b = blockIdx.x;
    double r = 0.0, g = 0.0;
    for (int k = 0; k < nrr; ++k) r += rr[(size_t)b * nrr + k];
    for (int k = 0; k < ngg; ++k) g += gg[(size_t)b * ngg + k];
    alpha[b] = (float)(r / g);
}
// V += alpha Res; S = soft(V, tau_S / rho) on real and imaginary parts (proposed_algorithm.m:50,56)
__global__ void __launch_bounds__(256) k_lg_vstep(cx<float>* __restrict__ V, const cx<float>* __restrict__ Res, cx<float>* __restrict__ S, const float* __restrict__ alpha,
                                                  const double* __restrict__ tauS, const double* __restrict__ rho, size_t per, const unsigned char* __restrict__ mask) {
    const int b = blockIdx.y;
    const size_t t = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (t >= per) return;
    const float al = alpha[b], thr = (float)(tauS[b] / rho[b]);
    const size_t i = (size_t)b * per + t;
    cx<float> v = V[i]; const cx<float> r = Res[i];
    v = mk<float>(v.re + al * r.re, v.im + al * r.im);
    V[i] = v;
    cx<float> sv = mk<float>(soft1<float>(v.re, thr), soft1<float>(v.im, thr));
    if (mask && !mask[i]) sv = mk<float>(0.f, 0.f);                  // K3 * s: the support ranking of proposed_algorithm_angles.m:36,68
    S[i] = sv;
}
inline bool make_map_e(const unsigned short* E, int Mext, int groups_total, int box_cols, int box_groups, CUtensorMap* map) {
    auto enc = tc::encode_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {8, (cuuint64_t)Mext, (cuuint64_t)groups_total};
    cuuint64_t strides[2] = {16, (cuuint64_t)Mext * 16};
    cuuint32_t box[3] = {8, (cuuint32_t)box_cols, (cuuint32_t)box_groups}, es[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<unsigned short*>(E), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shapes this route takes (fp32, 'approximate', pilots entry, no diagnostics, no support ranking)
inline bool large_shape(int N, int M, int G, int Nt, int Gt, int L) {
    return N % 8 == 0 && N >= 16 && N <= 64 && G >= 1 && G <= 64 && Nt % 64 == 0 && Nt <= 256 && Gt >= 1 && L >= 1 && L <= MAXL && M % CPC == 0 && !(N == 16 && Nt == 64);
}

// ---- host driver ----------------------------------------------------------------------------------------------------------------
static int run_large(Handle* h, const jstsp_admm_desc* d, int mem, const void* subY_, const void* omega_, const void* A_, const PsiArgs* ps,
                     const double* tauY_, const double* tauS_, const double* rho_, void* S_, void* Y_, const int* indx_ = nullptr) {
    const int N = d->N, M = d->M, G = d->G, P = d->P, batch = d->batch, imax = d->imax, Nt = ps->Nt, Gt = ps->Gt, L = ps->L;
    const int N2 = 2 * N, nkg = Nt / 4, Mext = M + 8;
    const int cps = M % 2048 == 0 ? 2048 : CPC, KS = M / cps, ngram = M / CPC, ntile = M / 128;
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    if (!tc::encode_fn()) return fail(h, JSTSP_E_CUDA, "cuTensorMapEncodeTiled is not available");
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    if (!host && (!al16(subY_) || !al16(omega_) || !al16(Y_) || (d->ld_subY * 8) % 16 || (d->ld_omega * 4) % 16 || (d->ld_Y * 8) % 16))
        return fail(h, JSTSP_E_UNSUPPORTED, "large-array route: device buffers must be 16-byte aligned");
    const size_t NM = (size_t)N * M, GP = (size_t)G * P, NP = (size_t)N * P, NLN = (size_t)N * L * Nt;
    const bool shA = d->ld_A == 0, shD = ps->ld_Dt == 0, shP = ps->ld_Psi == 0;
    const size_t SB = (size_t)lg_stage_bytes(N2);
    const size_t dimg_b = (size_t)(M / SC2) * SB, qimg_b = (size_t)L * (nkg / 4) * SB, part_f = (size_t)KS * L * 2 * Nt * N2, e_us = (size_t)nkg * Mext * 8;
    // trials per pass: ~0.3 GB of state and operands per trial at config 4
    const size_t per_trial = 8 * NM * 8 + NM * 4 + dimg_b + qimg_b + part_f * 4 + e_us * 2 + (size_t)Nt * M * 8 + (size_t)ngram * 2 * N * N * 8 + 8 * GP * 8 + 4 * NLN * 8;
    int chunk = batch;
    if (h->max_chunk > 0 && chunk > h->max_chunk) chunk = h->max_chunk;
    size_t freeb = 0, totalb = 0;
    cudaMemGetInfo(&freeb, &totalb);
    const size_t budget = (freeb + h->ws_bytes) / 2;
    while (chunk > 1 && (size_t)chunk * per_trial > budget) chunk = (chunk + 1) / 2;
    if (chunk > 16) chunk = 16;
    cx<float> *X, *V1, *V2, *Xs, *XV, *Gm, *Yb, *sY, *dA, *dDt, *dPil, *T1c, *R1, *Res, *V, *S, *AR, *Q, *W;
    float *om, *scale, *part, *alpha; unsigned short* E; unsigned char *dimg, *qimg, *smask = nullptr; double *rr, *gg, *gram, *Uprev, *dtau, *dtaus, *drho; int *bad, *dindx = nullptr;
    const bool angles = indx_ != nullptr;
    const bool shI = d->ld_indx == 0;
    auto layout = [&](Arena& a, int nb) {
        X = a.take<cx<float>>(NM * nb); V1 = a.take<cx<float>>(NM * nb); V2 = a.take<cx<float>>(NM * nb);
        Xs = a.take<cx<float>>(NM * nb); XV = a.take<cx<float>>(NM * nb); Gm = a.take<cx<float>>(NM * nb);
        Yb = host ? a.take<cx<float>>(NM * nb) : nullptr;
        sY = host ? a.take<cx<float>>(NM * nb) : nullptr; om = host ? a.take<float>(NM * nb) : nullptr;
        dA = host ? a.take<cx<float>>((size_t)N * G * (shA ? 1 : nb)) : nullptr;
        dDt = host ? a.take<cx<float>>((size_t)Nt * Gt * (shD ? 1 : nb)) : nullptr;
        dPil = host ? a.take<cx<float>>((size_t)Nt * M * (shP ? 1 : nb)) : nullptr;
        E = a.take<unsigned short>(e_us * (shP ? 1 : nb)); scale = a.take<float>(nb); bad = a.take<int>(4); alpha = a.take<float>(nb);
        dimg = a.take<unsigned char>(dimg_b * nb); qimg = a.take<unsigned char>(qimg_b * nb); part = a.take<float>(part_f * nb);
        T1c = a.take<cx<float>>(NLN * nb); Q = a.take<cx<float>>(NLN * nb); R1 = a.take<cx<float>>(NP * nb); AR = a.take<cx<float>>(NP * nb);
        Res = a.take<cx<float>>(GP * nb); V = a.take<cx<float>>(GP * nb); S = a.take<cx<float>>(GP * nb);
        rr = a.take<double>((size_t)64 * nb); gg = a.take<double>((size_t)ntile * nb);
        gram = a.take<double>((size_t)ngram * 2 * N * N * nb); Uprev = a.take<double>((size_t)2 * N * N * nb); W = a.take<cx<float>>((size_t)N * N * nb);
        dtau = a.take<double>(nb); dtaus = a.take<double>(nb); drho = a.take<double>(nb);
        if (angles) { smask = a.take<unsigned char>(GP * nb); dindx = host ? a.take<int>((size_t)d->n_indx * (shI ? 1 : nb)) : nullptr; }
    };
    { Arena probe(nullptr, 0); layout(probe, chunk); int rc = ensure_workspace(h, probe.off); if (rc) return rc; }
    const size_t sm_j = JacobiSmem::bytes(N) + 2 * sizeof(double) * (size_t)N * N + 16;
    const size_t sm_state = (size_t)2 * 64 * 64 * 8 + (size_t)N2 * 65 * 4;
    const size_t sm0 = lg_smem(0, N2, nkg), sm1 = lg_smem(1, N2, nkg);
    int rc;
    if ((rc = set_smem(h, k_svt_weights<float>, sm_j))) return rc;
    if ((rc = set_smem(h, k_lg_state, sm_state))) return rc;
    if ((rc = set_smem(h, k_lg_mma<0>, sm0))) return rc;
    if ((rc = set_smem(h, k_lg_mma<1>, sm1))) return rc;
    JSTSP_CUDA(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), st));
    auto gemm = [&](int slot, int m, int n, int k, int nz, int nz2, const cx<float>* A, long long sA1, long long sA2, int ldA, int opA, const cx<float>* B, long long sB1,
                    long long sB2, int ldB, int opB, cx<float>* Cc, long long sC1, long long sC2, int ldC) {
        GemmArgs g{m, n, k, nz2, A, sA1, sA2, ldA, opA, B, sB1, sB2, ldB, opB, Cc, sC1, sC2, ldC};
        dim3 grid(ceil_div(m, 64), ceil_div(n, 64), nz);
        JSTSP_LAUNCH(h, slot, (k_cgemm64<<<grid, 256, 0, st>>>(g)));
    };
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = (batch - b0) < chunk ? (batch - b0) : chunk;
        Arena ar(h->ws, h->ws_bytes);
        layout(ar, chunk);
        const cx<float>*pY, *pA, *pDt, *pPil; const float* pO; const double *pTauY, *pTauS, *pRho;
        long long ldY_in, ldO, ldA, ldDt, ldPil;
        if (host) {
            auto up = [&](void* dst, const void* src, size_t per, long long ld, size_t el) -> cudaError_t {
                if (ld == 0) return cudaMemcpyAsync(dst, src, per * el, cudaMemcpyHostToDevice, st);
                return cudaMemcpy2DAsync(dst, per * el, (const char*)src + (size_t)b0 * ld * el, (size_t)ld * el, per * el, nb, cudaMemcpyHostToDevice, st);
            };
            JSTSP_CUDA(h, up(sY, subY_, NM, d->ld_subY ? d->ld_subY : (batch == 1 ? (long long)NM : 0), 8));
            JSTSP_CUDA(h, up(om, omega_, NM, d->ld_omega ? d->ld_omega : (batch == 1 ? (long long)NM : 0), 4));
            JSTSP_CUDA(h, up(dA, A_, (size_t)N * G, d->ld_A, 8));
            JSTSP_CUDA(h, up(dDt, ps->Dt, (size_t)Nt * Gt, ps->ld_Dt, 8));
            JSTSP_CUDA(h, up(dPil, ps->Psi, (size_t)Nt * M, ps->ld_Psi, 8));
            JSTSP_CUDA(h, cudaMemcpyAsync(dtau, tauY_ + b0, 8 * nb, cudaMemcpyHostToDevice, st));
            JSTSP_CUDA(h, cudaMemcpyAsync(dtaus, tauS_ + b0, 8 * nb, cudaMemcpyHostToDevice, st));
            JSTSP_CUDA(h, cudaMemcpyAsync(drho, rho_ + b0, 8 * nb, cudaMemcpyHostToDevice, st));
            pY = sY; ldY_in = (d->ld_subY || batch == 1) ? (long long)NM : 0; pO = om; ldO = (d->ld_omega || batch == 1) ? (long long)NM : 0;
            pA = dA; ldA = shA ? 0 : (long long)N * G; pDt = dDt; ldDt = shD ? 0 : (long long)Nt * Gt; pPil = dPil; ldPil = shP ? 0 : (long long)Nt * M;
            pTauY = dtau; pTauS = dtaus; pRho = drho;
            if (angles) JSTSP_CUDA(h, up(dindx, indx_, (size_t)d->n_indx, d->ld_indx, sizeof(int)));
        } else {
            pY = (const cx<float>*)subY_ + (long long)b0 * d->ld_subY; ldY_in = d->ld_subY; pO = (const float*)omega_ + (long long)b0 * d->ld_omega; ldO = d->ld_omega;
            pA = (const cx<float>*)A_ + (long long)b0 * d->ld_A; ldA = d->ld_A; pDt = (const cx<float>*)ps->Dt + (long long)b0 * ps->ld_Dt; ldDt = ps->ld_Dt;
            pPil = (const cx<float>*)ps->Psi + (long long)b0 * ps->ld_Psi; ldPil = ps->ld_Psi;
            pTauY = tauY_ + b0; pTauS = tauS_ + b0; pRho = rho_ + b0;
        }
        cx<float>* Yd = host ? Yb : (Y_ ? (cx<float>*)Y_ + (long long)b0 * d->ld_Y : nullptr);
        const long long ldYd = host ? (long long)NM : d->ld_Y;
        // pilots -> sign image, checked on the device
        const int nE = shP ? 1 : nb;
        JSTSP_CUDA(h, cudaMemsetAsync(bad, 0, sizeof(int) * 4, st));
        { dim3 g(ceil_div(Mext, 32), nE); JSTSP_LAUNCH(h, PK_SETUP, (k_lg_pack_e<<<g, 256, 0, st>>>(pPil, ldPil, E, scale, bad, Nt, M, Mext, L))); }
        if (shP && nb > 1) JSTSP_LAUNCH(h, PK_SETUP, (k_lg_spread_scale<<<ceil_div(nb, 256), 256, 0, st>>>(scale, nb)));
        int bad_h = 0;
        JSTSP_CUDA(h, cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        JSTSP_CUDA(h, cudaStreamSynchronize(st));
        if (bad_h) return fail(h, JSTSP_E_UNSUPPORTED, "large-array route needs 4-QAM pilot sequences (entries +-a +-ja, exact in bf16 after the common scaling)");
        CUtensorMap map0, map1;
        if (!make_map_e(E, Mext, nkg * nE, WIN, nkg, &map0) || !make_map_e(E, Mext, nkg * nE, SC2, nkg, &map1)) return fail(h, JSTSP_E_CUDA, "cuTensorMapEncodeTiled failed for the pilot image");
        for (cx<float>* z : {X, V1, V2, Xs, XV, Gm}) JSTSP_CUDA(h, cudaMemsetAsync(z, 0, 8 * NM * nb, st));
        JSTSP_CUDA(h, cudaMemsetAsync(alpha, 0, sizeof(float) * nb, st));
        if (angles) JSTSP_CUDA(h, cudaMemsetAsync(smask, 0, GP * nb, st));
        const int* pIndx = angles ? (host ? dindx : indx_ + (long long)b0 * d->ld_indx) : nullptr;
        const long long ldIndx = angles ? (host ? (shI ? 0 : (long long)d->n_indx) : d->ld_indx) : 0;
        JSTSP_CUDA(h, cudaMemsetAsync(V, 0, 8 * GP * nb, st));
        JSTSP_CUDA(h, cudaMemsetAsync(S, 0, 8 * GP * nb, st));
        JSTSP_CUDA(h, cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)ngram * 2 * N * N * nb, st));
        for (int it = 0; it < imax; ++it) {
            AdmmP<float> qe{};
            qe.N = N; qe.gram = gram; qe.nmc = ngram; qe.Uprev = Uprev; qe.iter = it; qe.tauY = pTauY; qe.rho = pRho; qe.W = W;
            // W(it) from the Gram matrix the previous state kernel left: iteration 0 on the main stream, later ones on the side stream,
            // hidden behind the three products of the previous iteration (one CTA per trial, ~1.7 ms at 64 rows)
            if (it == 0) JSTSP_LAUNCH(h, PK_EIG, (k_svt_weights<float><<<nb, 128, sm_j, st>>>(qe)));
            else JSTSP_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
            StateArgs sa{N, M, it + 1 == imax ? 1 : 0, pY, ldY_in, pO, ldO, pRho, W, X, V1, Yd, ldYd, V2, XV, Xs, Gm, alpha, dimg, (long long)dimg_b, gram};
            if (!Yd) sa.last = 0;
            { dim3 g(ngram, nb); JSTSP_LAUNCH(h, PK_LG_STATE, (k_lg_state<<<g, 256, sm_state, st>>>(sa))); }
            if (it + 1 < imax) {
                AdmmP<float> qn = qe; qn.iter = it + 1;
                JSTSP_CUDA(h, cudaEventRecord(h->ev_fork, st));
                JSTSP_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
                k_svt_weights<float><<<nb, 128, sm_j, h->side>>>(qn); h->launches++;
                JSTSP_CUDA(h, cudaEventRecord(h->ev_join, h->side));
            }
            MmaArgs m1{dimg, (long long)dimg_b, part, (long long)part_f, nullptr, N2, L, nkg, shP ? 1 : 0, cps, M};
            { dim3 g(KS, L, nb); JSTSP_LAUNCH(h, PK_LG_PASS2, (k_lg_mma<1><<<g, MMA_THREADS, sm1, st>>>(map1, m1))); }
            { dim3 g(L * Nt / 4, nb), blk(N, 4); JSTSP_LAUNCH(h, PK_LG_SMALL, (k_lg_t1c<<<g, blk, 0, st>>>(part, (long long)part_f, T1c, scale, N, Nt, L, KS))); }
            // Res_l = A' (T1'_l Dt)
            gemm(PK_LG_SMALL, N, Gt, Nt, nb * L, L, T1c, (long long)NLN, (long long)N * Nt, N, OPN, pDt, ldDt, 0, Nt, OPN, R1, (long long)NP, (long long)N * Gt, N);
            gemm(PK_LG_SMALL, G, P, N, nb, 1, pA, ldA, 0, N, OPH, R1, (long long)NP, 0, N, OPN, Res, (long long)GP, 0, G);
            { dim3 g(64, nb); JSTSP_LAUNCH(h, PK_LG_SMALL, (k_lg_sumsq<<<g, 256, 0, st>>>(Res, GP, rr))); }
            // G = (A Res_l Dt') e
            gemm(PK_LG_SMALL, N, P, G, nb, 1, pA, ldA, 0, N, OPN, Res, (long long)GP, 0, G, OPN, AR, (long long)NP, 0, N);
            gemm(PK_LG_SMALL, N, Nt, Gt, nb * L, L, AR, (long long)NP, (long long)N * Gt, N, OPN, pDt, ldDt, 0, Nt, OPH, Q, (long long)NLN, (long long)N * Nt, N);
            { dim3 g(ceil_div(L * nkg * N2, 256), nb); JSTSP_LAUNCH(h, PK_LG_SMALL, (k_lg_qimg<<<g, 256, 0, st>>>(Q, qimg, (long long)qimg_b, scale, N, Nt, L))); }
            MmaArgs m0{qimg, (long long)qimg_b, reinterpret_cast<float*>(Gm), (long long)NM * 2, gg, N2, L, nkg, shP ? 1 : 0, 0, M};
            { dim3 g(ntile, nb); JSTSP_LAUNCH(h, PK_LG_PASS1, (k_lg_mma<0><<<g, MMA_THREADS, sm0, st>>>(map0, m0))); }
            JSTSP_LAUNCH(h, PK_LG_SMALL, (k_lg_alpha<<<nb, 1, 0, st>>>(rr, 64, gg, ntile, alpha)));
            if (angles) {      // Omega_S(indx_S(1 : min(10 + 5 i, G P))) = 1 (proposed_algorithm_angles.m:36), the mask kernel of the dense route
                AdmmP<float> qm{};
                qm.G = G; qm.P = P; qm.iter = it; qm.n_indx = d->n_indx; qm.indx = pIndx; qm.ld_indx = ldIndx; qm.smask = smask;
                dim3 g(1, nb); JSTSP_LAUNCH(h, PK_LG_SMALL, (k_mask_grow<float><<<g, 64, 0, st>>>(qm)));
            }
            { dim3 g(ceil_div((int)GP, 256), nb); JSTSP_LAUNCH(h, PK_LG_SMALL, (k_lg_vstep<<<g, 256, 0, st>>>(V, Res, S, alpha, pTauS, pRho, GP, smask))); }
            // Xs = (A S_l Dt') e
            gemm(PK_LG_SMALL, N, P, G, nb, 1, pA, ldA, 0, N, OPN, S, (long long)GP, 0, G, OPN, AR, (long long)NP, 0, N);
            gemm(PK_LG_SMALL, N, Nt, Gt, nb * L, L, AR, (long long)NP, (long long)N * Gt, N, OPN, pDt, ldDt, 0, Nt, OPH, Q, (long long)NLN, (long long)N * Nt, N);
            { dim3 g(ceil_div(L * nkg * N2, 256), nb); JSTSP_LAUNCH(h, PK_LG_SMALL, (k_lg_qimg<<<g, 256, 0, st>>>(Q, qimg, (long long)qimg_b, scale, N, Nt, L))); }
            MmaArgs m2 = m0; m2.out = reinterpret_cast<float*>(Xs); m2.gg = nullptr;
            { dim3 g(ntile, nb); JSTSP_LAUNCH(h, PK_LG_PASS1, (k_lg_mma<0><<<g, MMA_THREADS, sm0, st>>>(map0, m2))); }
        }
        JSTSP_CUDA(h, cudaGetLastError());
        JSTSP_LAUNCH(h, PK_OTHER, (k_count_nonfinite<float><<<nb, 128, 0, st>>>(S, GP, nb, h->d_flag)));
        const long long ldS = d->ld_S ? d->ld_S : (long long)GP;
        if (host) {
            JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)S_ + (size_t)b0 * ldS * 8, (size_t)ldS * 8, S, GP * 8, GP * 8, nb, cudaMemcpyDeviceToHost, st));
            if (Y_) { const long long ldy = d->ld_Y ? d->ld_Y : (long long)NM; JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)Y_ + (size_t)b0 * ldy * 8, (size_t)ldy * 8, Yb, NM * 8, NM * 8, nb, cudaMemcpyDeviceToHost, st)); }
            JSTSP_CUDA(h, cudaStreamSynchronize(st));
        } else {
            JSTSP_CUDA(h, cudaMemcpy2DAsync((cx<float>*)S_ + (long long)b0 * ldS, (size_t)ldS * 8, S, GP * 8, GP * 8, nb, cudaMemcpyDeviceToDevice, st));
        }
    }
    h->last_path = 3; h->last_variant = 0;
    int nbad = 0;
    if (host) { JSTSP_CUDA(h, cudaMemcpyAsync(&nbad, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st)); JSTSP_CUDA(h, cudaStreamSynchronize(st)); }
    return nbad;
}

}  // namespace lg
}  // namespace jstsp
