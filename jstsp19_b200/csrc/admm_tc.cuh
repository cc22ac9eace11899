// admm_tc.cuh - tcgen05 (5th-gen tensor core) path of the proposed ADMM iteration, fp32 storage, 3xTF32 products.
//
// One kernel per iteration and (trial b, 128-column chunk) replaces k_xs_fast(i-1) + k_xupd_t1_fast(i):
//
//   pass 1   Xs(:,chunk) = (A S) B(:,chunk)                 (proposed_algorithm.m:58 of iteration i-1)
//   update   C, V2 (:61,:65 of i-1) ; Y = W Z (svt.m via SURVEY A.2) ; X, V1 (:38-40,:64 of i) ; K = X - V2/rho - C (:43)
//   pass 2   T1 partial = K(:,chunk) B(:,chunk)^H           (:47, summed over chunks by k_vstep_fast)
//   Gram     partial Gram of the next SVT input X - V1/rho  (consumed by k_svt_weights on the side stream)
//
// so C and Xs never touch HBM and the dictionary chunk B(:,chunk) (P x 128 complex, the dominant stream) is read
// from HBM once per iteration (pass 2 re-reads it from L2).
//
// Tensor-core formulation (all operands real fp32 interpreted as tf32, fp32 accumulation in TMEM):
//   * B is used in place, in its natural layout: row m of the 2-D float view [M][2P] is column m of B with (re,im)
//     interleaved.  TMA tensor copies bring 16 KiB tiles into a shared-memory ring:
//       pass 1: box 32 floats x 128 rows, SWIZZLE_128B            -> K-major  A operand (M = 128 columns of the chunk)
//       pass 2: box 32 floats x 32 rows x 4 groups, SWIZZLE_128B_ATOM_32B -> MN-major A operand (M = 128 floats = 64 p)
//     (MN-major tf32 exists only in that layout; both verified by tools/umma_probe2.cu).
//   * complex products are expressed through the small operand (the N side, K-major, SWIZZLE_NONE):
//       pass 1 rows (n,re) = [ASr, -ASi], (n,im) = [ASi, ASr] along k = (p,re),(p,im)   -> D1[m][(n,c)] = Xs^T
//       pass 2 rows (n,re) = Kr(n,:), (n,im) = Ki(n,:) along k = m                     -> D2[(p,cB)][(n,cK)], combined
//                                                                                         across lane pairs in the epilogue
//   * 3xTF32: the tensor core truncates fp32 inputs to tf32 (tools/umma_probe.cu), so the raw tile is the "hi" part.
//     hi MMA:  tile_hi x [S_hi ; S_lo]  (N = 4N rows, one pass over the tile for two of the three terms)
//     then the worker warps overwrite the tile IN PLACE with x - trunc(x) and a second MMA adds tile_lo x S_hi.
//
// Warp roles (320 threads): warps 0-7 workers (lo rewrite, element-wise update, epilogues, Gram), warp 8 TMA producer,
// warp 9 MMA issuer.  Two CTAs are resident per SM so that one CTA's element-wise phase overlaps the other's streaming.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include "stream_core.cuh"

namespace jstsp {
namespace tc {

constexpr int MC = 128;            // columns of X per CTA (= MMA M)
constexpr int STAGE = 16384;       // bytes per dictionary tile
constexpr int WORKERS = 256;
constexpr int NWW = WORKERS / 32;
constexpr int THREADS = WORKERS + 64;
constexpr int ZP = MC + 1;         // pitch of the planar Z tile [row][column]
constexpr int LAG = 2;             // the lo MMA of stage g is issued after the hi MMA of stage g + LAG - 1

template <int N, int NST> struct Geo {
    static constexpr int NS = 4 * N;                     // rows of the small operand: (n,c) hi | (n,c) lo
    static constexpr int NG = NS / 8;                    // 8-row core-matrix groups
    static constexpr int OP1 = 8 * NG * 128;             // bytes of one pass-1 operand slice (32 k-floats)
    static constexpr int KOP_LBO = NG * 128 + 16;        // pass-2 operand: stride between 4-float k groups (+16: bank spread)
    static constexpr int KOP = (MC / 4) * KOP_LBO;
    static constexpr int OPREG = ((NST * OP1 > KOP ? NST * OP1 : KOP) + 15) / 16 * 16;
    static constexpr int ZREG = 2 * N * ZP * 4;
    static constexpr int BARS = 256;
    static constexpr int TMEM_COLS = 4 * NS <= 32 ? 32 : 4 * NS <= 64 ? 64 : 4 * NS <= 128 ? 128 : 4 * NS <= 256 ? 256 : 512;   // 2 x D1, 2 x D2
    static constexpr size_t SMEM = (size_t)NST * STAGE + OPREG + ZREG + BARS;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)layout << 61);          // layout: 0 none, 1 128B_BASE32B, 2 128B
}
__host__ __device__ constexpr uint32_t instr_desc(int M, int Nn, int a_mn) {   // kind::tf32, fp32 accumulate, B operand K-major
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); }
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                   "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
// three 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld16x3(uint32_t t0, uint32_t t1, uint32_t t2, float (&a)[16], float (&b)[16], float (&c)[16]) {
    uint32_t r[48];
#define JSTSP_LD16(base, T)                                                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                               \
                 : "=r"(r[base + 0]), "=r"(r[base + 1]), "=r"(r[base + 2]), "=r"(r[base + 3]), "=r"(r[base + 4]), "=r"(r[base + 5]), "=r"(r[base + 6]),   \
                   "=r"(r[base + 7]), "=r"(r[base + 8]), "=r"(r[base + 9]), "=r"(r[base + 10]), "=r"(r[base + 11]), "=r"(r[base + 12]), "=r"(r[base + 13]), \
                   "=r"(r[base + 14]), "=r"(r[base + 15])                                                                                                \
                 : "r"(T))
    JSTSP_LD16(0, t0); JSTSP_LD16(16, t1); JSTSP_LD16(32, t2);
#undef JSTSP_LD16
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) { a[j] = __uint_as_float(r[j]); b[j] = __uint_as_float(r[16 + j]); c[j] = __uint_as_float(r[32 + j]); }
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }   // what the tensor core keeps
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }                                        // exact in fp32

// ---- A S -> pass-1 small operand, hi | lo split, in the shared-memory image the kernel bulk-copies per stage ----------------
// asop[b][stage][kg 0..7][row group 0..NG-1][8 rows][4 k]   (k = 2 (p % 16) + c inside stage p / 16)
template <int N>
__global__ void __launch_bounds__(256) k_expand_as(const cx<float>* __restrict__ AS, float* __restrict__ asop, int P) {
    constexpr int NS = 4 * N, NG = NS / 8, OP1F = 8 * NG * 32;
    const int b = blockIdx.y;
    const cx<float>* as = AS + (size_t)b * N * P;
    float* out = asop + (size_t)b * (size_t)(P / 16) * OP1F;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * P; t += gridDim.x * blockDim.x) {
        const int n = t % N, pp = t / N;
        const cx<float> a = as[t];
        float* st = out + (size_t)(pp / 16) * OP1F;
        const int kre = 2 * (pp % 16), kim = kre + 1;
        auto put = [&](int row, int k, float v) {
            st[((k / 4) * NG + row / 8) * 32 + (row % 8) * 4 + (k % 4)] = tf32_hi(v);
            const int rl = row + 2 * N;
            st[((k / 4) * NG + rl / 8) * 32 + (rl % 8) * 4 + (k % 4)] = tf32_lo(v);
        };
        put(2 * n, kre, a.re); put(2 * n, kim, -a.im);          // Xs_re = sum ASr Br - ASi Bi
        put(2 * n + 1, kre, a.im); put(2 * n + 1, kim, a.re);   // Xs_im = sum ASi Br + ASr Bi
    }
}

// ---- the fused iteration kernel ------------------------------------------------------------------------------------------
template <int N, int NST>
__global__ void __launch_bounds__(THREADS, 2) k_fused_tc(AdmmP<float> p, const __grid_constant__ CUtensorMap mapB1, const __grid_constant__ CUtensorMap mapB2,
                                                          const float* __restrict__ asop, int b_shared) {
    static_assert(N == 16, "the worker mapping below is written for N = 16 rows");
    using G = Geo<N, NST>;
    constexpr int NS = G::NS, NG = G::NG, OP1 = G::OP1, KOP_LBO = G::KOP_LBO, NH = N / 2;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* ring = smem;
    unsigned char* opnd = smem + NST * STAGE;
    float* Zre = reinterpret_cast<float*>(opnd + G::OPREG);
    float* Zim = Zre + N * ZP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(opnd + G::OPREG + G::ZREG);
    uint64_t *full = bars, *hi_done = bars + NST, *lo_ready = bars + 2 * NST, *empty = bars + 3 * NST;
    uint64_t *d1_full = bars + 4 * NST, *kop_ready = d1_full + 1, *d2_full = d1_full + 2, *d2_empty = d1_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d1_full + 6);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int b = blockIdx.y, chunk = blockIdx.x, c0 = chunk * MC;
    const int P = p.P, M = p.M;
    const int S1 = p.iter > 0 ? (2 * P) / 32 : 0;            // pass-1 stages (iteration 0: Xs = C = V2 = 0, nothing to compute)
    const int NBLK = (2 * P) / 128, S2 = NBLK * 4, GT = S1 + S2;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&hi_done[s], 1); mbar_init(&lo_ready[s], NWW); mbar_init(&empty[s], 1); }
        mbar_init(d1_full, 1); mbar_init(kop_ready, NWW);
        mbar_init(&d2_full[0], 1); mbar_init(&d2_full[1], 1); mbar_init(&d2_empty[0], NWW); mbar_init(&d2_empty[1], NWW);
        mbar_fence_init();
    }
    if (warp == NWW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)G::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;
    // The tensor core truncates when it adds into the accumulator, so chains are kept short: two alternating pass-1 accumulators
    // (summed in fp32 by the workers) and the small lo x hi term goes to the columns of the other small term (hi x lo).
    const uint32_t D1[2] = {tm, tm + NS}, D2[2] = {tm + 2 * NS, tm + 3 * NS};

    if (warp == NWW) {
        // ===== TMA producer =====
        if (lane == 0) {
            const int bb = b_shared ? 0 : b;
            const float* as_b = asop + (size_t)b * (size_t)(P / 16) * (OP1 / 4);
            const bool dbg5 = p.dbg && p.dbg_kernel == 5;
            long long w_empty = 0, t_begin = dbg5 ? clock64() : 0, tq = 0;
            for (int g = 0; g < GT; ++g) {
                const int slot = g % NST, use = g / NST;
                if (dbg5) tq = clock64();
                if (use > 0) mbar_wait(&empty[slot], (use - 1) & 1);
                if (dbg5) w_empty += clock64() - tq;
                if (g < S1) {
                    mbar_expect_tx(&full[slot], STAGE + OP1);
                    tma_3d(ring + slot * STAGE, &mapB1, 32 * g, c0, bb, &full[slot]);
                    tma_bulk_g2s(opnd + slot * OP1, as_b + (size_t)g * (OP1 / 4), OP1, &full[slot]);
                } else {
                    const int j = g - S1, blk = j / 4, mq = j % 4;
                    mbar_expect_tx(&full[slot], STAGE);
                    tma_4d(ring + slot * STAGE, &mapB2, 0, c0 + 32 * mq, 4 * blk, bb, &full[slot]);
                }
            }
            if (dbg5) { long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8; o[0] = 1; o[1] = clock64() - t_begin; o[2] = w_empty; }
        }
    } else if (warp == NWW + 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t id_hi1 = instr_desc(128, NS, 0), id_lo1 = instr_desc(128, NS / 2, 0);
            constexpr uint32_t id_hi2 = instr_desc(128, NS, 1), id_lo2 = instr_desc(128, NS / 2, 1);
            const uint32_t kop_base = smem_u32(opnd);
            auto issue = [&](int g, bool lo) {
                const int slot = g % NST;
                const uint32_t a_base = smem_u32(ring + slot * STAGE);
                if (g < S1) {
                    const uint32_t b_base = smem_u32(opnd + slot * OP1);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32(D1[g & 1] + (lo ? 2 * N : 0), smem_desc(a_base + ks * 32, 16, 1024, 2), smem_desc(b_base + ks * 2 * NG * 128, NG * 128, 128, 0),
                                  lo ? id_lo1 : id_hi1, lo ? 1u : ((g >= 2 || ks) ? 1u : 0u));
                } else {
                    const int j = g - S1, blk = j / 4, mq = j % 4;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_tf32(D2[blk & 1] + (lo ? 2 * N : 0), smem_desc(a_base + ks * 1024, 4096, 512, 1),
                                  smem_desc(kop_base + (mq * 4 + ks) * 2 * KOP_LBO, KOP_LBO, 128, 0), lo ? id_lo2 : id_hi2, lo ? 1u : ((mq | ks) ? 1u : 0u));
                }
            };
            const bool dbg4 = p.dbg && p.dbg_kernel == 4;
            long long w_full = 0, w_lo = 0, w_kop = 0, w_d2e = 0, w_issue = 0, w_commit = 0, t_begin = dbg4 ? clock64() : 0, tq = 0;
            auto do_lo = [&](int gl) {
                const int slot = gl % NST, use = gl / NST;
                if (dbg4) tq = clock64();
                mbar_wait(&lo_ready[slot], use & 1);
                if (dbg4) { long long t = clock64(); w_lo += t - tq; tq = t; }
                tc_fence_after();
                issue(gl, true);
                if (dbg4) { long long t = clock64(); w_issue += t - tq; tq = t; }
                umma_commit(&empty[slot]);
                if (gl == S1 - 1) umma_commit(d1_full);
                if (gl >= S1 && (gl - S1) % 4 == 3) umma_commit(&d2_full[((gl - S1) / 4) & 1]);
                if (dbg4) w_commit += clock64() - tq;
            };
            int next_lo = 0;
            for (int g = 0; g < GT; ++g) {
                if (g == S1) {                                   // pass boundary: finish pass 1, then wait for the K operand
                    while (next_lo < S1) do_lo(next_lo++);
                    if (dbg4) tq = clock64();
                    mbar_wait(kop_ready, 0);
                    if (dbg4) w_kop += clock64() - tq;
                }
                if (g - next_lo >= LAG) do_lo(next_lo++);
                const int slot = g % NST, use = g / NST;
                if (g >= S1) {
                    const int j = g - S1, blk = j / 4;
                    if (dbg4) tq = clock64();
                    if (j % 4 == 0 && blk >= 2) mbar_wait(&d2_empty[blk & 1], ((blk >> 1) - 1) & 1);
                    if (dbg4) w_d2e += clock64() - tq;
                }
                if (dbg4) tq = clock64();
                mbar_wait(&full[slot], use & 1);
                if (dbg4) { long long t = clock64(); w_full += t - tq; tq = t; }
                tc_fence_after();
                issue(g, false);
                if (dbg4) { long long t = clock64(); w_issue += t - tq; tq = t; }
                umma_commit(&hi_done[slot]);
                if (dbg4) w_commit += clock64() - tq;
            }
            while (next_lo < GT) do_lo(next_lo++);
            if (dbg4) {
                long long* o = p.dbg + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
                o[0] = 1; o[1] = clock64() - t_begin; o[2] = w_full; o[3] = w_lo; o[4] = w_kop; o[5] = w_d2e; o[6] = w_issue; o[7] = w_commit;
            }
        }
    } else {
        // ===== workers =====
        const int quad = warp % 4, half = warp / 4;
        const int m = quad * 32 + lane;                          // column of the chunk (pass 1) / row of the 128-float block (pass 2)
        const int n0 = half * NH;
        const float rho = (float)p.rho[b];
        const float irho = 1.0f / rho, kap = rho / (rho + 1.0f);
        const size_t eoff = (size_t)b * N * M + (size_t)(c0 + m) * N + n0;
        cx<float>* __restrict__ Xg = p.X + eoff;
        cx<float>* __restrict__ V1g = p.V1 + eoff;
        cx<float>* __restrict__ V2g = p.V2 + eoff;
        const cx<float>* __restrict__ SYg = p.subY + (long long)b * p.ld_subY + (size_t)(c0 + m) * N + n0;
        const float* __restrict__ OMg = p.omega + (long long)b * p.ld_omega + (size_t)(c0 + m) * N + n0;
        // warm L2 with the state this thread reads after pass 1
        prefetch_l2(Xg); prefetch_l2(V1g); prefetch_l2(V2g); prefetch_l2(SYg); prefetch_l2(OMg);

        const bool dbg6 = p.dbg && p.dbg_kernel == 6 && tid == 0;
        long long w_hi = 0, t_rw = 0, t_fa = 0, t_epi = 0, tq = 0;
        auto rewrite = [&](int g) {                              // tile <- tile - trunc_tf32(tile), in place
            const int slot = g % NST, use = g / NST;
            if (dbg6) tq = clock64();
            mbar_wait(&hi_done[slot], use & 1);
            if (dbg6) { long long t = clock64(); w_hi += t - tq; tq = t; }
            float4* t4 = reinterpret_cast<float4*>(ring + slot * STAGE);
#pragma unroll
            for (int u = 0; u < STAGE / 16 / WORKERS; ++u) {
                float4 v = t4[tid + u * WORKERS];
                v.x = tf32_lo(v.x); v.y = tf32_lo(v.y); v.z = tf32_lo(v.z); v.w = tf32_lo(v.w);
                t4[tid + u * WORKERS] = v;
            }
            if (dbg6) { long long t = clock64(); t_rw += t - tq; tq = t; }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&lo_ready[slot]);
            if (dbg6) t_fa += clock64() - tq;
        };
        const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;
        JSTSP_STAMP(p, 3, cta_id, 0);
        for (int g = 0; g < S1; ++g) rewrite(g);
        JSTSP_STAMP(p, 3, cta_id, 1);

        // ---- element-wise update of this thread's NH rows of column c0 + m ----
        float xs_r[NH], xs_i[NH], c_r[NH], c_i[NH];
        if (S1 > 0) {
            mbar_wait(d1_full, 0);
            tc_fence_after();
            float a[16], l[16];
            const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
            tmem_ld16(D1[0] + lane_base + 2 * n0, a);
            tmem_ld16(D1[0] + lane_base + 2 * N + 2 * n0, l);
#pragma unroll
            for (int r = 0; r < NH; ++r) { xs_r[r] = a[2 * r]; xs_i[r] = a[2 * r + 1]; c_r[r] = l[2 * r]; c_i[r] = l[2 * r + 1]; }
            if (S1 > 1) {
                tmem_ld16(D1[1] + lane_base + 2 * n0, a);
                tmem_ld16(D1[1] + lane_base + 2 * N + 2 * n0, l);
#pragma unroll
                for (int r = 0; r < NH; ++r) { xs_r[r] += a[2 * r]; xs_i[r] += a[2 * r + 1]; c_r[r] += l[2 * r]; c_i[r] += l[2 * r + 1]; }
            }
#pragma unroll
            for (int r = 0; r < NH; ++r) { xs_r[r] += c_r[r]; xs_i[r] += c_i[r]; }     // big sums first, the two 2^-11-sized correction sums last
        } else {
#pragma unroll
            for (int r = 0; r < NH; ++r) { xs_r[r] = 0.f; xs_i[r] = 0.f; }
        }
        JSTSP_STAMP(p, 3, cta_id, 2);
        cx<float> xo[NH], v1[NH], v2[NH];
        {
            cx<float> t[4];
#pragma unroll
            for (int hf = 0; hf < NH / 4; ++hf) {
                ld4c<float>(Xg + 4 * hf, t);
#pragma unroll
                for (int u = 0; u < 4; ++u) xo[4 * hf + u] = t[u];
                ld4c<float>(V1g + 4 * hf, t);
#pragma unroll
                for (int u = 0; u < 4; ++u) v1[4 * hf + u] = t[u];
                ld4c<float>(V2g + 4 * hf, t);
#pragma unroll
                for (int u = 0; u < 4; ++u) v2[4 * hf + u] = t[u];
            }
        }
#pragma unroll
        for (int r = 0; r < NH; ++r) {
            // C = rho/(rho+1) (X - Xs - V2/rho) ; V2 += rho (C - X + Xs)      (.m:61,65 of the previous iteration; all zero at i = 1)
            c_r[r] = kap * (xo[r].re - xs_r[r] - irho * v2[r].re); c_i[r] = kap * (xo[r].im - xs_i[r] - irho * v2[r].im);
            v2[r].re += rho * (c_r[r] - xo[r].re + xs_r[r]); v2[r].im += rho * (c_i[r] - xo[r].im + xs_i[r]);
            Zre[(n0 + r) * ZP + m] = xo[r].re - irho * v1[r].re; Zim[(n0 + r) * ZP + m] = xo[r].im - irho * v1[r].im;   // SVT input (.m:35)
        }
        worker_sync();
        // Y = W Z  (W = U diag(max(0,1-tau/sigma)) U^H from k_svt_weights)
        float y_r[NH], y_i[NH];
#pragma unroll
        for (int r = 0; r < NH; ++r) { y_r[r] = 0.f; y_i[r] = 0.f; }
        {
            const cx<float>* __restrict__ Wg = p.W + (size_t)b * N * N + n0;
#pragma unroll 4
            for (int k = 0; k < N; ++k) {
                const float zr = Zre[k * ZP + m], zi = Zim[k * ZP + m];
                cx<float> w[NH];
#pragma unroll
                for (int hf = 0; hf < NH / 4; ++hf) {
                    cx<float> t[4];
                    ld4c<float>(Wg + (size_t)N * k + 4 * hf, t);
#pragma unroll
                    for (int u = 0; u < 4; ++u) w[4 * hf + u] = t[u];
                }
#pragma unroll
                for (int r = 0; r < NH; ++r) cmac<float>(y_r[r], y_i[r], w[r].re, w[r].im, zr, zi);
            }
        }
        const bool last = (p.iter == p.imax - 1) && p.Yout != nullptr;
        float kt_r[NH], kt_i[NH], zn_r[NH], zn_i[NH];
        {
            cx<float> sy[NH]; float om[NH];
#pragma unroll
            for (int hf = 0; hf < NH / 4; ++hf) {
                cx<float> t[4]; float o4[4];
                ld4c<float>(SYg + 4 * hf, t); ld4r<float>(OMg + 4 * hf, o4);
#pragma unroll
                for (int u = 0; u < 4; ++u) { sy[4 * hf + u] = t[u]; om[4 * hf + u] = o4[u]; }
            }
#pragma unroll
            for (int r = 0; r < NH; ++r) {
                const float d = 1.0f / (om[r] + 2.0f * rho);                                                                    // iK1 (.m:20)
                const float xr = (v1[r].re + rho * y_r[r] + sy[r].re + v2[r].re + rho * c_r[r] + rho * xs_r[r]) * d;             // .m:38-40
                const float xi = (v1[r].im + rho * y_i[r] + sy[r].im + v2[r].im + rho * c_i[r] + rho * xs_i[r]) * d;
                v1[r].re += rho * (y_r[r] - xr); v1[r].im += rho * (y_i[r] - xi);                                                // .m:64
                kt_r[r] = xr - irho * v2[r].re - c_r[r]; kt_i[r] = xi - irho * v2[r].im - c_i[r];                                // .m:43
                zn_r[r] = xr - irho * v1[r].re; zn_i[r] = xi - irho * v1[r].im;                                                  // next SVT input
                xo[r] = mk<float>(xr, xi);
            }
        }
#pragma unroll
        for (int hf = 0; hf < NH / 4; ++hf) {
            cx<float> t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = xo[4 * hf + u];
            st4c<float>(Xg + 4 * hf, t);
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = v1[4 * hf + u];
            st4c<float>(V1g + 4 * hf, t);
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = v2[4 * hf + u];
            st4c<float>(V2g + 4 * hf, t);
            if (last) {
#pragma unroll
                for (int u = 0; u < 4; ++u) t[u] = mk<float>(y_r[4 * hf + u], y_i[4 * hf + u]);
                st4c<float>(p.Yout + (long long)b * p.ld_Y + (size_t)(c0 + m) * N + n0 + 4 * hf, t);
            }
        }
        // pass-2 small operand: rows (n,c) hi | lo, k = m, K-major SWIZZLE_NONE with a padded k-group stride
        {
            float* kop = reinterpret_cast<float*>(opnd);
            const int kbase = (m / 4) * (KOP_LBO / 4) + (m % 4);
#pragma unroll
            for (int r = 0; r < NH; ++r) {
                const int rr = 2 * (n0 + r);
                const float v[2] = {kt_r[r], kt_i[r]};
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int row = rr + c, rl = row + 2 * N;
                    kop[kbase + (row / 8) * 32 + (row % 8) * 4] = tf32_hi(v[c]);
                    kop[kbase + (rl / 8) * 32 + (rl % 8) * 4] = tf32_lo(v[c]);
                }
            }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(kop_ready);
        JSTSP_STAMP(p, 3, cta_id, 3);
        worker_sync();                                           // every read of Z (the SVT input) is done
#pragma unroll
        for (int r = 0; r < NH; ++r) { Zre[(n0 + r) * ZP + m] = zn_r[r]; Zim[(n0 + r) * ZP + m] = zn_i[r]; }

        // ---- pass 2: lo rewrites and per-block epilogues (T1 partial, row-major [N][P]) ----
        float* __restrict__ T1f = reinterpret_cast<float*>(p.T1 + ((size_t)b * p.nmc + chunk) * (size_t)N * P);
        auto epilogue = [&](int blk) {
            mbar_wait(&d2_full[blk & 1], (blk >> 1) & 1);
            tc_fence_after();
            float a[16], l[16];
            const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
            tmem_ld16(D2[blk & 1] + lane_base + 2 * n0, a);
            tmem_ld16(D2[blk & 1] + lane_base + 2 * N + 2 * n0, l);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[blk & 1]);
            // lane = float index q of the block: even lanes hold B_re rows, odd lanes B_im rows of the same p
            //   T1r = D[(p,re),(n,re)] + D[(p,im),(n,im)] ; T1i = D[(p,re),(n,im)] - D[(p,im),(n,re)]
#pragma unroll
            for (int r = 0; r < NH; ++r) {
                const float vre = a[2 * r] + l[2 * r], vim = a[2 * r + 1] + l[2 * r + 1];
                const float other = __shfl_xor_sync(0xffffffffu, vim, 1);
                const float out = (lane & 1) ? (other - vre) : (vre + other);
                T1f[(size_t)(n0 + r) * 2 * P + 128 * blk + m] = out;
            }
        };
        for (int j = 0; j < S2; ++j) {
            rewrite(S1 + j);
            if (dbg6) tq = clock64();
            if (j % 4 == 0 && j >= 4) epilogue(j / 4 - 1);
            if (dbg6) t_epi += clock64() - tq;
        }
        if (dbg6) { long long* o = p.dbg + (size_t)cta_id * 8; o[0] = 1; o[1] = w_hi; o[2] = t_rw; o[3] = t_fa; o[4] = t_epi; }
        JSTSP_STAMP(p, 3, cta_id, 4);
        epilogue(NBLK - 1);
        JSTSP_STAMP(p, 3, cta_id, 5);

        // ---- partial Gram of the next SVT input (all MMAs have completed: the ring is scratch now) ----
        worker_sync();
        {
            float* scratch = reinterpret_cast<float*>(ring);     // [slice][N*N][2]
            constexpr int NB4 = N / 4, COMBOS = NB4 * NB4, SLICES = WORKERS / COMBOS, CPS = MC / SLICES;
            static_assert(SLICES * N * N * 2 * 4 <= NST * STAGE, "Gram scratch does not fit the ring");
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
            worker_sync();
            double* out = p.gram + ((size_t)b * p.nmc + chunk) * 2 * N * N;
            for (int t = tid; t < N * N; t += WORKERS) {
                double re = 0.0, im = 0.0;
#pragma unroll 4
                for (int s = 0; s < SLICES; ++s) { re += (double)scratch[((size_t)s * N * N + t) * 2]; im += (double)scratch[((size_t)s * N * N + t) * 2 + 1]; }
                out[2 * t] = re; out[2 * t + 1] = im;
            }
        }
        JSTSP_STAMP(p, 3, cta_id, 6);
        if (p.dbg && p.dbg_kernel == 3 && threadIdx.x == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.dbg[(size_t)cta_id * 8 + 7] = sm; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NWW) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)G::TMEM_COLS));
}

// ---- host side: tensor maps over the dictionary ----------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* f = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
    }
    return fn;
}
// B: nB dictionaries of P x M complex fp32 (column-major), ld_B complex elements apart.  Returns false when the driver refuses.
inline bool make_maps(const cx<float>* B, long long ld_B, int nB, int P, int M, CUtensorMap* m1, CUtensorMap* m2) {
    auto enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t row_bytes = (cuuint64_t)2 * P * 4, trial_bytes = (cuuint64_t)(ld_B ? ld_B : (long long)P * M) * 8;
    void* base = const_cast<cx<float>*>(B);
    {
        cuuint64_t dims[3] = {(cuuint64_t)2 * P, (cuuint64_t)M, (cuuint64_t)nB}; cuuint64_t strides[2] = {row_bytes, trial_bytes};
        cuuint32_t box[3] = {32, 128, 1}, es[3] = {1, 1, 1};
        if (enc(m1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    {
        cuuint64_t dims[4] = {32, (cuuint64_t)M, (cuuint64_t)(2 * P / 32), (cuuint64_t)nB}; cuuint64_t strides[3] = {row_bytes, 128, trial_bytes};
        cuuint32_t box[4] = {32, 32, 4, 1}, es[4] = {1, 1, 1, 1};
        if (enc(m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
    }
    return true;
}

}  // namespace tc
}  // namespace jstsp
