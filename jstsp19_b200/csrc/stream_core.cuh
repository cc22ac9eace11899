// stream_core.cuh - TMA-fed streaming contraction, the fast path of every big product.
//
//   Out[r, o] = sum_{c < ncols} L[r, c] * op(Big[o, c])        op = conj or identity
//
// `Big` is column-major with the OUTPUT index o contiguous (B for K B^H, B^T for (A S) B,
// the Hermitian B B^H for V BBH), so the reduction columns of one pipeline stage are
// contiguous segments of global memory: each is moved by one `cp.async.bulk` (TMA bulk
// copy, SASS UBLKCP) into a shared-memory ring and completes on an mbarrier.  The small
// left operand L (<= 64 rows) sits in shared memory in planar re/im form; every thread
// keeps an 8-row x 2-output complex accumulator tile in registers, reads its two ring
// values with ONE conflict-free LDS.128 and the 8 L rows with 4 broadcast LDS.128 per 64
// FFMA.  A stage is always a full group of `stage_cols` columns (the tail re-reads the last
// valid column against zero-padded L columns), so the unrolled stage body is branch-free
// and the compiler software-pipelines the shared-memory loads under the FMAs.  No thread
// ever waits on a global load: HBM/L2 latency is covered by the stages in flight.
#pragma once
#include "common.cuh"
#include "gemm_cores.cuh"

namespace jstsp {

template <typename T> __host__ __device__ constexpr int stage_cols() { return sizeof(T) == 4 ? 8 : 4; }   // 16 KB per 256 outputs
constexpr int kStages = 4;     // default ring depth (barrier array size)
constexpr int kOW = 64;        // outputs per warp (2 per lane)

__host__ __device__ __forceinline__ int round_up_to(int n, int m) { return (n + m - 1) / m * m; }
__host__ __device__ __forceinline__ int cta_width(int NG) { return (kWarps / NG) * kOW; }   // outputs per CTA pass

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// ---- packed fp32x2 FMA (Blackwell FFMA2): two IEEE fp32 FMAs per issue slot ----------------------
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void ffma2(uint64_t& d, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); }

// Register tile of one thread: 8 rows x 2 outputs, complex.  fp32 keeps ROW PAIRS packed in 64-bit
// registers so that one complex MAC of two rows is 4 FFMA2 (the 8x2 tile update of one reduction
// column is 32 FFMA2 instead of 64 FFMA); fp64 uses plain DFMA.  Both round exactly like scalar FMAs.
template <typename T> struct AccTile;
template <> struct AccTile<float> {
    uint64_t R[4][2], I[4][2];
    __device__ __forceinline__ void load(const float (&ar)[kRB][2], const float (&ai)[kRB][2]) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 2; ++j) { R[q][j] = pack2(ar[2 * q][j], ar[2 * q + 1][j]); I[q][j] = pack2(ai[2 * q][j], ai[2 * q + 1][j]); }
    }
    __device__ __forceinline__ void store(float (&ar)[kRB][2], float (&ai)[kRB][2]) const {
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 2; ++j) { unpack2(R[q][j], ar[2 * q][j], ar[2 * q + 1][j]); unpack2(I[q][j], ai[2 * q][j], ai[2 * q + 1][j]); }
    }
    // tile += L(:, col) * [b0 b1]   ;  lre/lim point at this thread's 8 rows of the column
    __device__ __forceinline__ void mac(const float* __restrict__ lre, const float* __restrict__ lim, cx<float> b0, cx<float> b1) {
        const float4 r0 = *reinterpret_cast<const float4*>(lre), r1 = *reinterpret_cast<const float4*>(lre + 4);
        const float4 i0 = *reinterpret_cast<const float4*>(lim), i1 = *reinterpret_cast<const float4*>(lim + 4);
        const uint64_t LR[4] = {pack2(r0.x, r0.y), pack2(r0.z, r0.w), pack2(r1.x, r1.y), pack2(r1.z, r1.w)};
        const uint64_t LI[4] = {pack2(i0.x, i0.y), pack2(i0.z, i0.w), pack2(i1.x, i1.y), pack2(i1.z, i1.w)};
        const uint64_t B0R = pack2(b0.re, b0.re), B0I = pack2(b0.im, b0.im), B0N = pack2(-b0.im, -b0.im);
        const uint64_t B1R = pack2(b1.re, b1.re), B1I = pack2(b1.im, b1.im), B1N = pack2(-b1.im, -b1.im);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ffma2(R[q][0], LR[q], B0R); ffma2(I[q][0], LR[q], B0I);
            ffma2(R[q][1], LR[q], B1R); ffma2(I[q][1], LR[q], B1I);
            ffma2(R[q][0], LI[q], B0N); ffma2(I[q][0], LI[q], B0R);
            ffma2(R[q][1], LI[q], B1N); ffma2(I[q][1], LI[q], B1R);
        }
    }
};
template <> struct AccTile<double> {
    double r[kRB][2], i[kRB][2];
    __device__ __forceinline__ void load(const double (&ar)[kRB][2], const double (&ai)[kRB][2]) {
#pragma unroll
        for (int q = 0; q < kRB; ++q) { r[q][0] = ar[q][0]; r[q][1] = ar[q][1]; i[q][0] = ai[q][0]; i[q][1] = ai[q][1]; }
    }
    __device__ __forceinline__ void store(double (&ar)[kRB][2], double (&ai)[kRB][2]) const {
#pragma unroll
        for (int q = 0; q < kRB; ++q) { ar[q][0] = r[q][0]; ar[q][1] = r[q][1]; ai[q][0] = i[q][0]; ai[q][1] = i[q][1]; }
    }
    __device__ __forceinline__ void mac(const double* __restrict__ lre, const double* __restrict__ lim, cx<double> b0, cx<double> b1) {
        double lr[kRB], li[kRB];
        load_rows8<double>(lre, 0, 0, 0, lr);
        load_rows8<double>(lim, 0, 0, 0, li);
#pragma unroll
        for (int q = 0; q < kRB; ++q) {
            cmac<double>(r[q][0], i[q][0], lr[q], li[q], b0.re, b0.im);
            cmac<double>(r[q][1], i[q][1], lr[q], li[q], b1.re, b1.im);
        }
    }
};

template <typename T>
struct StreamRing {
    __host__ __device__ static size_t bytes(int width, int stages = kStages) { return sizeof(cx<T>) * (size_t)stages * stage_cols<T>() * width; }
};

// output index (within the warp's 64-wide group) of accumulator column j
template <typename T> __device__ __forceinline__ int out_of(int lane, int j) {
    if constexpr (sizeof(T) == 4) return 2 * lane + j;
    else return j * kWarp + lane;
}

// the two ring / exchange values of this lane at row pointer `rowp` (64 outputs wide)
template <typename T> __device__ __forceinline__ void load_pair(const cx<T>* __restrict__ rowp, int lane, cx<T>& b0, cx<T>& b1) {
    if constexpr (sizeof(T) == 4) {
        float4 v = *reinterpret_cast<const float4*>(rowp + 2 * lane);
        b0 = mk<T>(v.x, v.y); b1 = mk<T>(v.z, v.w);
    } else {
        b0 = rowp[lane]; b1 = rowp[kWarp + lane];
    }
}
template <typename T> __device__ __forceinline__ void store_pair(cx<T>* __restrict__ rowp, int lane, T r0, T i0, T r1, T i1) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(rowp + 2 * lane) = make_float4(r0, i0, r1, i1);
    } else {
        rowp[lane] = mk<T>(r0, i0); rowp[kWarp + lane] = mk<T>(r1, i1);
    }
}

// One streaming contraction.  All threads of the CTA must call start() and run().
//   ring   : shared memory, StreamRing<T>::bytes(width, STAGES), 16-byte aligned
//   bars   : STAGES "full" mbarriers followed by STAGES "empty" mbarriers - use PipeBars<STAGES> and
//            pipe_bars_init() once per kernel
//   it0    : running stage counter of this CTA (carried across calls so barrier phases stay consistent)
//   Big    : global pointer at (first output of this CTA's tile, column 0); ld in elements
//   width  : outputs covered by this CTA = (warps / NG) * 64 ; nvalid <= width actually present
//   L planes must hold round_up(ncols, stage_cols) columns, the padding columns zero.
template <int STAGES>
struct __align__(8) PipeBars {
    uint64_t full[STAGES];
    uint64_t empty[STAGES];     // one arrival per warp per use of the slot
};
// Initialise the barriers and zero the ring (so that never-written ring bytes are finite), then make
// the generic-proxy writes visible to the async proxy before the first bulk copy lands.
// (the zero fill is only needed when some reduction range is not a multiple of stage_cols)
template <int STAGES>
__device__ __forceinline__ void pipe_bars_init(PipeBars<STAGES>& pb, void* ring, size_t ring_bytes, bool zero_ring) {
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&pb.full[s], 1); mbar_init(&pb.empty[s], kWarps); }
        mbar_fence_init();
    }
    if (zero_ring) {
        float4* r4 = reinterpret_cast<float4*>(ring);
        for (size_t i = threadIdx.x; i < ring_bytes / 16; i += blockDim.x) r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <typename T, int STAGES = kStages>
struct StreamPipe {
    static constexpr int kSC = stage_cols<T>();
    cx<T>* ring; uint64_t* bars; uint64_t* empty; const cx<T>* big; long long ld; int width, nvalid, ncols, nst; uint32_t it0;
    bool contig;    // the reduction columns are back to back in global memory (ld == nvalid): one bulk copy per stage

    __device__ __forceinline__ void issue(int t) {   // called by ONE thread
        const uint32_t slot = (it0 + t) % STAGES;
        const int c0 = t * kSC;
        const uint32_t seg = (uint32_t)(nvalid * sizeof(cx<T>));
        cx<T>* dst = ring + (size_t)slot * kSC * width;
        if (contig) {
            // width == nvalid == ld: the whole stage is one contiguous block; a short tail stage leaves
            // stale (finite: the ring is zero-initialised) values against zero L columns
            const int nc = (ncols - c0) < kSC ? (ncols - c0) : kSC;
            mbar_expect_tx(&bars[slot], seg * nc);
            tma_bulk_g2s(dst, big + (long long)c0 * ld, seg * nc, &bars[slot]);
        } else {
            mbar_expect_tx(&bars[slot], seg * kSC);
#pragma unroll
            for (int c = 0; c < kSC; ++c) {
                const int col = (c0 + c) < ncols ? (c0 + c) : (ncols - 1);    // tail: any valid column (its L column is zero)
                tma_bulk_g2s(dst + (size_t)c * width, big + (long long)col * ld, seg, &bars[slot]);
            }
        }
    }
    __device__ __forceinline__ void start(cx<T>* ring_, uint64_t* bars_, uint32_t it0_, const cx<T>* big_, long long ld_, int width_, int nvalid_,
                                          int ncols_) {
        ring = ring_; bars = bars_; empty = bars_ + STAGES; it0 = it0_; big = big_; ld = ld_; nvalid = nvalid_; ncols = ncols_;
        contig = (ld_ == (long long)nvalid_);
        width = contig ? nvalid_ : width_;          // ring row pitch
        nst = (ncols + kSC - 1) / kSC;
        if (threadIdx.x == 0) {
            const int pre = nst < STAGES ? nst : STAGES;
            for (int t = 0; t < pre; ++t) issue(t);
        }
    }
    // acc[r][j]: row rg*8+r, output og*64 + out_of(lane, j).  Returns the advanced stage counter.
    template <bool CONJ>
    __device__ __forceinline__ uint32_t run(const T* __restrict__ Lre, const T* __restrict__ Lim, int RP, int NG, T (&ar)[kRB][2], T (&ai)[kRB][2]) {
        const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
        const int rg = warp % NG, og = warp / NG;
        const bool active = og < kWarps / NG;
        AccTile<T> acc;
        acc.load(ar, ai);
        for (int t = 0; t < nst; ++t) {
            const uint32_t g = it0 + t, slot = g % STAGES;
            mbar_wait(&bars[slot], (g / STAGES) & 1u);
            if (active) {
                const cx<T>* st = ring + (size_t)slot * kSC * width + og * kOW;
                const T* lre = Lre + (size_t)t * kSC * RP + rg * kRB;
                const T* lim = Lim + (size_t)t * kSC * RP + rg * kRB;
#pragma unroll
                for (int c = 0; c < kSC; ++c) {
                    cx<T> b0, b1;
                    load_pair<T>(st + (size_t)c * width, lane, b0, b1);
                    if (CONJ) { b0.im = -b0.im; b1.im = -b1.im; }
                    acc.mac(lre + (size_t)c * RP, lim + (size_t)c * RP, b0, b1);
                }
            }
            // release the slot: every warp arrives on the slot's "empty" mbarrier once its reads are done (the proxy fence orders those
            // generic-proxy reads before the async-proxy refill); thread 0 refills a slot one stage late, after waiting for that barrier,
            // so no CTA-wide barrier is needed per stage and warps may run up to STAGES-1 stages apart.
#ifdef JSTSP_PIPE_CTA_SYNC
            // sanitizer build (make racecheck-lib): compute-sanitizer racecheck models CTA barriers but neither mbarriers nor the async proxy, so
            // it reports every slot read against the refill whatever the handshake.  This variant releases the slot with a CTA-wide barrier, which
            // the tool does follow: a clean run of it shows that no OTHER shared-memory hazard hides behind those reports.
            __syncthreads();
            if (threadIdx.x == 0 && t + STAGES < nst) issue(t + STAGES);
            continue;
#endif
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[slot])) : "memory");
            }
            if (threadIdx.x == 0 && t >= 1 && t - 1 + STAGES < nst) {
                const uint32_t gp = it0 + t - 1;
                mbar_wait(&empty[gp % STAGES], (gp / STAGES) & 1u);
                issue(t - 1 + STAGES);
            }
        }
        acc.store(ar, ai);
        __syncthreads();                           // ring and L planes may be reused by the caller
        return it0 + nst;
    }
};

// Small product with BOTH operands in shared memory (the row-mixing steps between the streamed
// products):  acc[r][j] += sum_{k<nk} L[rg*8+r, k] * R2[k][og*64 + out(lane,j)],  R2 row pitch in elements.
template <typename T>
__device__ __forceinline__ void smem_contract(const T* __restrict__ Lre, const T* __restrict__ Lim, int RP, int rg, const cx<T>* __restrict__ R2,
                                              int pitch, int og, int nk, T (&ar)[kRB][2], T (&ai)[kRB][2]) {
    const int lane = threadIdx.x % kWarp;
    const cx<T>* rp = R2 + og * kOW;
    AccTile<T> acc;
    acc.load(ar, ai);
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
        cx<T> b0, b1;
        load_pair<T>(rp + (size_t)k * pitch, lane, b0, b1);
        acc.mac(Lre + (size_t)k * RP + rg * kRB, Lim + (size_t)k * RP + rg * kRB, b0, b1);
    }
    acc.store(ar, ai);
}

template <typename T> __device__ __forceinline__ void load_rows4(const T* __restrict__ p, T (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
        double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
}

// Register-blocked partial Gram matrix of a planar tile z (row pitch RP, n <= 64 rows, ncols
// columns): 4x4 output blocks per thread, the column range split into slices, products in T
// per slice, cross-slice sum in fp64.  `scratch` needs slices*n*n*2*sizeof(T) bytes of shared
// memory; out: n x n interleaved double.  Contains __syncthreads().
template <typename T>
__device__ inline void gram_blocked(const T* __restrict__ zre, const T* __restrict__ zim, int RP, int n, int ncols, T* __restrict__ scratch,
                                    size_t scratch_bytes, double* __restrict__ out) {
    const int nb4 = (n + 3) / 4, combos = nb4 * nb4;
    int slices = kThreads / combos; if (slices < 1) slices = 1;
    const int fit = (int)(scratch_bytes / ((size_t)n * n * 2 * sizeof(T)));
    if (slices > fit) slices = fit;
    if (slices > ncols) slices = ncols;
    if (slices < 1) slices = 1;
    const int cps = (ncols + slices - 1) / slices;
    for (int w = threadIdx.x; w < combos * slices; w += kThreads) {
        const int combo = w % combos, slice = w / combos;
        const int ib = combo % nb4, jb = combo / nb4;
        T ar[4][4] = {}, ai[4][4] = {};
        const int cbeg = slice * cps, cend = (cbeg + cps) < ncols ? (cbeg + cps) : ncols;
        for (int c = cbeg; c < cend; ++c) {
            T xr[4], xi[4], yr[4], yi[4];
            load_rows4<T>(zre + c * RP + ib * 4, xr); load_rows4<T>(zim + c * RP + ib * 4, xi);
            load_rows4<T>(zre + c * RP + jb * 4, yr); load_rows4<T>(zim + c * RP + jb * 4, yi);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) cmac<T>(ar[u][v], ai[u][v], xr[u], xi[u], yr[v], -yi[v]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int i = ib * 4 + u, j = jb * 4 + v;
                if (i < n && j < n) { scratch[((size_t)slice * n * n + i + n * j) * 2] = ar[u][v]; scratch[((size_t)slice * n * n + i + n * j) * 2 + 1] = ai[u][v]; }
            }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n * n; t += kThreads) {
        double re = 0.0, im = 0.0;
        for (int s = 0; s < slices; ++s) { re += (double)scratch[((size_t)s * n * n + t) * 2]; im += (double)scratch[((size_t)s * n * n + t) * 2 + 1]; }
        out[2 * t] = re; out[2 * t + 1] = im;
    }
    __syncthreads();
}

}  // namespace jstsp
