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
//   bars   : STAGES mbarriers (initialised once per kernel with mbar_init(.,1) + fence + sync)
//   it0    : running stage counter of this CTA (carried across calls so barrier phases stay consistent)
//   Big    : global pointer at (first output of this CTA's tile, column 0); ld in elements
//   width  : outputs covered by this CTA = (warps / NG) * 64 ; nvalid <= width actually present
//   L planes must hold round_up(ncols, stage_cols) columns, the padding columns zero.
template <typename T, int STAGES = kStages>
struct StreamPipe {
    static constexpr int kSC = stage_cols<T>();
    cx<T>* ring; uint64_t* bars; const cx<T>* big; long long ld; int width, nvalid, ncols, nst; uint32_t it0;

    __device__ __forceinline__ void issue(int t) {   // called by ONE thread
        const uint32_t slot = (it0 + t) % STAGES;
        const int c0 = t * kSC;
        const uint32_t seg = (uint32_t)(nvalid * sizeof(cx<T>));
        mbar_expect_tx(&bars[slot], seg * kSC);
        cx<T>* dst = ring + (size_t)slot * kSC * width;
#pragma unroll
        for (int c = 0; c < kSC; ++c) {
            const int col = (c0 + c) < ncols ? (c0 + c) : (ncols - 1);    // tail: any valid column (its L column is zero)
            tma_bulk_g2s(dst + (size_t)c * width, big + (long long)col * ld, seg, &bars[slot]);
        }
    }
    __device__ __forceinline__ void start(cx<T>* ring_, uint64_t* bars_, uint32_t it0_, const cx<T>* big_, long long ld_, int width_, int nvalid_,
                                          int ncols_) {
        ring = ring_; bars = bars_; it0 = it0_; big = big_; ld = ld_; width = width_; nvalid = nvalid_; ncols = ncols_;
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
                    T lr[kRB], li[kRB];
                    load_rows8<T>(lre, RP, c, 0, lr);
                    load_rows8<T>(lim, RP, c, 0, li);
                    const T b0i = CONJ ? -b0.im : b0.im, b1i = CONJ ? -b1.im : b1.im;
#pragma unroll
                    for (int r = 0; r < kRB; ++r) {
                        cmac<T>(ar[r][0], ai[r][0], lr[r], li[r], b0.re, b0i);
                        cmac<T>(ar[r][1], ai[r][1], lr[r], li[r], b1.re, b1i);
                    }
                }
            }
            __syncthreads();                       // every warp is done with this ring slot
            if (threadIdx.x == 0 && t + STAGES < nst) issue(t + STAGES);
        }
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
#pragma unroll 4
    for (int k = 0; k < nk; ++k) {
        cx<T> b0, b1;
        load_pair<T>(rp + (size_t)k * pitch, lane, b0, b1);
        T lr[kRB], li[kRB];
        load_rows8<T>(Lre, RP, k, rg, lr);
        load_rows8<T>(Lim, RP, k, rg, li);
#pragma unroll
        for (int r = 0; r < kRB; ++r) {
            cmac<T>(ar[r][0], ai[r][0], lr[r], li[r], b0.re, b0.im);
            cmac<T>(ar[r][1], ai[r][1], lr[r], li[r], b1.re, b1.im);
        }
    }
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
