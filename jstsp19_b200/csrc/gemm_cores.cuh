// gemm_cores.cuh - the two SIMT complex-GEMM cores every solver is built from.
//
// Both multiply a SMALL left operand (R <= 64 rows, staged in shared memory in planar
// re/im form so that 8 consecutive rows are one or two broadcast LDS.128) with a BIG
// column-major right operand that is streamed from global memory exactly once per call:
//
//   contract_cols:  Out[r,k] = sum_{c in chunk} L[r,c] * conj(Big[k,c])     (A^H K B^H pattern,
//                   lanes <-> k, Big read with coalesced 8/16-byte loads)   proposed_algorithm.m:47)
//   expand_cols:    Out[r,c] = sum_k L[r,k] * Big[k,c], c in chunk          (A S B pattern,
//                   lanes <-> c, Big tile transposed through shared memory) proposed_algorithm.m:58)
//
// Register blocking: each warp owns one 8-row group x (32*CB) columns, each thread 8 x CB
// complex accumulators; per inner step one thread issues 8*CB complex MACs (4 FMA each)
// against 4 (fp32) / 8 (fp64) broadcast LDS.128 - measured on B200 this keeps the FMA pipe
// >95% busy (profiles/r01_pipe_probe.txt).
#pragma once
#include "common.cuh"

namespace jstsp {

constexpr int kRB = 8;          // rows per warp-level row group
constexpr int kThreads = 256;   // threads per CTA for all GEMM-shaped kernels
constexpr int kWarps = kThreads / kWarp;

__host__ __device__ __forceinline__ int round_up8(int n) { return (n + 7) & ~7; }

template <typename T> struct KTile { static constexpr int value = 32; };
template <> struct KTile<double> { static constexpr int value = 16; };

// Load the 8 planar rows [rg*8, rg*8+8) of column `c` of a planar tile (row pitch RP).
template <typename T>
__device__ __forceinline__ void load_rows8(const T* __restrict__ plane, int RP, int c, int rg, T (&v)[kRB]) {
    const T* p = plane + (size_t)c * RP + rg * kRB;
    if constexpr (sizeof(T) == 4) {
        float4 a = *reinterpret_cast<const float4*>(p);
        float4 b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double2 a = *reinterpret_cast<const double2*>(p + 2 * i);
            v[2 * i] = a.x; v[2 * i + 1] = a.y;
        }
    }
}

// ---------------------------------------------------------------------------------------
// contract_cols: one pass = NKG*32*KB output columns k; accumulates over `ncols` columns of
// the chunk.  Lre/Lim: planar [ncols][RP].  Big points at (row 0, first column of chunk).
// epi(r, k, re, im) is called for every valid output element.
// ---------------------------------------------------------------------------------------
template <typename T, int KB, typename Epi>
__device__ __forceinline__ void contract_cols(const T* __restrict__ Lre, const T* __restrict__ Lim, int RP, int NG,
                                              int ncols, const cx<T>* __restrict__ Big, long long ldb, int K,
                                              int R, Epi epi) {
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int rg = warp % NG, kg = warp / NG, NKG = kWarps / NG;
    if (kg >= NKG) return;
    const int span = NKG * kWarp * KB;
    for (int k0 = 0; k0 < K; k0 += span) {
        T ar[kRB][KB], ai[kRB][KB];
#pragma unroll
        for (int r = 0; r < kRB; ++r)
#pragma unroll
            for (int j = 0; j < KB; ++j) { ar[r][j] = T(0); ai[r][j] = T(0); }
        int kk[KB];
        const cx<T>* bp[KB];
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            kk[j] = k0 + kg * kWarp * KB + j * kWarp + lane;
            bp[j] = Big + (kk[j] < K ? kk[j] : 0);
        }
#pragma unroll 4
        for (int c = 0; c < ncols; ++c) {
            cx<T> b[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j) b[j] = bp[j][(long long)c * ldb];
            T lr[kRB], li[kRB];
            load_rows8<T>(Lre, RP, c, rg, lr);
            load_rows8<T>(Lim, RP, c, rg, li);
#pragma unroll
            for (int r = 0; r < kRB; ++r)
#pragma unroll
                for (int j = 0; j < KB; ++j) cmac<T>(ar[r][j], ai[r][j], lr[r], li[r], b[j].re, -b[j].im);
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) {
            if (kk[j] < K) {
#pragma unroll
                for (int r = 0; r < kRB; ++r) {
                    int row = rg * kRB + r;
                    if (row < R) epi(row, kk[j], ar[r][j], ai[r][j]);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// expand_cols: chunk of CC = NCG*32*CB columns, contraction over K in tiles of KT.
//   Lg   : global left operand, R x K column-major (ld = R) - or fetched through `lfetch(r,k)`
//   Big  : global, column-major (ldb), pointing at (row 0, first column of chunk)
//   smem : Bs   [CC][KT+1] complex  +  Lre/Lim [KT][RP]
// Result left in acc registers: acc?[r][j] is row rg*8+r, column cg*32*CB + j*32 + lane.
// ---------------------------------------------------------------------------------------
template <typename T, int CB>
struct ExpandSmem {
    static constexpr int KT = KTile<T>::value;
    __host__ __device__ static int ncg(int NG) { int n = kWarps / NG; return n > 4 ? 4 : (n < 1 ? 1 : n); }
    __host__ __device__ static int chunk_cols(int NG) { return ncg(NG) * kWarp * CB; }
    __host__ __device__ static size_t bytes(int NG, int RP) {
        return sizeof(cx<T>) * (size_t)chunk_cols(NG) * (KT + 1) + 2 * sizeof(T) * (size_t)KT * RP;
    }
};

template <typename T, int CB, typename LFetch>
__device__ __forceinline__ void expand_cols(void* smem_raw, int RP, int NG, int R, int K, LFetch lfetch,
                                            const cx<T>* __restrict__ Big, long long ldb, int ncols_valid,
                                            T (&ar)[kRB][CB], T (&ai)[kRB][CB]) {
    constexpr int KT = KTile<T>::value;
    const int NCG = ExpandSmem<T, CB>::ncg(NG);
    const int CC = NCG * kWarp * CB;
    cx<T>* Bs = reinterpret_cast<cx<T>*>(smem_raw);
    T* Lre = reinterpret_cast<T*>(Bs + (size_t)CC * (KT + 1));
    T* Lim = Lre + (size_t)KT * RP;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int rg = warp % NG, cg = warp / NG;
    const bool active = cg < NCG;
#pragma unroll
    for (int r = 0; r < kRB; ++r)
#pragma unroll
        for (int j = 0; j < CB; ++j) { ar[r][j] = T(0); ai[r][j] = T(0); }
    for (int k0 = 0; k0 < K; k0 += KT) {
        const int kt = (K - k0) < KT ? (K - k0) : KT;
        __syncthreads();   // previous tile fully consumed
        // stage Big tile: consecutive threads read consecutive k of one column (coalesced)
        for (int idx = threadIdx.x; idx < CC * KT; idx += kThreads) {
            int col = idx / KT, k = idx % KT;
            cx<T> v = mk<T>(T(0), T(0));
            if (col < ncols_valid && k < kt) v = Big[(long long)col * ldb + k0 + k];
            Bs[(size_t)col * (KT + 1) + k] = v;
        }
        // stage L tile (planar), zero padded rows / k
        for (int idx = threadIdx.x; idx < KT * RP; idx += kThreads) {
            int r = idx % RP, k = idx / RP;
            cx<T> v = mk<T>(T(0), T(0));
            if (r < R && k < kt) v = lfetch(r, k0 + k);
            Lre[idx] = v.re; Lim[idx] = v.im;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int k = 0; k < KT; ++k) {
                cx<T> b[CB];
#pragma unroll
                for (int j = 0; j < CB; ++j) b[j] = Bs[(size_t)(cg * kWarp * CB + j * kWarp + lane) * (KT + 1) + k];
                T lr[kRB], li[kRB];
                load_rows8<T>(Lre, RP, k, rg, lr);
                load_rows8<T>(Lim, RP, k, rg, li);
#pragma unroll
                for (int r = 0; r < kRB; ++r)
#pragma unroll
                    for (int j = 0; j < CB; ++j) cmac<T>(ar[r][j], ai[r][j], lr[r], li[r], b[j].re, b[j].im);
            }
        }
    }
}

// Partial Gram matrix  out[i,j] = sum_{c<ncols} z[i,c] conj(z[j,c])  of a planar tile (row
// pitch RP): products accumulate in T over 16-column segments, segments in fp64, so the
// fp32 path loses nothing beyond the rounding of its inputs.  out: n x n interleaved double.
template <typename T>
__device__ __forceinline__ void gram_partial(const T* __restrict__ zre, const T* __restrict__ zim, int RP, int n, int ncols,
                                             double* __restrict__ out) {
    // 2 x 2 outputs per thread (rows beyond n inside the RP-padded tile are zero); per output the summation order is unchanged:
    // 16-column partial sums in T, accumulated in double
    const int h = (n + 1) / 2;
    for (int t = threadIdx.x; t < h * h; t += blockDim.x) {
        const int i0 = 2 * (t % h), j0 = 2 * (t / h);
        const int i1 = i0 + 1 < n ? i0 + 1 : i0, j1 = j0 + 1 < n ? j0 + 1 : j0;
        double dre[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, dim[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            T re[2][2] = {{0, 0}, {0, 0}}, im[2][2] = {{0, 0}, {0, 0}};
            int ce = c0 + 16 < ncols ? c0 + 16 : ncols;
            for (int c = c0; c < ce; ++c) {
                const T x0r = zre[c * RP + i0], x0i = zim[c * RP + i0], x1r = zre[c * RP + i1], x1i = zim[c * RP + i1];
                const T y0r = zre[c * RP + j0], y0i = zim[c * RP + j0], y1r = zre[c * RP + j1], y1i = zim[c * RP + j1];
                cmac<T>(re[0][0], im[0][0], x0r, x0i, y0r, -y0i); cmac<T>(re[1][0], im[1][0], x1r, x1i, y0r, -y0i);
                cmac<T>(re[0][1], im[0][1], x0r, x0i, y1r, -y1i); cmac<T>(re[1][1], im[1][1], x1r, x1i, y1r, -y1i);
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int bq = 0; bq < 2; ++bq) { dre[a][bq] += (double)re[a][bq]; dim[a][bq] += (double)im[a][bq]; }
        }
        out[2 * (i0 + n * j0)] = dre[0][0]; out[2 * (i0 + n * j0) + 1] = dim[0][0];
        if (i1 != i0) { out[2 * (i1 + n * j0)] = dre[1][0]; out[2 * (i1 + n * j0) + 1] = dim[1][0]; }
        if (j1 != j0) {
            out[2 * (i0 + n * j1)] = dre[0][1]; out[2 * (i0 + n * j1) + 1] = dim[0][1];
            if (i1 != i0) { out[2 * (i1 + n * j1)] = dre[1][1]; out[2 * (i1 + n * j1) + 1] = dim[1][1]; }
        }
    }
}


}  // namespace jstsp
