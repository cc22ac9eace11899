// admm_fast.cuh - TMA-pipelined fast path of the proposed ADMM iteration ('approximate').
// Included by admm.cu after AdmmP<T>.  Three kernels per iteration:
//   k_xupd_t1_fast : SVT apply + X / V1 update + T1 = K B^H + next Gram        (grid nmc x batch)
//   k_vstep_fast   : Res = A^H T1 - AHA V BBH ; alpha ; V ; S ; A S           (one CTA per trial)
//   k_xs_fast      : Xs = (A S) B through the transposed copy B^T ; C ; V2    (grid nxc x batch)
// Every product - streamed (stream_core.cuh: StreamPipe) or shared-memory resident
// (smem_contract) - uses the same register tile: 8 rows x 2 outputs per thread.
// Preconditions checked on the host (otherwise the generic kernels in admm.cu run):
//   N % 8 == 0, G % 8 == 0, N, G <= 64, 16-byte aligned operands / segments (P, M even in
//   fp32), and for k_vstep_fast P <= outputs of one CTA pass.
#pragma once
#include "stream_core.cuh"

namespace jstsp {

// phase timestamp (developer hook, see jstsp_debug_buffer)
#define JSTSP_STAMP(p, kid, cta, slot)                                                              \
    do {                                                                                            \
        if ((p).dbg && (p).dbg_kernel == (kid) && threadIdx.x == 0) (p).dbg[(size_t)(cta) * 8 + (slot)] = clock64(); \
    } while (0)

template <typename T> struct FastCfg {
    static constexpr int MC = sizeof(T) == 4 ? 128 : 64;      // X-update column chunk
};
constexpr int kXupdStages = 3;   // 3-deep ring keeps two CTAs of k_xupd_t1_fast resident per SM

// 4 consecutive complex values, 16-byte aligned
template <typename T> __device__ __forceinline__ void ld4c(const cx<T>* __restrict__ p, cx<T> (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 2);
        v[0] = mk<T>(a.x, a.y); v[1] = mk<T>(a.z, a.w); v[2] = mk<T>(b.x, b.y); v[3] = mk<T>(b.z, b.w);
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = p[u];
    }
}
template <typename T> __device__ __forceinline__ void st4c(cx<T>* __restrict__ p, const cx<T> (&v)[4]) {
    if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0].re, v[0].im, v[1].re, v[1].im);
        *reinterpret_cast<float4*>(p + 2) = make_float4(v[2].re, v[2].im, v[3].re, v[3].im);
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = v[u];
    }
}
template <typename T> __device__ __forceinline__ void ld4r(const T* __restrict__ p, T (&v)[4]) { load_rows4<T>(p, v); }
// 2 consecutive complex values (16-byte aligned in fp32, two 16-byte accesses in fp64)
template <typename T> __device__ __forceinline__ void ld2c(const cx<T>* __restrict__ p, cx<T> (&v)[2]) {
    if constexpr (sizeof(T) == 4) { float4 a = *reinterpret_cast<const float4*>(p); v[0] = mk<T>(a.x, a.y); v[1] = mk<T>(a.z, a.w); }
    else { v[0] = p[0]; v[1] = p[1]; }
}
template <typename T> __device__ __forceinline__ void st2c(cx<T>* __restrict__ p, const cx<T> (&v)[2]) {
    if constexpr (sizeof(T) == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0].re, v[0].im, v[1].re, v[1].im);
    else { p[0] = v[0]; p[1] = v[1]; }
}
template <typename T> __device__ __forceinline__ void ld2r(const T* __restrict__ p, T (&v)[2]) {
    if constexpr (sizeof(T) == 4) { float2 a = *reinterpret_cast<const float2*>(p); v[0] = a.x; v[1] = a.y; }
    else { double2 a = *reinterpret_cast<const double2*>(p); v[0] = a.x; v[1] = a.y; }
}

// Tiled B^T copy: Bt[(chunk*P + p)*W + (m - chunk*W)] = B[p + P*m], chunk = m / W; columns beyond M are zero.
// (grid ceil(P/32) x ceil(Mpad/32) x nB, Mpad = ceil(M/W)*W)
template <typename T>
__global__ void __launch_bounds__(256) k_transpose_b(const cx<T>* __restrict__ B, long long ld_B, cx<T>* __restrict__ Bt, long long ld_Bt, int P, int M, int W) {
    __shared__ cx<T> tile[32][33];
    const cx<T>* src = B + (long long)blockIdx.z * ld_B;
    cx<T>* dst = Bt + (long long)blockIdx.z * ld_Bt;
    const int p0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
    for (int i = ty; i < 32; i += 8) {
        cx<T> v = mk<T>(T(0), T(0));
        if (p0 + tx < P && m0 + i < M) v = src[(p0 + tx) + (long long)P * (m0 + i)];
        tile[i][tx] = v;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int m = m0 + tx, pp = p0 + i;
        if (pp < P) dst[((long long)(m / W) * P + pp) * W + (m % W)] = tile[tx][i];
    }
}

// =======================================================================================
// kernel 1
// =======================================================================================
template <typename T>
struct XupdFastSmem {
    static constexpr int MC = FastCfg<T>::MC;
    static size_t bytes(int N, int NG, bool conv) {      // N % 8 == 0 -> RP == N
        size_t ring = StreamRing<T>::bytes(cta_width(NG), kXupdStages);
        size_t zt = 2 * sizeof(T) * (size_t)N * (MC + 1);
        size_t planes = (size_t)(conv ? 8 : 4) * N * MC * sizeof(T);
        size_t w = 2 * sizeof(T) * (size_t)N * N;
        return ring + ((zt + 15) & ~size_t(15)) + planes + w;
    }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 2 : 1) k_xupd_t1_fast(AdmmP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ PipeBars<kXupdStages> pb;
    uint64_t* bars = pb.full;
    constexpr int MC = FastCfg<T>::MC, ZP = MC + 1;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int N = p.N, RP = p.RP, NG = p.NG, P = p.P;      // RP == N here
    const int W = cta_width(NG);
    const int c0 = chunk * MC;
    const int ncols = (p.M - c0) < MC ? (p.M - c0) : MC;
    const bool conv = p.convd != nullptr;
    const size_t ringb = StreamRing<T>::bytes(W, kXupdStages);
    cx<T>* ring = reinterpret_cast<cx<T>*>(smem);
    T* Ztre = reinterpret_cast<T*>(smem + ringb);                       // [N][ZP]  Z transposed, padded pitch
    T* Ztim = Ztre + (size_t)N * ZP;
    T* Kre = reinterpret_cast<T*>(smem + ringb + ((2 * sizeof(T) * (size_t)N * ZP + 15) & ~size_t(15)));   // [MC][RP]
    T* Kim = Kre + (size_t)RP * MC;
    T* Nre = Kim + (size_t)RP * MC;                                      // next SVT input
    T* Nim = Nre + (size_t)RP * MC;
    T* Cx = Nim + (size_t)RP * MC;                                       // conv: X re/im, V1 re/im planes
    T* Wre = Nim + (size_t)RP * MC * (conv ? 5 : 1);                     // [N][RP]: Wre[k*RP + r] = W[r,k]
    T* Wim = Wre + (size_t)N * RP;
    const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;
    JSTSP_STAMP(p, 0, cta_id, 0);
    pipe_bars_init(pb, smem, ringb, (ncols % stage_cols<T>()) != 0);
    JSTSP_STAMP(p, 0, cta_id, 1);
    // start streaming B(:, chunk) right away; the element-wise prologue below hides the first latency
    const cx<T>* Bc = p.B + (long long)b * p.ld_B + (long long)c0 * P;
    StreamPipe<T, kXupdStages> pipe;
    uint32_t it = 0;
    pipe.start(ring, bars, it, Bc, (long long)P, W, P < W ? P : W, ncols);

    const T rho = (T)p.rho[b];
    const T irho = T(1) / rho;
    const size_t off = (size_t)b * N * p.M + (size_t)c0 * N;
    cx<T>* __restrict__ X = p.X + off; cx<T>* __restrict__ V1 = p.V1 + off;
    const cx<T>* __restrict__ V2 = p.V2 + off; const cx<T>* __restrict__ C = p.C + off; const cx<T>* __restrict__ Xs = p.Xs + off;
    const cx<T>* __restrict__ subY = p.subY + (long long)b * p.ld_subY + (size_t)c0 * N;
    const T* __restrict__ om = p.omega + (long long)b * p.ld_omega + (size_t)c0 * N;
    const cx<T>* Wg = p.W + (size_t)b * N * N;
    {   // warm L2 with the tiles the element-wise phase reads ~20k cycles from now
        const size_t tile_bytes = sizeof(cx<T>) * (size_t)N * ncols;
        for (size_t o = (size_t)threadIdx.x * 128; o < tile_bytes; o += (size_t)kThreads * 128) {
            prefetch_l2(reinterpret_cast<const char*>(V2) + o); prefetch_l2(reinterpret_cast<const char*>(C) + o);
            prefetch_l2(reinterpret_cast<const char*>(Xs) + o); prefetch_l2(reinterpret_cast<const char*>(subY) + o);
            if (o < tile_bytes / 2) prefetch_l2(reinterpret_cast<const char*>(om) + o);
        }
    }
    for (int t = threadIdx.x; t < N * N; t += kThreads) { cx<T> w = Wg[t]; Wre[t] = w.re; Wim[t] = w.im; }   // (r + N*k) == k*RP + r
    const int nel = N * ncols;                 // elements of this chunk; global index e = c*N + r (N == RP)
    {   // Z = X - V1/rho, transposed into [k][c]: coalesced 16-byte global reads, padded pitch -> conflict-free
        for (int e0 = 0; e0 < N * MC; e0 += 2 * kThreads * 4) {
            cx<T> x[4][2], v[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int e = e0 + 2 * (u * kThreads + threadIdx.x); if (e < nel) { ld2c<T>(X + e, x[u]); ld2c<T>(V1 + e, v[u]); } }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = e0 + 2 * (u * kThreads + threadIdx.x);
                if (e < N * MC) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int r = (e + q) % N, c = (e + q) / N;
                        T zr = 0, zi = 0;
                        if (e < nel) { zr = x[u][q].re - irho * v[u][q].re; zi = x[u][q].im - irho * v[u][q].im; }
                        Ztre[r * ZP + c] = zr; Ztim[r * ZP + c] = zi;
                    }
                }
            }
        }
        if (ncols < MC) {    // padding columns of the streamed operand / Gram planes must be zero
            for (int t = threadIdx.x + ncols * RP; t < MC * RP; t += kThreads) { Kre[t] = 0; Kim[t] = 0; Nre[t] = 0; Nim[t] = 0; }
            if (conv) for (int t = threadIdx.x + ncols * RP; t < MC * RP; t += kThreads) { Cx[t] = 0; Cx[t + RP * MC] = 0; Cx[t + 2 * RP * MC] = 0; Cx[t + 3 * RP * MC] = 0; }
        }
    }
    __syncthreads();
    JSTSP_STAMP(p, 0, cta_id, 2);
    // Y = W Z, register-blocked (8 rows x 1 column per thread); parked in the N planes        (svt.m via SURVEY A.2)
    for (int item = threadIdx.x; item < NG * MC; item += kThreads) {
        const int c = item % MC, rg = item / MC;
        if (c >= ncols) continue;
        T yr[kRB], yi[kRB];
#pragma unroll
        for (int r = 0; r < kRB; ++r) { yr[r] = 0; yi[r] = 0; }
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
            const T zr = Ztre[k * ZP + c], zi = Ztim[k * ZP + c];
            T wr[kRB], wi[kRB];
            load_rows8<T>(Wre, RP, k, rg, wr);
            load_rows8<T>(Wim, RP, k, rg, wi);
#pragma unroll
            for (int r = 0; r < kRB; ++r) cmac<T>(yr[r], yi[r], wr[r], wi[r], zr, zi);
        }
#pragma unroll
        for (int r = 0; r < kRB; ++r) { Nre[c * RP + rg * kRB + r] = yr[r]; Nim[c * RP + rg * kRB + r] = yi[r]; }
    }
    __syncthreads();
    // element-wise updates with fully coalesced 16-byte global accesses          (proposed_algorithm.m:38-43,64)
    const bool last = (p.iter == p.imax - 1) && p.Yout != nullptr;
    for (int e0 = 0; e0 < nel; e0 += 2 * kThreads * 2) {
        cx<T> v1[2][2], v2[2][2], cc[2][2], xs[2][2], sy[2][2]; T omv[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = e0 + 2 * (u * kThreads + threadIdx.x);
            if (e < nel) { ld2c<T>(V1 + e, v1[u]); ld2c<T>(V2 + e, v2[u]); ld2c<T>(C + e, cc[u]); ld2c<T>(Xs + e, xs[u]); ld2c<T>(subY + e, sy[u]); ld2r<T>(om + e, omv[u]); }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = e0 + 2 * (u * kThreads + threadIdx.x);
            if (e >= nel) continue;
            cx<T> xo[2], n1[2], yo[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int si = e + q;                                  // == c*RP + r
                const T y_r = Nre[si], y_i = Nim[si];
                const T d = T(1) / (omv[u][q] + T(2) * rho);                                                             // iK1 (.m:20)
                const T xr = (v1[u][q].re + rho * y_r + sy[u][q].re + v2[u][q].re + rho * cc[u][q].re + rho * xs[u][q].re) * d;   // .m:38-40
                const T xi = (v1[u][q].im + rho * y_i + sy[u][q].im + v2[u][q].im + rho * cc[u][q].im + rho * xs[u][q].im) * d;
                const T n1r = v1[u][q].re + rho * (y_r - xr), n1i = v1[u][q].im + rho * (y_i - xi);                      // .m:64
                xo[q] = mk<T>(xr, xi); n1[q] = mk<T>(n1r, n1i); yo[q] = mk<T>(y_r, y_i);
                Kre[si] = xr - irho * v2[u][q].re - cc[u][q].re; Kim[si] = xi - irho * v2[u][q].im - cc[u][q].im;        // .m:43
                Nre[si] = xr - irho * n1r; Nim[si] = xi - irho * n1i;                                                    // next SVT input (.m:35)
                if (conv) { Cx[si] = xr; Cx[RP * MC + si] = xi; Cx[2 * RP * MC + si] = n1r; Cx[3 * RP * MC + si] = n1i; }
            }
            st2c<T>(X + e, xo); st2c<T>(V1 + e, n1);
            if (last) { cx<T>* yp = p.Yout + (long long)b * p.ld_Y + (size_t)c0 * N + e; yp[0] = yo[0]; yp[1] = yo[1]; }
        }
    }
    __syncthreads();
    JSTSP_STAMP(p, 0, cta_id, 3);
    // T1 partial = K(:,chunk) B(:,chunk)^H, stored row-major [N][P]                (.m:47)
    cx<T>* T1 = p.T1 + ((size_t)b * p.nmc + chunk) * (size_t)N * P;
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int rg = warp % NG, og = warp / NG;
    for (int o0 = 0; o0 < P; o0 += W) {
        if (o0 > 0) pipe.start(ring, bars, it, Bc + o0, (long long)P, W, (P - o0) < W ? (P - o0) : W, ncols);
        T ar[kRB][2], ai[kRB][2];
#pragma unroll
        for (int r = 0; r < kRB; ++r) { ar[r][0] = ar[r][1] = ai[r][0] = ai[r][1] = T(0); }
        it = pipe.template run<true>(Kre, Kim, RP, NG, ar, ai);
        if (og < kWarps / NG) {
            const int ob = o0 + og * kOW;
            if (ob + out_of<T>(lane, 0) < P) {
#pragma unroll
                for (int r = 0; r < kRB; ++r) {
                    cx<T>* rowp = T1 + (size_t)(rg * kRB + r) * P + ob;
                    if (ob + out_of<T>(lane, 1) < P) store_pair<T>(rowp, lane, ar[r][0], ai[r][0], ar[r][1], ai[r][1]);
                    else rowp[out_of<T>(lane, 0)] = mk<T>(ar[r][0], ai[r][0]);
                }
            }
        }
    }
    JSTSP_STAMP(p, 0, cta_id, 4);
    // partial Gram of the next SVT input (the ring is idle now and serves as scratch)
    T* scratch = reinterpret_cast<T*>(smem);
    gram_blocked<T>(Nre, Nim, RP, N, ncols, scratch, ringb, p.gram + ((size_t)b * p.nmc + chunk) * 2 * N * N);
    JSTSP_STAMP(p, 0, cta_id, 5);
    if (p.dbg && p.dbg_kernel == 0 && threadIdx.x == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.dbg[(size_t)cta_id * 8 + 7] = sm; }
    if (conv) {
        size_t cg = (size_t)p.nmc * 2 * N * N;
        double* base = p.cgramA + (size_t)b * 2 * cg + (size_t)chunk * 2 * N * N;
        gram_blocked<T>(Cx + 2 * (size_t)RP * MC, Cx + 3 * (size_t)RP * MC, RP, N, ncols, scratch, ringb, base);          // V1
        gram_blocked<T>(Cx, Cx + (size_t)RP * MC, RP, N, ncols, scratch, ringb, base + cg);                              // X
    }
}

// =======================================================================================
// kernel 2 (one CTA per trial)
// =======================================================================================
template <typename T>
struct VstepFastSmem {
    __host__ __device__ static size_t region0(int N, int G, int GNG, int P) {
        size_t ring = StreamRing<T>::bytes(cta_width(GNG));
        size_t exch = sizeof(cx<T>) * (size_t)(N + G) * round_up_to(P, kOW);
        return ring > exch ? ring : exch;
    }
    __host__ __device__ static size_t bytes(int N, int G, int GNG, int P) {
        size_t lv = 2 * sizeof(T) * (size_t)G * round_up_to(P, stage_cols<T>());
        size_t small = 2 * sizeof(T) * ((size_t)(N + G) * G + (size_t)G * G + (size_t)G * N);
        return region0(N, G, GNG, P) + lv + small;
    }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 2 : 1) k_vstep_fast(AdmmP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ PipeBars<kStages> pb;
    uint64_t* bars = pb.full;
    __shared__ double red[kWarps][3];
    __shared__ double s_alpha;
    const int b = blockIdx.x;
    const int N = p.N, G = p.G, P = p.P, RG = p.GP8, NGg = p.GNG, RN = p.RP, NGn = p.NG;     // RG == G, RN == N
    const int W = cta_width(NGg);
    const int pitch = round_up_to(P, kOW);
    const int Pc = round_up_to(P, stage_cols<T>());
    cx<T>* ring = reinterpret_cast<cx<T>*>(smem);
    cx<T>* E = reinterpret_cast<cx<T>*>(smem);                         // exchange rows [N+G][pitch], aliases the ring
    T* Lre = reinterpret_cast<T*>(smem + VstepFastSmem<T>::region0(N, G, NGg, P));   // planar [Pc][RG]: V, then Res
    T* Lim = Lre + (size_t)RG * Pc;
    T* A1re = Lim + (size_t)RG * Pc;                                   // [A^H | -AHA] : [(N+G)][RG]
    T* A1im = A1re + (size_t)(N + G) * RG;
    T* Qre = A1im + (size_t)(N + G) * RG;                              // AHA : [G][RG]
    T* Qim = Qre + (size_t)G * RG;
    T* Sre = Qim + (size_t)G * RG;                                     // A : [G][RN]
    T* Sim = Sre + (size_t)G * RN;
    pipe_bars_init(pb, smem, StreamRing<T>::bytes(W), (P % stage_cols<T>()) != 0);
    const cx<T>* BBH = p.BBH + (long long)b * p.ld_BBH;
    StreamPipe<T> pipe;
    uint32_t it = 0;
    cx<T>* __restrict__ V = p.V + (size_t)b * G * P;
    cx<T>* __restrict__ VBg = p.VB + (size_t)b * G * P;              // V BBH, row-major [G][P], carried across iterations
    const cx<T>* A = p.A + (long long)b * p.ld_A;
    const cx<T>* AHA = p.AHA + (long long)b * p.ld_AHA;
    {   // BBH is streamed once the exchange rows (which alias the ring) are consumed: warm L2 meanwhile
        const size_t bytes = sizeof(cx<T>) * (size_t)P * P;
        for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)kThreads * 128) prefetch_l2(reinterpret_cast<const char*>(BBH) + o);
    }
    for (int t = threadIdx.x; t < N * G; t += kThreads) {
        const int n = t % N, g = t / N;
        const cx<T> a = A[t];
        A1re[n * RG + g] = a.re; A1im[n * RG + g] = -a.im;             // (A^H)[g,n] = conj(A[n,g])
        Sre[g * RN + n] = a.re; Sim[g * RN + n] = a.im;
    }
    for (int t = threadIdx.x; t < G * G; t += kThreads) {
        const int g = t % G, k = t / G;
        const cx<T> a = AHA[t];
        A1re[(N + k) * RG + g] = -a.re; A1im[(N + k) * RG + g] = -a.im;
        Qre[k * RG + g] = a.re; Qim[k * RG + g] = a.im;
    }
    for (int t = threadIdx.x + G * P; t < RG * Pc; t += kThreads) { Lre[t] = T(0); Lim[t] = T(0); }   // padding columns of the streamed operand
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int rg = warp % NGg, og = warp / NGg;
    const bool active = og < kWarps / NGg;
    const int o_0 = og * kOW + out_of<T>(lane, 0), o_1 = og * kOW + out_of<T>(lane, 1);
    T ar[kRB][2], ai[kRB][2];
    // exchange rows: [0,N) = summed T1 (row-major partials), [N,N+G) = V BBH.
    // V BBH is not recomputed: V_{i+1} = V_i + alpha Res_i  =>  V_{i+1} BBH = V_i BBH + alpha (Res_i BBH), and Res_i BBH is
    // formed below for the line search anyway (.m:47-50) - one pass over BBH per iteration instead of two.
    for (int t = threadIdx.x; t < G * P; t += kThreads) E[(size_t)(N + t / P) * pitch + (t % P)] = VBg[t];
    const int nt1 = p.nt1 ? p.nt1 : p.nmc;
    const cx<T>* T1 = p.T1 + (size_t)b * nt1 * N * P;
    for (int t = threadIdx.x; t < N * P; t += kThreads) {
        T re = 0, im = 0;
        for (int k = 0; k < nt1; ++k) { cx<T> v = T1[(size_t)k * N * P + t]; re += v.re; im += v.im; }
        E[(size_t)(t / P) * pitch + (t % P)] = mk<T>(re, im);
    }
    __syncthreads();
    // Res = A^H T1 - AHA (V BBH)        (.m:47)
    T rr_[kRB][2], ri_[kRB][2];
#pragma unroll
    for (int r = 0; r < kRB; ++r) { rr_[r][0] = rr_[r][1] = ri_[r][0] = ri_[r][1] = T(0); }
    if (active && og * kOW < pitch) smem_contract<T>(A1re, A1im, RG, rg, E, pitch, og, N + G, rr_, ri_);
    double rr = 0.0, vv = 0.0;
    if (active) {
#pragma unroll
        for (int r = 0; r < kRB; ++r) {
            if (o_0 < P) rr += (double)rr_[r][0] * rr_[r][0] + (double)ri_[r][0] * ri_[r][0];
            if (o_1 < P) rr += (double)rr_[r][1] * rr_[r][1] + (double)ri_[r][1] * ri_[r][1];
        }
    }
    __syncthreads();                                                   // V planes and E fully consumed
    if (active) {
#pragma unroll
        for (int r = 0; r < kRB; ++r) {
            if (o_0 < P) { Lre[o_0 * RG + rg * kRB + r] = rr_[r][0]; Lim[o_0 * RG + rg * kRB + r] = ri_[r][0]; }
            if (o_1 < P) { Lre[o_1 * RG + rg * kRB + r] = rr_[r][1]; Lim[o_1 * RG + rg * kRB + r] = ri_[r][1]; }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // exchange rows (generic proxy) -> ring refill (async proxy)
    __syncthreads();
    pipe.start(ring, bars, it, BBH, (long long)P, W, P, P);
#pragma unroll
    for (int r = 0; r < kRB; ++r) { ar[r][0] = ar[r][1] = ai[r][0] = ai[r][1] = T(0); }
    it = pipe.template run<true>(Lre, Lim, RG, NGg, ar, ai);
    if (active && og * kOW < pitch) {
#pragma unroll
        for (int r = 0; r < kRB; ++r) store_pair<T>(E + (size_t)(rg * kRB + r) * pitch + og * kOW, lane, ar[r][0], ai[r][0], ar[r][1], ai[r][1]);
    }
    __syncthreads();
    // <Res, Q>, Q = AHA (Res BBH)        (.m:48)
#pragma unroll
    for (int r = 0; r < kRB; ++r) { ar[r][0] = ar[r][1] = ai[r][0] = ai[r][1] = T(0); }
    if (active && og * kOW < pitch) smem_contract<T>(Qre, Qim, RG, rg, E, pitch, og, G, ar, ai);
    double rq = 0.0;
    if (active) {
#pragma unroll
        for (int r = 0; r < kRB; ++r) {
            if (o_0 < P) rq += (double)rr_[r][0] * ar[r][0] + (double)ri_[r][0] * ai[r][0];
            if (o_1 < P) rq += (double)rr_[r][1] * ar[r][1] + (double)ri_[r][1] * ai[r][1];
        }
        // |V|^2 of the iterate being left (third convergence diagnostic, .m:51)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int o = j == 0 ? o_0 : o_1;
            if (o < P) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    cx<T> v[4];
                    ld4c<T>(V + (size_t)o * G + rg * kRB + 4 * hf, v);
#pragma unroll
                    for (int u = 0; u < 4; ++u) vv += (double)v[u].re * v[u].re + (double)v[u].im * v[u].im;
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) { rr += __shfl_down_sync(0xffffffffu, rr, o); rq += __shfl_down_sync(0xffffffffu, rq, o); vv += __shfl_down_sync(0xffffffffu, vv, o); }
    if (lane == 0) { red[warp][0] = rr; red[warp][1] = rq; red[warp][2] = vv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0, d = 0;
        for (int w = 0; w < kWarps; ++w) { a += red[w][0]; c += red[w][1]; d += red[w][2]; }
        s_alpha = a / c;                                               // alpha = res'res / (res' R res)  (.m:48)
        double* o = p.dots + (size_t)b * p.npc * 4;
        o[0] = a; o[1] = c; o[2] = d; o[3] = 0.0;
        for (int k = 1; k < p.npc; ++k) { o[4 * k] = 0.0; o[4 * k + 1] = 0.0; o[4 * k + 2] = 0.0; o[4 * k + 3] = 0.0; }
    }
    __syncthreads();
    // V += alpha Res ; S = soft(V) (masked)          (.m:50,56 ; _angles.m:68)
    const T alpha = (T)s_alpha;
    const T thr = (T)(p.tauS[b] / p.rho[b]);
    cx<T>* __restrict__ S = p.S + (size_t)b * G * P;
    const unsigned char* mask = p.angles ? p.smask + (size_t)b * G * P : nullptr;
    if (active && og * kOW < pitch) {
        // V BBH <- V BBH + alpha (Res BBH); Res BBH still sits in exchange rows [0,G)
#pragma unroll
        for (int r = 0; r < kRB; ++r) {
            const int row = rg * kRB + r;
            cx<T> b0, b1;
            load_pair<T>(E + (size_t)row * pitch + og * kOW, lane, b0, b1);
            if (o_1 < P) {
                cx<T> v0, v1;
                load_pair<T>(VBg + (size_t)row * P + og * kOW, lane, v0, v1);
                store_pair<T>(VBg + (size_t)row * P + og * kOW, lane, v0.re + alpha * b0.re, v0.im + alpha * b0.im, v1.re + alpha * b1.re, v1.im + alpha * b1.im);
            } else if (o_0 < P) {
                cx<T>* q = VBg + (size_t)row * P + o_0;
                *q = mk<T>(q->re + alpha * b0.re, q->im + alpha * b0.im);
            }
        }
    }
    __syncthreads();                                                   // exchange rows are rewritten with S below
    if (active) {
        T s_re[kRB][2], s_im[kRB][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int o = j == 0 ? o_0 : o_1;
#pragma unroll
            for (int r = 0; r < kRB; ++r) { s_re[r][j] = 0; s_im[r][j] = 0; }
            if (o < P) {
                const size_t gi = (size_t)o * G + rg * kRB;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    cx<T> v[4], s[4];
                    ld4c<T>(V + gi + 4 * hf, v);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = 4 * hf + u;
                        v[u] = mk<T>(v[u].re + alpha * rr_[r][j], v[u].im + alpha * ri_[r][j]);
                        s[u] = mk<T>(soft1<T>(v[u].re, thr), soft1<T>(v[u].im, thr));
                        if (mask && !mask[gi + r]) s[u] = mk<T>(T(0), T(0));
                        s_re[r][j] = s[u].re; s_im[r][j] = s[u].im;
                    }
                    st4c<T>(V + gi + 4 * hf, v); st4c<T>(S + gi + 4 * hf, s);
                }
            }
        }
        if (og * kOW < pitch) {
#pragma unroll
            for (int r = 0; r < kRB; ++r) store_pair<T>(E + (size_t)(rg * kRB + r) * pitch + og * kOW, lane, s_re[r][0], s_im[r][0], s_re[r][1], s_im[r][1]);
        }
    }
    __syncthreads();
    // A S      (.m:58, left factor)
    {
        const int rgn = warp % NGn, ogn = warp / NGn;
        if (ogn < kWarps / NGn && ogn * kOW < pitch) {
#pragma unroll
            for (int r = 0; r < kRB; ++r) { ar[r][0] = ar[r][1] = ai[r][0] = ai[r][1] = T(0); }
            smem_contract<T>(Sre, Sim, RN, rgn, E, pitch, ogn, G, ar, ai);
            cx<T>* __restrict__ AS = p.AS + (size_t)b * N * P;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int o = ogn * kOW + out_of<T>(lane, j);
                if (o < P) {
                    const size_t gi = (size_t)o * N + rgn * kRB;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        cx<T> v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = mk<T>(ar[4 * hf + u][j], ai[4 * hf + u][j]);
                        st4c<T>(AS + gi + 4 * hf, v);
                    }
                }
            }
        }
    }
}

// =======================================================================================
// kernel 3
// =======================================================================================
constexpr int kPCH = 256;   // reduction columns of A S held in shared memory at a time
template <typename T>
struct XsFastSmem {
    __host__ __device__ static size_t region0(int N, int NG, bool conv) {
        size_t ring = StreamRing<T>::bytes(cta_width(NG));
        size_t xt = sizeof(cx<T>) * (size_t)cta_width(NG) * (N + 2);           // product tile parked for the coalesced epilogue
        size_t sc = conv ? (size_t)N * N * 2 * sizeof(T) : 0;                  // at least one Gram slice
        size_t m = ring > xt ? ring : xt;
        return m > sc ? m : sc;
    }
    static size_t bytes(int N, int NG, int P, bool conv) {
        int pch = round_up_to(P < kPCH ? P : kPCH, stage_cols<T>());
        size_t l = 2 * sizeof(T) * (size_t)N * pch;
        size_t cv = conv ? 2 * sizeof(T) * (size_t)N * cta_width(NG) : 0;      // V2 planes reuse the (A S) planes
        return region0(N, NG, conv) + (l > cv ? l : cv);
    }
};

template <typename T>
__global__ void __launch_bounds__(kThreads, sizeof(T) == 4 ? 2 : 1) k_xs_fast(AdmmP<T> p, const cx<T>* __restrict__ Bt, long long ld_Bt) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ PipeBars<kStages> pb;
    uint64_t* bars = pb.full;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int N = p.N, P = p.P, M = p.M, RP = p.RP, NG = p.NG;      // RP == N
    const int W = cta_width(NG);
    const int m0 = chunk * W;
    const int nvalid = (M - m0) < W ? (M - m0) : W;
    const bool conv = p.convd != nullptr;
    cx<T>* ring = reinterpret_cast<cx<T>*>(smem);
    const size_t r0b = XsFastSmem<T>::region0(N, NG, conv);
    T* Lre = reinterpret_cast<T*>(smem + r0b);
    const int pchmax = P < kPCH ? P : kPCH;
    const int pchp = round_up_to(pchmax, stage_cols<T>());
    T* Lim = Lre + (size_t)RP * pchp;
    const int cta_id = blockIdx.y * gridDim.x + blockIdx.x;
    JSTSP_STAMP(p, 2, cta_id, 0);
    pipe_bars_init(pb, smem, StreamRing<T>::bytes(W), (P % stage_cols<T>()) != 0 || (pchmax % stage_cols<T>()) != 0);
    {   // warm L2 with the X / V2 tiles the epilogue will read (written by earlier kernels of this iteration)
        const size_t off = (size_t)b * N * M + (size_t)m0 * N;
        const size_t tile_bytes = sizeof(cx<T>) * (size_t)N * nvalid;
        for (size_t o = (size_t)threadIdx.x * 128; o < tile_bytes; o += (size_t)kThreads * 128) {
            prefetch_l2(reinterpret_cast<const char*>(p.X + off) + o);
            prefetch_l2(reinterpret_cast<const char*>(p.V2 + off) + o);
        }
    }
    JSTSP_STAMP(p, 2, cta_id, 1);
    // B^T is stored tiled: [m-chunk][p][W] so that the reduction columns of a stage are contiguous
    const cx<T>* Btc = Bt + (long long)b * ld_Bt + (long long)chunk * P * W;
    const cx<T>* __restrict__ AS = p.AS + (size_t)b * N * P;
    StreamPipe<T> pipe;
    uint32_t it = 0;
    T ar[kRB][2], ai[kRB][2];
#pragma unroll
    for (int r = 0; r < kRB; ++r) { ar[r][0] = ar[r][1] = ai[r][0] = ai[r][1] = T(0); }
    for (int p0 = 0; p0 < P; p0 += pchmax) {
        const int np = (P - p0) < pchmax ? (P - p0) : pchmax;
        pipe.start(ring, bars, it, Btc + (long long)p0 * W, (long long)W, W, W, np);
        // stage (A S)(:, p0:p0+np) planar; zero the padding columns
        for (int t0 = 0; t0 < RP * pchp; t0 += 2 * 8 * kThreads) {
            cx<T> v[8][2];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + 2 * (u * kThreads + threadIdx.x);
                v[u][0] = mk<T>(T(0), T(0)); v[u][1] = v[u][0];
                if (t < N * np) ld2c<T>(AS + (size_t)N * p0 + t, v[u]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + 2 * (u * kThreads + threadIdx.x);
                if (t < RP * pchp) { Lre[t] = v[u][0].re; Lim[t] = v[u][0].im; Lre[t + 1] = v[u][1].re; Lim[t + 1] = v[u][1].im; }
            }
        }
        __syncthreads();
        JSTSP_STAMP(p, 2, cta_id, 2);
        it = pipe.template run<false>(Lre, Lim, RP, NG, ar, ai);
        JSTSP_STAMP(p, 2, cta_id, 3);
    }
    const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
    const int rg = warp % NG, og = warp / NG;
    // park the product tile in shared memory ([column][N+2] complex, the ring is idle now) so that the
    // element-wise C / V2 update below runs with fully coalesced 16-byte global accesses
    const int NPp = N + 2;
    cx<T>* Xt = reinterpret_cast<cx<T>*>(smem);
    T* Vre = Lre;                                   // conv: V2 planes reuse the (A S) planes
    T* Vim = Lre + (size_t)RP * W;
    if (og < kWarps / NG) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = og * kOW + out_of<T>(lane, j);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                cx<T> v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = mk<T>(ar[4 * hf + u][j], ai[4 * hf + u][j]);
                st4c<T>(Xt + (size_t)c * NPp + rg * kRB + 4 * hf, v);
            }
        }
    }
    if (conv) for (int t = threadIdx.x + nvalid * RP; t < 2 * RP * W; t += kThreads) if (t < RP * W || t >= RP * W + nvalid * RP) Vre[t] = 0;
    __syncthreads();
    {
        const T rho = (T)p.rho[b];
        const T irho = T(1) / rho, kap = rho / (rho + T(1));
        const size_t off = (size_t)b * N * M + (size_t)m0 * N;
        const cx<T>* __restrict__ X = p.X + off; cx<T>* __restrict__ V2 = p.V2 + off; cx<T>* __restrict__ C = p.C + off; cx<T>* __restrict__ Xs = p.Xs + off;
        const int nel = N * nvalid;
        constexpr int EB = sizeof(T) == 4 ? 8 : 4;      // element pairs in flight per thread
        for (int e0 = 0; e0 < nel; e0 += 2 * kThreads * EB) {
            cx<T> x[EB][2], v2[EB][2];
#pragma unroll
            for (int u = 0; u < EB; ++u) { const int e = e0 + 2 * (u * kThreads + threadIdx.x); if (e < nel) { ld2c<T>(X + e, x[u]); ld2c<T>(V2 + e, v2[u]); } }
#pragma unroll
            for (int u = 0; u < EB; ++u) {
                const int e = e0 + 2 * (u * kThreads + threadIdx.x);
                if (e >= nel) continue;
                const int c = e / N, r = e % N;
                cx<T> sv[2], so[2], co[2], vo[2];
                ld2c<T>(Xt + (size_t)c * NPp + r, sv);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const T sr = sv[q].re, si = sv[q].im;
                    const T cr = kap * (x[u][q].re - sr - irho * v2[u][q].re), ci = kap * (x[u][q].im - si - irho * v2[u][q].im);   // .m:61
                    const T nr = v2[u][q].re + rho * (cr - x[u][q].re + sr), ni = v2[u][q].im + rho * (ci - x[u][q].im + si);       // .m:65
                    so[q] = mk<T>(sr, si); co[q] = mk<T>(cr, ci); vo[q] = mk<T>(nr, ni);
                    if (conv) { Vre[e + q] = nr; Vim[e + q] = ni; }
                }
                st2c<T>(Xs + e, so); st2c<T>(C + e, co); st2c<T>(V2 + e, vo);
            }
        }
    }
    JSTSP_STAMP(p, 2, cta_id, 4);
    if (p.dbg && p.dbg_kernel == 2 && threadIdx.x == 0) { unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm)); p.dbg[(size_t)cta_id * 8 + 7] = sm; }
    if (conv) {
        __syncthreads();
        gram_blocked<T>(Vre, Vim, RP, N, nvalid, reinterpret_cast<T*>(smem), r0b, p.cgramB + ((size_t)b * p.nxc + chunk) * 2 * N * N);
    }
}

}  // namespace jstsp
