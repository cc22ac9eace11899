// system.cu - measurement model and driver-side metric kernels (batched, one CTA per trial).
//
//   jstsp_wideband_mmwave_channel             replaces basic_system_functions/wideband_mmwave_channel.m:1-62
//   jstsp_proposed_hbf / jstsp_hbf            replace  basic_system_functions/proposed_hbf.m:1-44, hbf.m:1-26
//   jstsp_wideband_hybBF_comm_system_training replaces basic_system_functions/wideband_hybBF_comm_system_training.m:1-58
//   jstsp_nmse                                the driver metric norm(S-Zbar)^2/norm(Zbar)^2, clipped (plot_errorVSsnr.m:138-141)
//   jstsp_admm_parameters                     tau_Y, tau_Z, rho of plot_errorVSsnr.m:127-130 (eigs -> 6th largest eigenvalue)
//
// Randomness is an INPUT: the functions that draw in the reference (randn / rand / randperm) take
// the draws as arrays in the reference's consumption order, so a MEX gateway can obtain them from
// MATLAB's own generator (mexCallMATLAB) and the batched engine from its counter-based generator.
#include <algorithm>
#include "common.cuh"
#include "jacobi.cuh"

namespace jstsp {

// ---------------------------------------------------------------------------------------------
// channel generator
// ---------------------------------------------------------------------------------------------
template <typename T>
struct ChanP {
    int L, Mr, Mt, ncl, nray, Gr, Gt;
    const double* normals;   // [b][L*Np*2]  (re, im) of the Rayleigh coefficient, per (l, ray)   (.m:19)
    const double* uniforms;  // [b][L*Np*2]  (u for phi_r, u for phi_t), per (l, ray)             (.m:20,22)
    cx<T>*H, *Zbar, *Ar, *At, *Dr, *Dt;      // outputs (any may be null)
    cx<T>* ws;                                // [b][Mr*Mt + Mr*Gt] scratch
};

__device__ __forceinline__ double laplacian_angle(double u) {        // genLaplacianSamples (.m:56-62)
    const double beta = 1.0 / (1.0 - exp(-sqrt(2.0) * M_PI / 50.0));
    return beta * (exp(-sqrt(2.0) / 50.0 * M_PI) - cosh(u));
}

template <typename T>
__global__ void __launch_bounds__(256) k_channel(ChanP<T> p) {
    const int b = blockIdx.x, L = p.L, Mr = p.Mr, Mt = p.Mt, Np = p.ncl * p.nray, Gr = p.Gr, Gt = p.Gt;
    extern __shared__ __align__(16) unsigned char smem[];
    double* phr = reinterpret_cast<double*>(smem);      // sin(-phi_r) of tap-1 rays (page-1 quirk, .m:24)
    double* pht = phr + Np;
    const double* nrm = p.normals + (size_t)b * L * Np * 2;
    const double* uni = p.uniforms + (size_t)b * L * Np * 2;
    for (int k = threadIdx.x; k < Np; k += blockDim.x) {
        phr[k] = sin(0.0 - laplacian_angle(uni[2 * k]));           // angle(): exp(-1j*pi*sin(phi0-phi)*(0:M-1)')  (.m:42-52)
        pht[k] = sin(0.0 - laplacian_angle(uni[2 * k + 1]));
    }
    __syncthreads();
    // steering matrices returned to the caller use every tap's own angles (.m:21-22)
    if (p.Ar) for (int t = threadIdx.x; t < Mr * Np * L; t += blockDim.x) {
        const int i = t % Mr, k = (t / Mr) % Np, l = t / (Mr * Np);
        double s, c; sincos(-M_PI * sin(0.0 - laplacian_angle(uni[2 * (l * Np + k)])) * i, &s, &c);
        p.Ar[(size_t)b * Mr * Np * L + t] = mk<T>((T)c, (T)s);
    }
    if (p.At) for (int t = threadIdx.x; t < Mt * Np * L; t += blockDim.x) {
        const int i = t % Mt, k = (t / Mt) % Np, l = t / (Mt * Np);
        double s, c; sincos(-M_PI * sin(0.0 - laplacian_angle(uni[2 * (l * Np + k) + 1])) * i, &s, &c);
        p.At[(size_t)b * Mt * Np * L + t] = mk<T>((T)c, (T)s);
    }
    if (p.Dr) for (int t = threadIdx.x; t < Mr * Gr; t += blockDim.x) {
        const int i = t % Mr, g = t / Mr;
        double s, c; sincos(-2.0 * M_PI * (double)i * g / Gr, &s, &c);
        p.Dr[(size_t)b * Mr * Gr + t] = mk<T>((T)(c / sqrt((double)Mr)), (T)(s / sqrt((double)Mr)));          // .m:9
    }
    if (p.Dt) for (int t = threadIdx.x; t < Mt * Gt; t += blockDim.x) {
        const int i = t % Mt, g = t / Mt;
        double s, c; sincos(-2.0 * M_PI * (double)i * g / Gt, &s, &c);
        p.Dt[(size_t)b * Mt * Gt + t] = mk<T>((T)(c / sqrt((double)Mt)), (T)(s / sqrt((double)Mt)));          // .m:10
    }
    // tables: steering vectors of the tap-1 rays, DFT twiddles of both grids (the products below index them instead of calling sincos per term)
    double* arr = pht + Np;                    // [Np][Mr][2]
    double* att = arr + 2 * Np * Mr;           // [Np][Mt][2]
    double* twr = att + 2 * Np * Mt;           // [Gr][2]   exp(+2 pi j q / Gr)
    double* twt = twr + 2 * Gr;                // [Gt][2]   exp(-2 pi j q / Gt)
    for (int t = threadIdx.x; t < Np * Mr; t += blockDim.x) { double sn, cs; sincos(-M_PI * phr[t / Mr] * (t % Mr), &sn, &cs); arr[2 * t] = cs; arr[2 * t + 1] = sn; }
    for (int t = threadIdx.x; t < Np * Mt; t += blockDim.x) { double sn, cs; sincos(-M_PI * pht[t / Mt] * (t % Mt), &sn, &cs); att[2 * t] = cs; att[2 * t + 1] = sn; }
    for (int t = threadIdx.x; t < Gr; t += blockDim.x) { double sn, cs; sincos(2.0 * M_PI * (double)t / Gr, &sn, &cs); twr[2 * t] = cs; twr[2 * t + 1] = sn; }
    for (int t = threadIdx.x; t < Gt; t += blockDim.x) { double sn, cs; sincos(-2.0 * M_PI * (double)t / Gt, &sn, &cs); twt[2 * t] = cs; twt[2 * t + 1] = sn; }
    __syncthreads();
    cx<T>* Hl = p.ws + (size_t)b * (Mr * Mt + Mr * Gt);
    cx<T>* tmp = Hl + Mr * Mt;
    const double scale = 1.0 / sqrt((double)Np);
    for (int l = 0; l < L; ++l) {
        // H_l = 1/sqrt(Np) sum_ray w_ray coef a_r a_t^H ; cluster c rays are counted (ncl - c) times (.m:24-33)
        for (int t = threadIdx.x; t < Mr * Mt; t += blockDim.x) {
            const int i = t % Mr, j = t / Mr;
            double re = 0.0, im = 0.0;
            for (int k = 0; k < Np; ++k) {
                const double w = (double)(p.ncl - k / p.nray) * scale / sqrt(2.0);
                const double cr = nrm[2 * (l * Np + k)] * w, ci = nrm[2 * (l * Np + k) + 1] * w;
                const double ac = arr[2 * (k * Mr + i)], as = arr[2 * (k * Mr + i) + 1], bc = att[2 * (k * Mt + j)], bs = att[2 * (k * Mt + j) + 1];
                const double c = ac * bc + as * bs, sn = as * bc - ac * bs;         // a_r(i) conj(a_t(j))
                re += cr * c - ci * sn; im += cr * sn + ci * c;
            }
            Hl[t] = mk<T>((T)re, (T)im);
            if (p.H) p.H[(size_t)b * Mr * Mt * L + (size_t)l * Mr * Mt + t] = Hl[t];
        }
        __syncthreads();
        if (p.Zbar) {
            // Z_l = Dr' H_l Dt ; Zbar(:, l*Gt + g) = Z_l(:, g)  (.m:35,38)
            for (int t = threadIdx.x; t < Mr * Gt; t += blockDim.x) {       // tmp = H_l Dt
                const int i = t % Mr, g = t / Mr;
                double re = 0.0, im = 0.0;
                for (int j = 0; j < Mt; ++j) {
                    const int q = (int)(((long long)j * g) % Gt);
                    const double c = twt[2 * q], sn = twt[2 * q + 1];
                    const cx<T> hv = Hl[i + Mr * j];
                    re += hv.re * c - hv.im * sn; im += hv.re * sn + hv.im * c;
                }
                tmp[t] = mk<T>((T)(re / sqrt((double)Mt)), (T)(im / sqrt((double)Mt)));
            }
            __syncthreads();
            for (int t = threadIdx.x; t < Gr * Gt; t += blockDim.x) {
                const int gr = t % Gr, g = t / Gr;
                double re = 0.0, im = 0.0;
                for (int i = 0; i < Mr; ++i) {
                    const int q = (int)(((long long)i * gr) % Gr);                    // conj(Dr(i,gr))
                    const double c = twr[2 * q], sn = twr[2 * q + 1];
                    const cx<T> v = tmp[i + Mr * g];
                    re += v.re * c - v.im * sn; im += v.re * sn + v.im * c;
                }
                p.Zbar[(size_t)b * Gr * Gt * L + (size_t)(l * Gt + g) * Gr + gr] = mk<T>((T)(re / sqrt((double)Mr)), (T)(im / sqrt((double)Mr)));
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// measurement synthesis:  Y = sum_l H_l Psi_l ; out = [Omega .*] (W_e' (Y + N))
// ---------------------------------------------------------------------------------------------
template <typename T>
struct MeasP {
    int Nr, Nt, L, T_, Wc, Lr;          // T_ training columns; W_e = W(:, 1:Wc); Lr ones per mask column (0 = no mask)
    const cx<T>* H;  long long ld_H;     // Nr x Nt x L
    const cx<T>* N;  long long ld_N;     // Nr x T
    const cx<T>* Psi; long long ld_Psi;  // psi_mode 0: Psi_i (Tp x Tp x Nt, only rows 1..L read); 1: pilots s_k (Nt x T, row k = s_k)
    int psi_mode, Tp;
    const cx<T>* W;  long long ld_W;     // Nr x (>= Wc)
    const int* perm; long long ld_perm;  // mask draws: T x Wc int32 (1-based randperm per column), or null
    cx<T>* Yout; cx<T>* Ynl; cx<T>* We; cx<T>* Psibar; T* Omega;     // outputs (any may be null)
    cx<T>* ws;                            // [b][Nr * T] scratch for R
};

template <typename T>
__device__ __forceinline__ cx<T> psi_at(const MeasP<T>& p, const cx<T>* Psi, int k, int t, int l) {
    if (p.psi_mode == 0) return Psi[l + (size_t)p.Tp * t + (size_t)p.Tp * p.Tp * k];            // Psi_bar(k,t,l) = Psi_i(l,t,k)  (proposed_hbf.m:17)
    const int d = t - l;                                                                         // row l of toeplitz(s_k)
    cx<T> v = Psi[k + (size_t)p.Nt * (d >= 0 ? d : -d)];
    if (d < 0) v.im = -v.im;
    return v;
}

template <typename T>
__global__ void __launch_bounds__(256) k_measure(MeasP<T> p) {
    const int b = blockIdx.y, Nr = p.Nr, Nt = p.Nt, L = p.L, TT = p.T_, Wc = p.Wc;
    const cx<T>* H = p.H + (long long)b * p.ld_H;
    const cx<T>* Psi = p.Psi + (long long)b * p.ld_Psi;
    const cx<T>* W = p.W + (long long)b * p.ld_W;
    const cx<T>* N = p.N ? p.N + (long long)b * p.ld_N : nullptr;
    cx<T>* R = p.ws + (size_t)b * Nr * TT;
    // column tile of this CTA
    const int tile = (TT + gridDim.x - 1) / gridDim.x;
    const int t0 = blockIdx.x * tile, t1 = (t0 + tile) < TT ? (t0 + tile) : TT;
    for (int e = threadIdx.x; e < Nr * (t1 - t0); e += blockDim.x) {
        const int r = e % Nr, t = t0 + e / Nr;
        T re = 0, im = 0;
        for (int l = 0; l < L; ++l)
            for (int k = 0; k < Nt; ++k) {
                const cx<T> hv = H[r + (size_t)Nr * k + (size_t)Nr * Nt * l];
                const cx<T> ps = psi_at<T>(p, Psi, k, t, l);
                cmac<T>(re, im, hv.re, hv.im, ps.re, ps.im);
            }
        if (p.Ynl) p.Ynl[(size_t)b * Nr * TT + (size_t)t * Nr + r] = mk<T>(re, im);       // noiseless Y (proposed_hbf.m:14-20)
        if (N) { const cx<T> nv = N[r + (size_t)Nr * t]; re += nv.re; im += nv.im; }      // R = Y + N (:22)
        R[(size_t)t * Nr + r] = mk<T>(re, im);
    }
    if (p.Psibar) for (int e = threadIdx.x; e < Nt * (t1 - t0) * L; e += blockDim.x) {
        const int k = e % Nt, t = t0 + (e / Nt) % (t1 - t0), l = e / (Nt * (t1 - t0));
        p.Psibar[(size_t)b * Nt * TT * L + (size_t)l * Nt * TT + (size_t)t * Nt + k] = psi_at<T>(p, Psi, k, t, l);
    }
    if (p.We && blockIdx.x == 0) for (int e = threadIdx.x; e < Nr * Wc; e += blockDim.x) p.We[(size_t)b * Nr * Wc + e] = W[e];
    __syncthreads();
    for (int e = threadIdx.x; e < Wc * (t1 - t0); e += blockDim.x) {
        const int q = e % Wc, t = t0 + e / Wc;
        T re = 0, im = 0;
        for (int r = 0; r < Nr; ++r) { const cx<T> w = W[r + (size_t)Nr * q], x = R[(size_t)t * Nr + r]; cmac<T>(re, im, w.re, -w.im, x.re, x.im); }
        T om = T(1);
        if (p.perm) {          // Omega(indices(1:Lr), t) = 1 with indices = randperm(Wc)  (proposed_hbf.m:36-41)
            om = T(0);
            const int* pr = p.perm + (long long)b * p.ld_perm + (size_t)t * Wc;
            for (int j = 0; j < p.Lr; ++j) if (pr[j] == q + 1) om = T(1);
            if (p.Omega) p.Omega[(size_t)b * Wc * TT + (size_t)t * Wc + q] = om;
        }
        if (p.Yout) p.Yout[(size_t)b * Wc * TT + (size_t)t * Wc + q] = mk<T>(om * re, om * im);     // Omega .* (W_e' R) (:42)
    }
}

// Same synthesis for pilot-sequence input (psi_mode 1) with the operands staged in shared memory: H (all taps), the (columns + L - 1) window
// of e_k(t) = s_k(t) | conj(s_k(-t)), then R and W.  Thread = (column, 8 rows): per (tap, antenna) one pilot load and one broadcast 8-row
// slice of H feed 8 complex MACs.  grid (ceil(T / CTc), batch), block 256, CTc = 256 / (Nr / 8) columns per CTA.
template <typename T>
__global__ void __launch_bounds__(256) k_measure_tiled(MeasP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.y, Nr = p.Nr, Nt = p.Nt, L = p.L, TT = p.T_, Wc = p.Wc;
    const int RG = Nr / 8, CTc = 256 / RG, WP = CTc + L - 1;
    cx<T>* sH = reinterpret_cast<cx<T>*>(smem);                  // [(l * Nt + k)][Nr]
    cx<T>* sW = sH + (size_t)L * Nt * Nr;                        // [r][Wc + 1]: W(r, q) at r (Wc + 1) + q (lanes run over q: the padding spreads the banks)
    cx<T>* sP = sW + (size_t)Nr * (Wc + 1);                      // [k][WP]; reused as R [row][CTc + 1]
    const cx<T>* H = p.H + (long long)b * p.ld_H;
    const cx<T>* Pil = p.Psi + (long long)b * p.ld_Psi;
    const cx<T>* W = p.W + (long long)b * p.ld_W;
    const int t0 = blockIdx.x * CTc;
    if constexpr (sizeof(T) == 4) {      // fp32: planar re | im so that row pairs are 64-bit operands of the packed FMA below
        float* hre = reinterpret_cast<float*>(sH); float* him = hre + (size_t)L * Nt * Nr;
#pragma unroll 8
        for (int e = threadIdx.x; e < L * Nt * Nr; e += 256) { const cx<T> v = H[e]; hre[e] = v.re; him[e] = v.im; }
    } else {
        for (int e = threadIdx.x; e < L * Nt * Nr; e += 256) sH[e] = H[e];                   // H(r, k, l) at r + Nr k + Nr Nt l
    }
    for (int e = threadIdx.x; e < Nr * Wc; e += 256) sW[(e % Nr) * (Wc + 1) + e / Nr] = W[e];
    if (t0 - (L - 1) >= 0 && t0 + CTc <= TT) {           // interior tile: a straight copy, eight loads in flight per thread
#pragma unroll 8
        for (int e = threadIdx.x; e < Nt * WP; e += 256) sP[(size_t)(e % Nt) * WP + e / Nt] = Pil[(size_t)Nt * (t0 - (L - 1)) + e];
    } else {
        for (int e = threadIdx.x; e < Nt * WP; e += 256) {
            const int k = e % Nt, w = e / Nt, tau = t0 - (L - 1) + w;
            cx<T> v = mk<T>(T(0), T(0));
            if (tau >= 0 && tau < TT) v = Pil[k + (size_t)Nt * tau];
            else if (tau < 0 && -tau < TT) v = conj(Pil[k + (size_t)Nt * (-tau)]);            // row l of toeplitz(s_k) below the diagonal
            sP[(size_t)k * WP + w] = v;
        }
    }
    __syncthreads();
    const int c = threadIdx.x % CTc, rg = threadIdx.x / CTc, t = t0 + c;
    T ar[8] = {}, ai[8] = {};
    if constexpr (sizeof(T) == 4) {
        // packed fp32x2 FMAs (Blackwell FFMA2) on row pairs: R += Hre * p.re ; R += Him * (-p.im) ; I += Hre * p.im ; I += Him * p.re - the same
        // operations in the same order as cmac(), two rows per instruction (rounds exactly like the scalar form)
        const float* hre = reinterpret_cast<const float*>(sH); const float* him = hre + (size_t)L * Nt * Nr;
        uint64_t R[4] = {0, 0, 0, 0}, I[4] = {0, 0, 0, 0};
        auto pk = [](float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; };
        auto f2 = [](uint64_t& d, uint64_t a, uint64_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b)); };
        for (int l = 0; l < L; ++l) {
            const cx<T>* pw = sP + (c + (L - 1) - l);
            const size_t h0 = (size_t)l * Nt * Nr + 8 * rg;
#pragma unroll 4
            for (int k = 0; k < Nt; ++k) {
                const cx<T> ps = pw[(size_t)k * WP];
                const uint64_t PR = pk(ps.re, ps.re), PI = pk(ps.im, ps.im), PN = pk(-ps.im, -ps.im);
                const ulonglong2 a0 = *reinterpret_cast<const ulonglong2*>(hre + h0 + (size_t)k * Nr), a1 = *reinterpret_cast<const ulonglong2*>(hre + h0 + (size_t)k * Nr + 4);
                const ulonglong2 b0 = *reinterpret_cast<const ulonglong2*>(him + h0 + (size_t)k * Nr), b1 = *reinterpret_cast<const ulonglong2*>(him + h0 + (size_t)k * Nr + 4);
                const uint64_t HR[4] = {a0.x, a0.y, a1.x, a1.y}, HI[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                for (int q = 0; q < 4; ++q) { f2(R[q], HR[q], PR); f2(R[q], HI[q], PN); f2(I[q], HR[q], PI); f2(I[q], HI[q], PR); }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(R[q])); ar[2 * q] = lo; ar[2 * q + 1] = hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(I[q])); ai[2 * q] = lo; ai[2 * q + 1] = hi;
        }
    } else {
        for (int l = 0; l < L; ++l) {
            const cx<T>* pw = sP + (c + (L - 1) - l);
            const cx<T>* hh = sH + (size_t)l * Nt * Nr + 8 * rg;
            for (int k = 0; k < Nt; ++k) {
                const cx<T> ps = pw[(size_t)k * WP];
                const cx<T>* hv = hh + (size_t)k * Nr;
#pragma unroll
                for (int u = 0; u < 8; ++u) cmac<T>(ar[u], ai[u], hv[u].re, hv[u].im, ps.re, ps.im);
            }
        }
    }
    __syncthreads();                                             // the pilot window is dead: its space takes R
    cx<T>* sR = sP;
    if (t < TT) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = 8 * rg + u;
            T re = ar[u], im = ai[u];
            if (p.Ynl) p.Ynl[(size_t)b * Nr * TT + (size_t)t * Nr + r] = mk<T>(re, im);       // noiseless Y (proposed_hbf.m:14-20)
            if (p.N) { const cx<T> nv = p.N[(long long)b * p.ld_N + r + (size_t)Nr * t]; re += nv.re; im += nv.im; }   // R = Y + N (:22)
            sR[(size_t)r * (CTc + 1) + c] = mk<T>(re, im);
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < Wc * CTc; e += 256) {
        const int q = e % Wc, cc = e / Wc, tt = t0 + cc;
        if (tt >= TT) break;
        T re = 0, im = 0;
        for (int r = 0; r < Nr; ++r) { const cx<T> w = sW[r * (Wc + 1) + q], x = sR[(size_t)r * (CTc + 1) + cc]; cmac<T>(re, im, w.re, -w.im, x.re, x.im); }
        T om = T(1);
        if (p.perm) {          // Omega(indices(1:Lr), t) = 1 with indices = randperm(Wc)  (proposed_hbf.m:36-41)
            om = T(0);
            const int* pr = p.perm + (long long)b * p.ld_perm + (size_t)tt * Wc;
            for (int j = 0; j < p.Lr; ++j) if (pr[j] == q + 1) om = T(1);
            if (p.Omega) p.Omega[(size_t)b * Wc * TT + (size_t)tt * Wc + q] = om;
        }
        if (p.Yout) p.Yout[(size_t)b * Wc * TT + (size_t)tt * Wc + q] = mk<T>(om * re, om * im);     // Omega .* (W_e' R) (:42)
    }
}

// ---------------------------------------------------------------------------------------------
// spectral-norm NMSE and the driver-side ADMM parameters
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ void gram_rows(JacobiSmem& sm, const cx<T>* A, const cx<T>* Bsub, int g, int ncol) {
    // sm.A = E E^H with E = A - Bsub (Bsub may be null), E is g x ncol column-major
    for (int t = threadIdx.x; t < g * g; t += blockDim.x) {
        const int i = t % g, j = t / g;
        double re = 0.0, im = 0.0;
        for (int c = 0; c < ncol; ++c) {
            cx<T> a = A[i + (size_t)g * c], d = A[j + (size_t)g * c];
            if (Bsub) { const cx<T> x = Bsub[i + (size_t)g * c], y = Bsub[j + (size_t)g * c]; a = mk<T>(a.re - x.re, a.im - x.im); d = mk<T>(d.re - y.re, d.im - y.im); }
            re += (double)a.re * d.re + (double)a.im * d.im; im += (double)a.im * d.re - (double)a.re * d.im;
        }
        sm.Are[t] = re; sm.Aim[t] = im;
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(256) k_nmse(const cx<T>* S, long long ld_S, const cx<T>* Z, long long ld_Z, int G, int P, double* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    JacobiSmem sm; sm.carve(smem, G);
    const int b = blockIdx.x;
    const cx<T>* s = S + (long long)b * ld_S; const cx<T>* z = Z + (long long)b * ld_Z;
    double num = 0.0, den = 0.0;
    gram_rows<T>(sm, s, z, G, P);
    jacobi_hermitian_block(sm, G, 24, false, sizeof(T) == 4 ? 1e-10 : 1e-20);      // fp32 callers: eigenvalues to 1e-10 relative are beyond their inputs
    for (int k = 0; k < G; ++k) num = fmax(num, sm.Are[k + G * k]);
    __syncthreads();
    gram_rows<T>(sm, z, nullptr, G, P);
    jacobi_hermitian_block(sm, G, 24, false, sizeof(T) == 4 ? 1e-10 : 1e-20);
    for (int k = 0; k < G; ++k) den = fmax(den, sm.Are[k + G * k]);
    if (threadIdx.x == 0) { double e = num / den; out[b] = e > 1.0 ? 1.0 : e; }     // clipped at 1 (plot_errorVSsnr.m:139-141)
}

template <typename T>
__global__ void __launch_bounds__(256) k_params(const cx<T>* Y, long long ld_Y, int N, int M, const cx<T>* Z, long long ld_Z, int G, int P, int kth,
                                                double* tauY, double* tauZ, double* rho) {
    extern __shared__ __align__(16) unsigned char smem[];
    JacobiSmem sm; sm.carve(smem, N);
    __shared__ double red[8];
    const int b = blockIdx.x;
    const cx<T>* y = Y + (long long)b * ld_Y;
    gram_rows<T>(sm, y, nullptr, N, M);
    double fy = 0.0;                                         // ||Y||_F^2 = trace of the Gram matrix
    for (int k = 0; k < N; ++k) fy += sm.Are[k + N * k];
    __syncthreads();
    jacobi_hermitian_block(sm, N, 24, false, sizeof(T) == 4 ? 1e-10 : 1e-20);
    if (threadIdx.x == 0) {
        // eigs(Y'Y): the kth largest eigenvalue (6 by default -> min(eigs), 1 -> max) (plot_errorVSsnr.m:129-130)
        double ev[64];
        for (int k = 0; k < N; ++k) ev[k] = sm.Are[k + N * k];
        for (int i = 0; i < N; ++i) for (int j = i + 1; j < N; ++j) if (ev[j] > ev[i]) { double t = ev[i]; ev[i] = ev[j]; ev[j] = t; }
        // eigs works on the M x M matrix Y'Y: with fewer than kth eigenvalues it returns all min(N, M) non-trivial ones (plus exact zeros when M > N)
        const int kk = kth <= (N < M ? N : M) ? kth : (M > N ? kth : (N < M ? N : M));
        const double lam = kk <= N ? ev[kk - 1] : 0.0;
        tauY[b] = 1.0 / fy;
        rho[b] = sqrt(fmax(lam, 0.0) / fy);
    }
    if (Z && tauZ) {
        const cx<T>* z = Z + (long long)b * ld_Z;
        double f = 0.0;
        for (size_t t = threadIdx.x; t < (size_t)G * P; t += blockDim.x) f += (double)z[t].re * z[t].re + (double)z[t].im * z[t].im;
        for (int o = 16; o > 0; o >>= 1) f += __shfl_xor_sync(0xffffffffu, f, o);
        if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = f;
        __syncthreads();
        if (threadIdx.x == 0) { double a = 0.0; for (int w = 0; w < 8; ++w) a += red[w]; tauZ[b] = 0.5 / a; }     // 1/norm(Zbar,'fro')^2/2 (:128)
    }
}

// ---------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------
}  // namespace jstsp

using namespace jstsp;

template <typename T>
static int run_channel(Handle* h, int mem, int L, int Mr, int Mt, int ncl, int nray, int Gr, int Gt, int batch, const double* normals, const double* uniforms,
                       void* H, void* Zbar, void* Ar, void* At, void* Dr, void* Dt) {
    if (L <= 0 || Mr <= 0 || Mt <= 0 || ncl <= 0 || nray <= 0 || Gr <= 0 || Gt <= 0 || batch <= 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if (!normals || !uniforms) return fail(h, JSTSP_E_ARG, "NULL draws");
    const bool host = mem == JSTSP_HOST;
    const int Np = ncl * nray;
    const size_t nH = (size_t)Mr * Mt * L * batch, nZ = (size_t)Gr * Gt * L * batch, nAr = (size_t)Mr * Np * L * batch, nAt = (size_t)Mt * Np * L * batch,
                 nDr = (size_t)Mr * Gr * batch, nDt = (size_t)Mt * Gt * batch, nd = (size_t)L * Np * 2 * batch;
    cx<T>* outs[6] = {(cx<T>*)H, (cx<T>*)Zbar, (cx<T>*)Ar, (cx<T>*)At, (cx<T>*)Dr, (cx<T>*)Dt};
    const size_t ns[6] = {nH, nZ, nAr, nAt, nDr, nDt};
    ChanP<T> p{};
    for (int pass = 0; pass < 2; ++pass) {
        Arena ar(pass ? h->ws : nullptr, pass ? h->ws_bytes : 0);
        p = ChanP<T>{};
        p.L = L; p.Mr = Mr; p.Mt = Mt; p.ncl = ncl; p.nray = nray; p.Gr = Gr; p.Gt = Gt;
        p.ws = ar.take<cx<T>>((size_t)batch * (Mr * Mt + Mr * Gt));
        cx<T>* dev[6];
        for (int k = 0; k < 6; ++k) dev[k] = outs[k] ? (host ? ar.take<cx<T>>(ns[k]) : outs[k]) : nullptr;
        double* dn = host ? ar.take<double>(nd) : const_cast<double*>(normals);
        double* du = host ? ar.take<double>(nd) : const_cast<double*>(uniforms);
        if (!pass) { int rc = ensure_workspace(h, ar.off); if (rc) return rc; continue; }
        if (host) {
            JSTSP_CUDA(h, cudaMemcpyAsync(dn, normals, nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            JSTSP_CUDA(h, cudaMemcpyAsync(du, uniforms, nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        }
        p.normals = dn; p.uniforms = du;
        p.H = dev[0]; p.Zbar = dev[1]; p.Ar = dev[2]; p.At = dev[3]; p.Dr = dev[4]; p.Dt = dev[5];
        const size_t sm_ch = sizeof(double) * 2 * ((size_t)Np + (size_t)Np * Mr + (size_t)Np * Mt + Gr + Gt);
        { int rc = set_smem(h, k_channel<T>, sm_ch); if (rc) return rc; }
        JSTSP_LAUNCH(h, PK_OTHER, (k_channel<T><<<batch, 256, sm_ch, h->stream>>>(p)));
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            for (int k = 0; k < 6; ++k) if (outs[k]) JSTSP_CUDA(h, cudaMemcpyAsync(outs[k], dev[k], ns[k] * sizeof(cx<T>), cudaMemcpyDeviceToHost, h->stream));
            JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
        }
    }
    return JSTSP_OK;
}

extern "C" int jstsp_wideband_mmwave_channel(jstsp_handle* h, int dtype, int mem, int L, int Mr, int Mt, int ncl, int nray, int Gr, int Gt, int batch,
                                             const double* normals, const double* uniforms,
                                             void* H, void* Zbar, void* Ar, void* At, void* Dr, void* Dt) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_channel<float>(h, mem, L, Mr, Mt, ncl, nray, Gr, Gt, batch, normals, uniforms, H, Zbar, Ar, At, Dr, Dt);
    if (dtype == JSTSP_F64) return run_channel<double>(h, mem, L, Mr, Mt, ncl, nray, Gr, Gt, batch, normals, uniforms, H, Zbar, Ar, At, Dr, Dt);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

template <typename T>
static int run_measure(Handle* h, int mem, const jstsp_meas_desc* d, const void* H, const void* N, const void* Psi, const void* W, const int* perm,
                       void* Yout, void* We, void* Psibar, void* Omega, void* Ynl) {
    const int Nr = d->Nr, Nt = d->Nt, L = d->L, TT = d->T, Wc = d->Wc, batch = d->batch;
    if (Nr <= 0 || Nt <= 0 || L <= 0 || TT <= 0 || Wc <= 0 || Wc > Nr || batch <= 0 || d->Lr < 0 || d->Lr > Wc) return fail(h, JSTSP_E_ARG, "bad dimension");
    if (!H || !Psi || !W) return fail(h, JSTSP_E_ARG, "NULL buffer");
    if (d->psi_mode == 0 && (d->Tp < TT || d->Tp < L)) return fail(h, JSTSP_E_ARG, "Psi_i must be at least T x T with T >= L");
    if (d->psi_mode == 1 && L > TT) return fail(h, JSTSP_E_ARG, "more delay taps than training columns (the Toeplitz rows are built from T pilot symbols)");
    const bool host = mem == JSTSP_HOST;
    const size_t nH = (size_t)Nr * Nt * L, nN = (size_t)Nr * TT, nW = (size_t)Nr * Nr,
                 nPsi = d->psi_mode == 0 ? (size_t)d->Tp * d->Tp * Nt : (size_t)Nt * TT;
    MeasP<T> p{};
    for (int pass = 0; pass < 2; ++pass) {
        Arena ar(pass ? h->ws : nullptr, pass ? h->ws_bytes : 0);
        p = MeasP<T>{};
        p.Nr = Nr; p.Nt = Nt; p.L = L; p.T_ = TT; p.Wc = Wc; p.Lr = d->Lr; p.psi_mode = d->psi_mode; p.Tp = d->Tp;
        p.ws = ar.take<cx<T>>((size_t)batch * Nr * TT);
        auto in = [&](const void* src, size_t per, long long ld, long long& ld_out) -> const cx<T>* {
            if (!src) { ld_out = 0; return nullptr; }
            if (!host) { ld_out = ld; return (const cx<T>*)src; }
            const int cnt = ld ? batch : 1;
            cx<T>* dv = ar.take<cx<T>>(per * cnt);
            if (pass) {
                if (ld == 0 || (size_t)ld == per) cudaMemcpyAsync(dv, src, per * cnt * sizeof(cx<T>), cudaMemcpyHostToDevice, h->stream);
                else cudaMemcpy2DAsync(dv, per * sizeof(cx<T>), src, (size_t)ld * sizeof(cx<T>), per * sizeof(cx<T>), cnt, cudaMemcpyHostToDevice, h->stream);
            }
            ld_out = ld ? (long long)per : 0;
            return dv;
        };
        p.H = in(H, nH, d->ld_H, p.ld_H);
        p.N = in(N, nN, d->ld_N, p.ld_N);
        p.Psi = in(Psi, nPsi, d->ld_Psi, p.ld_Psi);
        p.W = in(W, nW, d->ld_W, p.ld_W);
        const size_t nperm = (size_t)TT * Wc;
        if (perm) {
            if (host) { int* dp = ar.take<int>(nperm * batch); if (pass) cudaMemcpyAsync(dp, perm, nperm * batch * sizeof(int), cudaMemcpyHostToDevice, h->stream); p.perm = dp; }
            else p.perm = perm;
            p.ld_perm = (long long)nperm;
        }
        const size_t nY = (size_t)Wc * TT * batch, nWe = (size_t)Nr * Wc * batch, nPb = (size_t)Nt * TT * L * batch, nYn = (size_t)Nr * TT * batch;
        p.Yout = Yout ? (host ? ar.take<cx<T>>(nY) : (cx<T>*)Yout) : nullptr;
        p.We = We ? (host ? ar.take<cx<T>>(nWe) : (cx<T>*)We) : nullptr;
        p.Psibar = Psibar ? (host ? ar.take<cx<T>>(nPb) : (cx<T>*)Psibar) : nullptr;
        p.Omega = Omega ? (host ? ar.take<T>(nY) : (T*)Omega) : nullptr;
        p.Ynl = Ynl ? (host ? ar.take<cx<T>>(nYn) : (cx<T>*)Ynl) : nullptr;
        if (!pass) { int rc = ensure_workspace(h, ar.off); if (rc) return rc; continue; }
        const int RG = Nr / 8, CTc = (Nr % 8 == 0 && RG >= 1 && 256 % RG == 0) ? 256 / RG : 0;
        const size_t sm_t = CTc ? sizeof(cx<T>) * ((size_t)L * Nt * Nr + (size_t)Nr * (Wc + 1) + std::max((size_t)Nt * (CTc + L - 1), (size_t)(CTc + 1) * Nr)) : 0;
        if (CTc && d->psi_mode == 1 && !Psibar && !We && sm_t <= h->smem_optin && getenv("JSTSP_MEASURE_SIMPLE") == nullptr) {
            { int rc = set_smem(h, k_measure_tiled<T>, sm_t); if (rc) return rc; }
            dim3 grid(ceil_div(TT, CTc), batch);
            JSTSP_LAUNCH(h, PK_OTHER, (k_measure_tiled<T><<<grid, 256, sm_t, h->stream>>>(p)));
        } else {
            int tiles = (TT + 31) / 32; if (tiles > 64) tiles = 64; if (tiles < 1) tiles = 1;
            dim3 grid(tiles, batch);
            JSTSP_LAUNCH(h, PK_OTHER, (k_measure<T><<<grid, 256, 0, h->stream>>>(p)));
        }
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            if (Yout) JSTSP_CUDA(h, cudaMemcpyAsync(Yout, p.Yout, nY * sizeof(cx<T>), cudaMemcpyDeviceToHost, h->stream));
            if (We) JSTSP_CUDA(h, cudaMemcpyAsync(We, p.We, nWe * sizeof(cx<T>), cudaMemcpyDeviceToHost, h->stream));
            if (Psibar) JSTSP_CUDA(h, cudaMemcpyAsync(Psibar, p.Psibar, nPb * sizeof(cx<T>), cudaMemcpyDeviceToHost, h->stream));
            if (Omega) JSTSP_CUDA(h, cudaMemcpyAsync(Omega, p.Omega, nY * sizeof(T), cudaMemcpyDeviceToHost, h->stream));
            if (Ynl) JSTSP_CUDA(h, cudaMemcpyAsync(Ynl, p.Ynl, nYn * sizeof(cx<T>), cudaMemcpyDeviceToHost, h->stream));
            JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
        }
    }
    return JSTSP_OK;
}

// ---- combiner codebooks and the 4-QAM alphabet (the inputs the drivers prepare around the estimator) -----------
namespace jstsp {
template <typename T>
__global__ void k_beamformer(int N, int type, const int* __restrict__ draws, cx<T>* __restrict__ B) {
    const double rs = 1.0 / sqrt((double)N);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N * N; t += gridDim.x * blockDim.x) {
        const int i = t % N, j = t / N;                        // B(i, j), column-major
        double ph = 0.0;                                       // B = rs * exp(-1j * ph) except for 'rand'
        switch (type) {
            case JSTSP_BF_FFT: ph = 2.0 * M_PI * (double)((long long)i * j % N) / N; break;                       // fft(eye(N)), createBeamformer.m:6
            case JSTSP_BF_RAND: {                                                                                 // randsrc(N,N,[1 -1 1j -1j]), :8
                const int d = draws[t];
                B[t] = mk<T>((T)(d == 0 ? rs : d == 1 ? -rs : 0.0), (T)(d == 2 ? rs : d == 3 ? -rs : 0.0));
                continue;
            }
            case JSTSP_BF_RAND_PS: ph = (double)i * 2.0 * M_PI * (double)draws[j] / 32.0; break;                  // randi(32,1,N), :10-11
            case JSTSP_BF_PS: ph = (double)i * 2.0 * M_PI * (double)j / N; break;                                 // :13-14
            case JSTSP_BF_ZC: ph = 11.0 * (double)i * M_PI * (double)(j + 1) / N; break;                          // :16-17
            default: {                                                                                            // quantized_4 (:19-24) / quantized (:26-31)
                const int nq = type == JSTSP_BF_QUANTIZED_4 ? 4 : 6, levels = 1 << nq;
                ph = (double)i * (2.0 * M_PI / levels) * (double)(j % levels);                                    // A(1:N) of the tiled 0..2^nq-1
            }
        }
        double sn, cs; sincos(-ph, &sn, &cs);
        B[t] = mk<T>((T)(rs * cs), (T)(rs * sn));
    }
}
// mode 0: symbols = alphabet(draw) with alphabet = [1+1j, -1+1j, 1-1j, -1-1j]/sqrt(2) (qam4mod.m:7-8);
// mode 1: hard decision of soft symbols (qam4mod.m:12-31; MATLAB orders complex numbers by their real parts, so the tests are on re / im >= 0, <= 0,
//         applied in the reference's order s2, s3, s4 - later rules win on the axes)
template <typename T>
__global__ void k_qam4(int mode, size_t n, const int* __restrict__ draws, const cx<T>* __restrict__ in, cx<T>* __restrict__ out) {
    const T a = (T)(1.0 / sqrt(2.0));                        // as the reference forms it, (1+1j)/sqrt(2): one ulp below the correctly rounded sqrt(1/2) in fp64
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        if (mode == 0) { const int d = draws[t]; out[t] = mk<T>((d & 1) ? -a : a, (d & 2) ? -a : a); continue; }
        const cx<T> v = in[t];
        cx<T> s = mk<T>(a, a);
        if (v.re >= 0 && v.im <= 0) s = mk<T>(a, -a);
        if (v.re <= 0 && v.im >= 0) s = mk<T>(-a, a);
        if (v.re <= 0 && v.im <= 0) s = mk<T>(-a, -a);
        out[t] = s;
    }
}
template <typename T>
static int run_small(Handle* h, int mem, size_t n_out, const int* draws, size_t n_draws, const void* in, void* out, int kind, int a0, int a1) {
    const bool host = mem == JSTSP_HOST;
    Arena probe(nullptr, 0);
    probe.take<cx<T>>(n_out); probe.take<cx<T>>(n_out); probe.take<int>(n_draws ? n_draws : 1);
    int rc = ensure_workspace(h, probe.off); if (rc) return rc;
    Arena ar(h->ws, h->ws_bytes);
    cx<T>* d_out = host ? ar.take<cx<T>>(n_out) : (cx<T>*)out;
    const cx<T>* d_in = (const cx<T>*)in;
    const int* d_draws = draws;
    if (host) {
        if (in) { cx<T>* t = ar.take<cx<T>>(n_out); JSTSP_CUDA(h, cudaMemcpyAsync(t, in, n_out * sizeof(cx<T>), cudaMemcpyHostToDevice, h->stream)); d_in = t; }
        if (draws) { int* t = ar.take<int>(n_draws); JSTSP_CUDA(h, cudaMemcpyAsync(t, draws, n_draws * sizeof(int), cudaMemcpyHostToDevice, h->stream)); d_draws = t; }
    }
    const int grid = (int)((n_out + 255) / 256 < 1184 ? (n_out + 255) / 256 : 1184);
    if (kind == 0) JSTSP_LAUNCH(h, PK_OTHER, (k_beamformer<T><<<grid, 256, 0, h->stream>>>(a0, a1, d_draws, d_out)));
    else JSTSP_LAUNCH(h, PK_OTHER, (k_qam4<T><<<grid, 256, 0, h->stream>>>(a0, n_out, d_draws, d_in, d_out)));
    JSTSP_CUDA(h, cudaGetLastError());
    if (host) {
        JSTSP_CUDA(h, cudaMemcpyAsync(out, d_out, n_out * sizeof(cx<T>), cudaMemcpyDeviceToHost, h->stream));
        JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return JSTSP_OK;
}
}  // namespace jstsp

extern "C" int jstsp_create_beamformer(jstsp_handle* h, int dtype, int mem, int N, int type, const int* draws, void* B) {
    if (!h) return JSTSP_E_ARG;
    if (N <= 0 || !B || type < JSTSP_BF_FFT || type > JSTSP_BF_QUANTIZED) return fail(h, JSTSP_E_ARG, "createBeamformer: bad size or unknown beamformer_type");
    if ((type == JSTSP_BF_RAND || type == JSTSP_BF_RAND_PS) && !draws) return fail(h, JSTSP_E_ARG, "createBeamformer: this codebook needs its random draws");
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    const size_t nd = type == JSTSP_BF_RAND ? (size_t)N * N : type == JSTSP_BF_RAND_PS ? (size_t)N : 0;
    if (dtype == JSTSP_F32) return run_small<float>(h, mem, (size_t)N * N, nd ? draws : nullptr, nd, nullptr, B, 0, N, type);
    if (dtype == JSTSP_F64) return run_small<double>(h, mem, (size_t)N * N, nd ? draws : nullptr, nd, nullptr, B, 0, N, type);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

extern "C" int jstsp_qam4mod(jstsp_handle* h, int dtype, int mem, int mode, long long n, const int* draws, const void* input, void* symbols) {
    if (!h) return JSTSP_E_ARG;
    if (n <= 0 || !symbols || (mode == 0 && !draws) || (mode == 1 && !input) || mode < 0 || mode > 1) return fail(h, JSTSP_E_ARG, "qam4mod: bad argument");
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_small<float>(h, mem, (size_t)n, mode == 0 ? draws : nullptr, mode == 0 ? (size_t)n : 0, mode == 1 ? input : nullptr, symbols, 1, mode, 0);
    if (dtype == JSTSP_F64) return run_small<double>(h, mem, (size_t)n, mode == 0 ? draws : nullptr, mode == 0 ? (size_t)n : 0, mode == 1 ? input : nullptr, symbols, 1, mode, 0);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

extern "C" int jstsp_measure(jstsp_handle* h, const jstsp_meas_desc* d, int dtype, int mem,
                             const void* H, const void* N, const void* Psi, const void* W, const int* perm,
                             void* Y_out, void* W_e, void* Psi_bar, void* Omega, void* Y_noiseless) {
    if (!h) return JSTSP_E_ARG;
    if (!d) return fail(h, JSTSP_E_ARG, "NULL descriptor");
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_measure<float>(h, mem, d, H, N, Psi, W, perm, Y_out, W_e, Psi_bar, Omega, Y_noiseless);
    if (dtype == JSTSP_F64) return run_measure<double>(h, mem, d, H, N, Psi, W, perm, Y_out, W_e, Psi_bar, Omega, Y_noiseless);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

template <typename T>
static int run_nmse(Handle* h, int mem, int G, int P, int batch, const void* S, long long ld_S, const void* Z, long long ld_Z, double* out) {
    if (G <= 0 || P <= 0 || batch <= 0 || !S || !Z || !out) return fail(h, JSTSP_E_ARG, "bad argument");
    if (G > 64) return fail(h, JSTSP_E_UNSUPPORTED, "nmse kernel covers <= 64 rows");
    const bool host = mem == JSTSP_HOST;
    const size_t GP = (size_t)G * P;
    if (!ld_S) ld_S = GP; if (!ld_Z) ld_Z = GP;
    const size_t smem = JacobiSmem::bytes(G);
    int rc = set_smem(h, k_nmse<T>, smem); if (rc) return rc;
    const cx<T>* ds = (const cx<T>*)S; const cx<T>* dz = (const cx<T>*)Z; double* dout = out;
    if (host) {
        rc = ensure_workspace(h, 2 * (GP * batch * sizeof(cx<T>) + 256) + batch * sizeof(double) + 256); if (rc) return rc;
        Arena ar(h->ws, h->ws_bytes);
        cx<T>* a = ar.take<cx<T>>(GP * batch); cx<T>* b = ar.take<cx<T>>(GP * batch); dout = ar.take<double>(batch);
        JSTSP_CUDA(h, cudaMemcpy2DAsync(a, GP * sizeof(cx<T>), S, (size_t)ld_S * sizeof(cx<T>), GP * sizeof(cx<T>), batch, cudaMemcpyHostToDevice, h->stream));
        JSTSP_CUDA(h, cudaMemcpy2DAsync(b, GP * sizeof(cx<T>), Z, (size_t)ld_Z * sizeof(cx<T>), GP * sizeof(cx<T>), batch, cudaMemcpyHostToDevice, h->stream));
        ds = a; dz = b; ld_S = GP; ld_Z = GP;
    }
    JSTSP_LAUNCH(h, PK_OTHER, (k_nmse<T><<<batch, 256, smem, h->stream>>>(ds, ld_S, dz, ld_Z, G, P, dout)));
    JSTSP_CUDA(h, cudaGetLastError());
    if (host) { JSTSP_CUDA(h, cudaMemcpyAsync(out, dout, batch * sizeof(double), cudaMemcpyDeviceToHost, h->stream)); JSTSP_CUDA(h, cudaStreamSynchronize(h->stream)); }
    return JSTSP_OK;
}

extern "C" int jstsp_nmse(jstsp_handle* h, int dtype, int mem, int G, int P, int batch,
                          const void* S, long long ld_S, const void* Zbar, long long ld_Z, double* nmse) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_nmse<float>(h, mem, G, P, batch, S, ld_S, Zbar, ld_Z, nmse);
    if (dtype == JSTSP_F64) return run_nmse<double>(h, mem, G, P, batch, S, ld_S, Zbar, ld_Z, nmse);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}


namespace jstsp {
// ---------------------------------------------------------------------------------------------
// rate metric: log2 det(I + c X X^H)   
// ---------------------------------------------------------------------------------------------
// The achievable-rate / capacity sweeps evaluate real(log2(det(eye(n) + c * X * X'))) per trial:
//   plot_rateVSframelength.m:113,130,135   X = Zbar (Nr x L*Gt),   c = 1/(Nr (sigma^2 + NMSE))
//   plot_capacity.m:47-66, plot_ee.m        X = W_c' * Y (Mr x T), c = 1/(sigma^2 Nt)
// Computed through the eigenvalues of the n x n Gram matrix (same fp64 Jacobi solver as the SVT): sum_k log2(1 + c lambda_k).
template <typename T>
__global__ void __launch_bounds__(256) k_log2det(const cx<T>* X, long long ld_X, int n, int m, const double* scale, double* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    JacobiSmem sm; sm.carve(smem, n);
    const int b = blockIdx.x;
    gram_rows<T>(sm, X + (long long)b * ld_X, nullptr, n, m);
    jacobi_hermitian_block(sm, n);
    if (threadIdx.x == 0) {
        const double c = scale[b];
        double acc = 0.0;
        for (int k = 0; k < n; ++k) acc += log2(1.0 + c * fmax(sm.Are[k + n * k], 0.0));
        out[b] = acc;
    }
}

template <typename T>
static int run_log2det(Handle* h, int mem, int n, int m, int batch, const void* X, long long ld_X, const double* scale, double* out) {
    if (n <= 0 || m <= 0 || batch <= 0 || !X || !scale || !out) return fail(h, JSTSP_E_ARG, "bad argument");
    if (n > 64) return fail(h, JSTSP_E_UNSUPPORTED, "rate kernel covers <= 64 rows");
    const bool host = mem == JSTSP_HOST;
    const size_t NM = (size_t)n * m;
    if (!ld_X) ld_X = NM;
    const size_t smem = JacobiSmem::bytes(n);
    int rc = set_smem(h, k_log2det<T>, smem); if (rc) return rc;
    const cx<T>* dx = (const cx<T>*)X; const double* ds = scale; double* dout = out;
    if (host) {
        rc = ensure_workspace(h, NM * batch * sizeof(cx<T>) + 2 * batch * sizeof(double) + 2048); if (rc) return rc;
        Arena ar(h->ws, h->ws_bytes);
        cx<T>* a = ar.take<cx<T>>(NM * batch);
        double* s2 = ar.take<double>(batch); double* o2 = ar.take<double>(batch);
        JSTSP_CUDA(h, cudaMemcpy2DAsync(a, NM * sizeof(cx<T>), X, (size_t)ld_X * sizeof(cx<T>), NM * sizeof(cx<T>), batch, cudaMemcpyHostToDevice, h->stream));
        JSTSP_CUDA(h, cudaMemcpyAsync(s2, scale, sizeof(double) * batch, cudaMemcpyHostToDevice, h->stream));
        dx = a; ld_X = NM; ds = s2; dout = o2;
    }
    JSTSP_LAUNCH(h, PK_OTHER, (k_log2det<T><<<batch, 256, smem, h->stream>>>(dx, ld_X, n, m, ds, dout)));
    JSTSP_CUDA(h, cudaGetLastError());
    if (host) {
        JSTSP_CUDA(h, cudaMemcpyAsync(out, dout, sizeof(double) * batch, cudaMemcpyDeviceToHost, h->stream));
        JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return JSTSP_OK;
}

// X = W(:, cols)' * Y  (Mr x T): the combiner columns a design keeps, applied to the noiseless received block
//   hbf.m:24 (W_c = W(:, 1:Lr)), plot_capacity.m:63-64 / plot_ee.m:63-64 (W(:, ind(1:Mr)), ind = randperm(Mr_e))
template <typename T>
__global__ void __launch_bounds__(256) k_select_combine(const cx<T>* Y, long long ld_Y, const cx<T>* W, long long ld_W, const int* cols, long long ld_cols,
                                                        int Nr, int T_, int Mr, cx<T>* X) {
    const int b = blockIdx.x;
    const cx<T>* y = Y + (long long)b * ld_Y; const cx<T>* w = W + (long long)b * ld_W;
    const int* cs = cols ? cols + (long long)b * ld_cols : nullptr;
    for (int e = threadIdx.x; e < Mr * T_; e += blockDim.x) {
        const int r = e % Mr, t = e / Mr, c = cs ? cs[r] - 1 : r;                 // cols are 1-based like ind(1:Mr)
        T re = 0, im = 0;
        for (int n = 0; n < Nr; ++n) { const cx<T> a = w[n + (size_t)Nr * c], v = y[n + (size_t)Nr * t]; cmac<T>(re, im, a.re, -a.im, v.re, v.im); }
        X[(size_t)b * Mr * T_ + e] = mk<T>(re, im);
    }
}

template <typename T>
static int run_capacity(Handle* h, int mem, int Nr, int T_, int Wc, int Mr, int batch, const void* Y, long long ld_Y, const void* W, long long ld_W,
                        const int* cols, long long ld_cols, const double* scale, double* out) {
    if (Nr <= 0 || T_ <= 0 || Mr <= 0 || Wc < Mr || batch <= 0 || !Y || !W || !scale || !out) return fail(h, JSTSP_E_ARG, "bad argument");
    if (Mr > 64) return fail(h, JSTSP_E_UNSUPPORTED, "rate kernel covers <= 64 rows");
    const bool host = mem == JSTSP_HOST;
    const size_t NT = (size_t)Nr * T_, NW = (size_t)Nr * Wc, MT = (size_t)Mr * T_, esz = sizeof(cx<T>);
    if (!ld_Y) ld_Y = NT;
    const int nW = ld_W ? batch : 1;
    if (cols && !ld_cols && batch > 1) ld_cols = 0;
    int rc = ensure_workspace(h, esz * (MT * batch + (host ? NT * batch + NW * nW : 0)) + (host ? sizeof(int) * (size_t)Mr * batch + 2 * sizeof(double) * batch : 0) + 8192); if (rc) return rc;
    Arena ar(h->ws, h->ws_bytes);
    cx<T>* X = ar.take<cx<T>>(MT * batch);
    const cx<T>* dY = (const cx<T>*)Y; const cx<T>* dW = (const cx<T>*)W; const int* dc = cols; const double* ds = scale; double* dout = out;
    if (host) {
        cx<T>* a = ar.take<cx<T>>(NT * batch); cx<T>* w = ar.take<cx<T>>(NW * nW);
        JSTSP_CUDA(h, cudaMemcpy2DAsync(a, NT * esz, Y, (size_t)ld_Y * esz, NT * esz, batch, cudaMemcpyHostToDevice, h->stream));
        JSTSP_CUDA(h, cudaMemcpy2DAsync(w, NW * esz, W, (size_t)(ld_W ? ld_W : NW) * esz, NW * esz, nW, cudaMemcpyHostToDevice, h->stream));
        dY = a; ld_Y = NT; dW = w; if (ld_W) ld_W = NW;
        if (cols) {
            const int nc = ld_cols ? batch : 1;
            int* c2 = ar.take<int>((size_t)Mr * nc);
            JSTSP_CUDA(h, cudaMemcpy2DAsync(c2, sizeof(int) * Mr, cols, sizeof(int) * (size_t)(ld_cols ? ld_cols : Mr), sizeof(int) * Mr, nc, cudaMemcpyHostToDevice, h->stream));
            dc = c2; if (ld_cols) ld_cols = Mr;
        }
        double* s2 = ar.take<double>(batch); double* o2 = ar.take<double>(batch);
        JSTSP_CUDA(h, cudaMemcpyAsync(s2, scale, sizeof(double) * batch, cudaMemcpyHostToDevice, h->stream));
        ds = s2; dout = o2;
    }
    JSTSP_LAUNCH(h, PK_OTHER, (k_select_combine<T><<<batch, 256, 0, h->stream>>>(dY, ld_Y, dW, ld_W, dc, ld_cols, Nr, T_, Mr, X)));
    const size_t smem = JacobiSmem::bytes(Mr);
    rc = set_smem(h, k_log2det<T>, smem); if (rc) return rc;
    JSTSP_LAUNCH(h, PK_OTHER, (k_log2det<T><<<batch, 256, smem, h->stream>>>(X, (long long)MT, Mr, T_, ds, dout)));
    JSTSP_CUDA(h, cudaGetLastError());
    if (host) {
        JSTSP_CUDA(h, cudaMemcpyAsync(out, dout, sizeof(double) * batch, cudaMemcpyDeviceToHost, h->stream));
        JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return JSTSP_OK;
}

}  // namespace jstsp
using namespace jstsp;
extern "C" int jstsp_log2det_rate(jstsp_handle* h, int dtype, int mem, int n, int m, int batch,
                                  const void* X, long long ld_X, const double* scale, double* rate) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_log2det<float>(h, mem, n, m, batch, X, ld_X, scale, rate);
    if (dtype == JSTSP_F64) return run_log2det<double>(h, mem, n, m, batch, X, ld_X, scale, rate);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

extern "C" int jstsp_capacity(jstsp_handle* h, int dtype, int mem, int Nr, int T, int Wc, int Mr, int batch,
                              const void* Y, long long ld_Y, const void* W, long long ld_W, const int* cols, long long ld_cols,
                              const double* scale, double* rate) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_capacity<float>(h, mem, Nr, T, Wc, Mr, batch, Y, ld_Y, W, ld_W, cols, ld_cols, scale, rate);
    if (dtype == JSTSP_F64) return run_capacity<double>(h, mem, Nr, T, Wc, Mr, batch, Y, ld_Y, W, ld_W, cols, ld_cols, scale, rate);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

// One call for the loop body of plot_capacity.m:35-66 / plot_ee.m:36-66 over an Mr range: the four receiver designs on the same noiseless blocks.
//   design 0 digital BF (all Nr columns of W_zc), 1 conventional HBF with phase shifters (W_q(:, 1:Mr)), 2 conventional HBF with ZC (W_zc(:, 1:Mr)),
//   3 proposed (W_q(:, ind(1:Mr)), ind = randperm(Mr_e) per trial).   out[(i_mr * 4 + design) * batch + b]
template <typename T>
static int run_capacity_sweep(Handle* h, int mem, int Nr, int T_, int batch, int n_mr, const int* mr_range, const void* Y, long long ld_Y, const void* Wzc, const void* Wq,
                              const int* ind, long long ld_ind, const double* scale, double* out) {
    if (n_mr <= 0 || !mr_range || !out || !ind) return fail(h, JSTSP_E_ARG, "bad argument");
    for (int i = 0; i < n_mr; ++i) {
        const int Mr = mr_range[i];
        if (Mr <= 0 || Mr > Nr) return fail(h, JSTSP_E_ARG, "Mr outside 1..Nr");
        double* o = out + (size_t)i * 4 * batch;
        int rc;
        if ((rc = run_capacity<T>(h, mem, Nr, T_, Nr, Nr, batch, Y, ld_Y, Wzc, 0, nullptr, 0, scale, o))) return rc;                       // plot_capacity.m:45-47
        if ((rc = run_capacity<T>(h, mem, Nr, T_, Nr, Mr, batch, Y, ld_Y, Wq, 0, nullptr, 0, scale, o + batch))) return rc;                // :50-52
        if ((rc = run_capacity<T>(h, mem, Nr, T_, Nr, Mr, batch, Y, ld_Y, Wzc, 0, nullptr, 0, scale, o + 2 * (size_t)batch))) return rc;   // :55-57
        if ((rc = run_capacity<T>(h, mem, Nr, T_, Nr, Mr, batch, Y, ld_Y, Wq, 0, ind, ld_ind, scale, o + 3 * (size_t)batch))) return rc;   // :61-64
    }
    return JSTSP_OK;
}
extern "C" int jstsp_capacity_sweep(jstsp_handle* h, int dtype, int mem, int Nr, int T, int batch, int n_mr, const int* mr_range,
                                    const void* Y, long long ld_Y, const void* W_zc, const void* W_q, const int* ind, long long ld_ind,
                                    const double* scale, double* out) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_capacity_sweep<float>(h, mem, Nr, T, batch, n_mr, mr_range, Y, ld_Y, W_zc, W_q, ind, ld_ind, scale, out);
    if (dtype == JSTSP_F64) return run_capacity_sweep<double>(h, mem, Nr, T, batch, n_mr, mr_range, Y, ld_Y, W_zc, W_q, ind, ld_ind, scale, out);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

template <typename T>
static int run_params(Handle* h, int mem, int N, int M, int G, int P, int batch, int kth, const void* Y, long long ld_Y, const void* Z, long long ld_Z,
                      double* tauY, double* tauZ, double* rho) {
    if (N <= 0 || M <= 0 || batch <= 0 || kth < 1 || !Y || !tauY || !rho) return fail(h, JSTSP_E_ARG, "bad argument");
    if (N > 64) return fail(h, JSTSP_E_UNSUPPORTED, "parameter kernel covers <= 64 rows");
    const bool host = mem == JSTSP_HOST;
    const size_t NM = (size_t)N * M, GP = (size_t)G * P;
    if (!ld_Y) ld_Y = NM; if (!ld_Z) ld_Z = GP;
    const size_t smem = JacobiSmem::bytes(N);
    int rc = set_smem(h, k_params<T>, smem); if (rc) return rc;
    const cx<T>* dy = (const cx<T>*)Y; const cx<T>* dz = (const cx<T>*)Z; double *t1 = tauY, *t2 = tauZ, *t3 = rho;
    if (host) {
        rc = ensure_workspace(h, (NM + GP) * batch * sizeof(cx<T>) + 3 * batch * sizeof(double) + 2048); if (rc) return rc;
        Arena ar(h->ws, h->ws_bytes);
        cx<T>* a = ar.take<cx<T>>(NM * batch);
        JSTSP_CUDA(h, cudaMemcpy2DAsync(a, NM * sizeof(cx<T>), Y, (size_t)ld_Y * sizeof(cx<T>), NM * sizeof(cx<T>), batch, cudaMemcpyHostToDevice, h->stream));
        dy = a; ld_Y = NM;
        if (Z) {
            cx<T>* b = ar.take<cx<T>>(GP * batch);
            JSTSP_CUDA(h, cudaMemcpy2DAsync(b, GP * sizeof(cx<T>), Z, (size_t)ld_Z * sizeof(cx<T>), GP * sizeof(cx<T>), batch, cudaMemcpyHostToDevice, h->stream));
            dz = b; ld_Z = GP;
        }
        t1 = ar.take<double>(batch); t2 = ar.take<double>(batch); t3 = ar.take<double>(batch);
    }
    JSTSP_LAUNCH(h, PK_OTHER, (k_params<T><<<batch, 256, smem, h->stream>>>(dy, ld_Y, N, M, dz, ld_Z, G, P, kth, t1, (Z && tauZ) ? t2 : nullptr, t3)));
    JSTSP_CUDA(h, cudaGetLastError());
    if (host) {
        JSTSP_CUDA(h, cudaMemcpyAsync(tauY, t1, batch * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (Z && tauZ) JSTSP_CUDA(h, cudaMemcpyAsync(tauZ, t2, batch * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        JSTSP_CUDA(h, cudaMemcpyAsync(rho, t3, batch * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return JSTSP_OK;
}

extern "C" int jstsp_admm_parameters(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int batch, int kth_eig,
                                     const void* Y, long long ld_Y, const void* Zbar, long long ld_Z,
                                     double* tau_Y, double* tau_Z, double* rho) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_params<float>(h, mem, N, M, G, P, batch, kth_eig, Y, ld_Y, Zbar, ld_Z, tau_Y, tau_Z, rho);
    if (dtype == JSTSP_F64) return run_params<double>(h, mem, N, M, G, P, batch, kth_eig, Y, ld_Y, Zbar, ld_Z, tau_Y, tau_Z, rho);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
