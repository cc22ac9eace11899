// jacobi.cuh - block-level Hermitian eigen-solver (parallel two-sided Jacobi, fp64) used
// for the singular-value thresholding of a short-and-wide matrix Z through its small
// Gram matrix G = Z Z^H: svt(Z,tau) = U diag(max(0,1-tau/sigma)) U^H Z.
// Replaces the full `svd` of benchmark_algorithms/svt.m:5 - the output of svt.m is basis
// independent, so the n x n Hermitian problem (n <= 64) is all that is needed.
#pragma once
#include "common.cuh"

namespace jstsp {

struct JacobiSmem {
    double *Are, *Aim, *Ure, *Uim;     // n*n each, column-major
    double *c, *s, *er, *ei, *offacc;  // npad/2 each
    int *pp, *qq;                      // npad/2 each
    double* red;                       // 4 doubles: fro2, off2, converged flag, spare
    __host__ __device__ static size_t bytes(int n) {
        int h = (n + 1) / 2;
        return sizeof(double) * (4 * (size_t)n * n + 5 * (size_t)h + 4) + sizeof(int) * 2 * (size_t)h;
    }
    __device__ void carve(void* base, int n) {
        int h = (n + 1) / 2;
        double* p = reinterpret_cast<double*>(base);
        Are = p; p += n * n; Aim = p; p += n * n; Ure = p; p += n * n; Uim = p; p += n * n;
        c = p; p += h; s = p; p += h; er = p; p += h; ei = p; p += h; offacc = p; p += h;
        red = p; p += 4;
        pp = reinterpret_cast<int*>(p); qq = pp + h;
    }
};

// Fast fp64 reciprocal square root / reciprocal: fp32 hardware seed + Newton steps in fp64 (full
// double accuracy after two/three steps).  The argument is range-reduced through its exponent so
// the fp32 seed never over/underflows.  x > 0 and finite.
__device__ __forceinline__ double pow2i(int k) { return __hiloint2double((1023 + k) << 20, 0); }   // exact 2^k, -1022 <= k <= 1023
__device__ __forceinline__ double fast_rsqrt(double x) {
    const int ex = ((__double2hiint(x) >> 20) & 0x7ff) - 1023;
    const int hshift = ex >> 1;
    const double xs = x * pow2i(-2 * hshift);                 // in [1, 4); one exact multiply instead of scalbn's library path
    double r = (double)rsqrtf((float)xs);                     // relative error <= 2^-22: two Newton steps (e <- 1.5 e^2) reach 1e-26
    r = r * (1.5 - 0.5 * xs * r * r);
    r = r * (1.5 - 0.5 * xs * r * r);
    return r * pow2i(-hshift);
}
__device__ __forceinline__ double fast_rcp_ge1(double d) {     // d >= 1 ; huge d (> fp32 range) returns 0
    double r = (double)__frcp_rn((float)d);                   // relative error <= 2^-24: two Newton steps (e <- e^2) reach 1e-29
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    return r;
}

// Round-robin ("circle method") pairing: npad players, step s in [0, npad-1).
__device__ __forceinline__ void rr_pair(int npad, int s, int k, int& p, int& q) {
    int a, b;
    const int r = npad - 1;
    if (k == 0) { a = r; b = s; }
    else { a = s + k; if (a >= r) a -= r; b = s - k + r; if (b >= r) b -= r; }      // (s +- k) mod (npad - 1) without a division
    p = a < b ? a : b; q = a < b ? b : a;
}

// In: Hermitian A (sm.Are/Aim).  Out: eigenvalues on the diagonal of A, eigenvectors in U.
// All threads of the block must call; uses __syncthreads().
// keep_U: U already holds a unitary matrix Q and A holds Q^H G Q (warm start); the rotations are
// accumulated onto Q so that U ends as the eigenvector matrix of G.
// stop_rel2: the iteration stops after a sweep that STARTED with off-diagonal mass <= stop_rel2 * |A|_F^2 (quadratic convergence:
// that sweep ends near the square of its starting level).
__device__ inline int jacobi_hermitian_block(JacobiSmem& sm, int n, int max_sweeps = 24, bool keep_U = false, double stop_rel2 = 1e-20) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int npad = n + (n & 1), h = npad / 2;
    const bool fast_t = stop_rel2 > 1e-15;          // the fp32 callers' tolerance
    if (!keep_U) for (int i = tid; i < n * n; i += nt) { sm.Ure[i] = (i % n == i / n) ? 1.0 : 0.0; sm.Uim[i] = 0.0; }
    {   // Frobenius norm^2 (block reduction; deterministic order)
        double f = 0.0;
        for (int i = tid; i < n * n; i += nt) f += sm.Are[i] * sm.Are[i] + sm.Aim[i] * sm.Aim[i];
        for (int o = 16; o > 0; o >>= 1) f += __shfl_down_sync(0xffffffffu, f, o);
        if (tid == 0) { sm.red[0] = 0.0; sm.red[2] = 0.0; }
        __syncthreads();
        // warps add their partials one after the other (fixed order)
        for (int w = 0; w < (nt + 31) / 32; ++w) {
            if (tid == w * 32) sm.red[0] += f;
            __syncthreads();
        }
    }
    if (n < 2 || sm.red[0] == 0.0) return 0;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (keep_U || sweep > 0) {
            // Is another sweep worth it?  Off-diagonal mass now (one block reduction) against stop_rel2^1.5 |A|_F^2: a warm-started or
            // already-swept matrix whose relative off-diagonal NORM is below stop_rel2^0.75 (3e-8 for the fp32 solves, 1e-15 for fp64) has
            // eigenvectors good to that level, and the sweep that would confirm it costs as much as the ones that did the work.
            double o2 = 0.0;
            for (int i = tid; i < n * n; i += nt) if (i % n != i / n) o2 += sm.Are[i] * sm.Are[i] + sm.Aim[i] * sm.Aim[i];
            for (int o = 16; o > 0; o >>= 1) o2 += __shfl_down_sync(0xffffffffu, o2, o);
            if (tid == 0) sm.red[1] = 0.0;
            __syncthreads();
            for (int w = 0; w < (nt + 31) / 32; ++w) {
                if (tid == w * 32) sm.red[1] += o2;
                __syncthreads();
            }
            if (sm.red[1] <= stop_rel2 * sqrt(stop_rel2) * sm.red[0]) break;
        }
        if (tid < h) sm.offacc[tid] = 0.0;
        for (int step = 0; step < npad - 1; ++step) {
            // phase 0: rotation parameters, one thread per pair
            if (tid < h) {
                int p, q; rr_pair(npad, step, tid, p, q);
                sm.pp[tid] = p; sm.qq[tid] = q;
                double c = 1.0, s = 0.0, er = 1.0, ei = 0.0;
                if (q < n) {
                    double ar = sm.Are[p + n * q], ai = sm.Aim[p + n * q];
                    double app = sm.Are[p + n * p], aqq = sm.Are[q + n * q];
                    double m2 = ar * ar + ai * ai;
                    if (m2 > 1e-290 && m2 > 1e-36 * fabs(app * aqq)) {
                        sm.offacc[tid] += m2;
                        const double rm = fast_rsqrt(m2);              // 1/|a_pq|
                        er = ar * rm; ei = ai * rm;
                        const double tau = (aqq - app) * 0.5 * rm;
                        double t;
                        if (fast_t) {
                            // fp32 solves: the tangent to single precision (|tau| beyond the float range gives t = 0, its limit).  (c, s, e) below still
                            // form an exactly unitary rotation in fp64; it merely leaves ~1e-7 |a_pq| behind, far under what the next sweep removes.
                            const float tf = (float)tau;
                            t = (double)(copysignf(1.f, tf) / (fabsf(tf) + sqrtf(fmaf(tf, tf, 1.f))));
                        } else {
                            const double s1 = 1.0 + tau * tau;
                            const double w = s1 < 1e300 ? s1 * fast_rsqrt(s1) : fabs(tau);     // sqrt(1 + tau^2)
                            t = (tau >= 0.0 ? 1.0 : -1.0) * fast_rcp_ge1(fabs(tau) + w);
                        }
                        c = fast_rsqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                sm.c[tid] = c; sm.s[tid] = s; sm.er[tid] = er; sm.ei[tid] = ei;
            }
            __syncthreads();
            // phase 1: every 2x2 block (k1,k2) <- J1^H * block * J2 ; U(:, {p,q}) <- U(:, {p,q}) * J
            for (int t = tid; t < h * h; t += nt) {
                int k1 = t % h, k2 = t / h;          // row pairs vary fastest across a warp: the column-major accesses below hit distinct banks
                                                       // (with k2 fastest all lanes share a row and the stride-n columns collide 16-way at n = 32)
                int p1 = sm.pp[k1], q1 = sm.qq[k1], p2 = sm.pp[k2], q2 = sm.qq[k2];
                bool v1 = q1 < n, v2 = q2 < n;
                double c1 = sm.c[k1], s1 = sm.s[k1], e1r = sm.er[k1], e1i = sm.ei[k1];
                double c2 = sm.c[k2], s2 = sm.s[k2], e2r = sm.er[k2], e2i = sm.ei[k2];
                if (s1 == 0.0 && s2 == 0.0) continue;
                // block entries b[r][c]
                double br[2][2] = {{0, 0}, {0, 0}}, bi[2][2] = {{0, 0}, {0, 0}};
                br[0][0] = sm.Are[p1 + n * p2]; bi[0][0] = sm.Aim[p1 + n * p2];
                if (v2) { br[0][1] = sm.Are[p1 + n * q2]; bi[0][1] = sm.Aim[p1 + n * q2]; }
                if (v1) { br[1][0] = sm.Are[q1 + n * p2]; bi[1][0] = sm.Aim[q1 + n * p2]; }
                if (v1 && v2) { br[1][1] = sm.Are[q1 + n * q2]; bi[1][1] = sm.Aim[q1 + n * q2]; }
                // rows: new_p = c*row_p - s*e*row_q ; new_q = s*row_p + c*e*row_q   (J^H on the left)
                if (s1 != 0.0) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        double xr = br[0][cc], xi = bi[0][cc], yr = br[1][cc], yi = bi[1][cc];
                        double eyr = e1r * yr - e1i * yi, eyi = e1r * yi + e1i * yr;   // e * y
                        br[0][cc] = c1 * xr - s1 * eyr; bi[0][cc] = c1 * xi - s1 * eyi;
                        br[1][cc] = s1 * xr + c1 * eyr; bi[1][cc] = s1 * xi + c1 * eyi;
                    }
                }
                // cols: new_p = c*col_p - s*conj(e)*col_q ; new_q = s*col_p + c*conj(e)*col_q  (J on the right)
                if (s2 != 0.0) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        double xr = br[rr][0], xi = bi[rr][0], yr = br[rr][1], yi = bi[rr][1];
                        double eyr = e2r * yr + e2i * yi, eyi = e2r * yi - e2i * yr;   // conj(e) * y
                        br[rr][0] = c2 * xr - s2 * eyr; bi[rr][0] = c2 * xi - s2 * eyi;
                        br[rr][1] = s2 * xr + c2 * eyr; bi[rr][1] = s2 * xi + c2 * eyi;
                    }
                }
                if (k1 == k2) {   // the annihilated pair: exact zeros off-diagonal (what is left of them with fp32 angles stays), real diagonal
                    if (!fast_t) { br[0][1] = bi[0][1] = br[1][0] = bi[1][0] = 0.0; }
                    else { br[1][0] = br[0][1]; bi[1][0] = -bi[0][1]; }
                    bi[0][0] = 0.0; bi[1][1] = 0.0;
                }
                sm.Are[p1 + n * p2] = br[0][0]; sm.Aim[p1 + n * p2] = bi[0][0];
                if (v2) { sm.Are[p1 + n * q2] = br[0][1]; sm.Aim[p1 + n * q2] = bi[0][1]; }
                if (v1) { sm.Are[q1 + n * p2] = br[1][0]; sm.Aim[q1 + n * p2] = bi[1][0]; }
                if (v1 && v2) { sm.Are[q1 + n * q2] = br[1][1]; sm.Aim[q1 + n * q2] = bi[1][1]; }
            }
            for (int t = tid; t < h * n; t += nt) {
                int k = t / n, i = t % n;
                double s2 = sm.s[k];
                if (s2 == 0.0) continue;
                int p = sm.pp[k], q = sm.qq[k];
                double c2 = sm.c[k], e2r = sm.er[k], e2i = sm.ei[k];
                double xr = sm.Ure[i + n * p], xi = sm.Uim[i + n * p], yr = sm.Ure[i + n * q], yi = sm.Uim[i + n * q];
                double eyr = e2r * yr + e2i * yi, eyi = e2r * yi - e2i * yr;
                sm.Ure[i + n * p] = c2 * xr - s2 * eyr; sm.Uim[i + n * p] = c2 * xi - s2 * eyi;
                sm.Ure[i + n * q] = s2 * xr + c2 * eyr; sm.Uim[i + n * q] = s2 * xi + c2 * eyi;
            }
            __syncthreads();
        }
        if (tid == 0) {
            double off = 0.0;
            for (int k = 0; k < h; ++k) off += sm.offacc[k];
            // `off` is the off-diagonal mass seen BEFORE this sweep's rotations; Jacobi converges
            // quadratically, so a sweep that started below 1e-10 (relative) ends at rounding level.
            sm.red[2] = (off <= stop_rel2 * sm.red[0]) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (sm.red[2] != 0.0) { ++sweep; break; }
    }
    return sweep;
}

// Warm start: A <- Q^H A Q with Q = U (n x n, in sm.U), using `tmp` (2*n*n doubles of shared memory).
__device__ inline void jacobi_similarity_block(JacobiSmem& sm, int n, double* tmp) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double* Tre = tmp; double* Tim = tmp + n * n;
    for (int t = tid; t < n * n; t += nt) {          // T = A Q
        const int i = t % n, j = t / n;
        double re = 0.0, im = 0.0;
        for (int k = 0; k < n; ++k) {
            const double ar = sm.Are[i + n * k], ai = sm.Aim[i + n * k], qr = sm.Ure[k + n * j], qi = sm.Uim[k + n * j];
            re += ar * qr - ai * qi; im += ar * qi + ai * qr;
        }
        Tre[t] = re; Tim[t] = im;
    }
    __syncthreads();
    for (int t = tid; t < n * n; t += nt) {          // A = Q^H T
        const int i = t % n, j = t / n;
        double re = 0.0, im = 0.0;
        for (int kk = 0; kk < n; ++kk) {
            int k = kk + i; if (k >= n) k -= n;      // skewed start: lanes read Q(k, i) for consecutive i - with a common k that is a stride-n (same bank) access
            const double qr = sm.Ure[k + n * i], qi = -sm.Uim[k + n * i], tr = Tre[k + n * j], ti = Tim[k + n * j];
            re += qr * tr - qi * ti; im += qr * ti + qi * tr;
        }
        sm.Are[t] = re; sm.Aim[t] = im;
    }
    __syncthreads();
    for (int t = tid; t < n * n; t += nt) {          // enforce exact Hermitian symmetry
        const int i = t % n, j = t / n;
        if (i < j) {
            const double re = 0.5 * (sm.Are[i + n * j] + sm.Are[j + n * i]), im = 0.5 * (sm.Aim[i + n * j] - sm.Aim[j + n * i]);
            sm.Are[i + n * j] = re; sm.Aim[i + n * j] = im; sm.Are[j + n * i] = re; sm.Aim[j + n * i] = -im;
        } else if (i == j) sm.Aim[t] = 0.0;
    }
    __syncthreads();
}

// Spectral weights of svt.m:7 applied on the left: W = U diag(f) U^H with
// f_k = max(0, 1 - tau/sigma_k), sigma_k = sqrt(lambda_k).  The reference returns an
// all-zero matrix when a singular value is exactly zero (svt.m:7-13, hit on the first ADMM
// iterate where the input is all-zero); here that rule fires when the Gram matrix is
// exactly zero.  Writes W (n x n, column-major, interleaved) through `store(i, j, re, im)`.
template <typename Store>
__device__ inline void svt_weights_block(JacobiSmem& sm, int n, double tau, Store store) {
    const int tid = threadIdx.x, nt = blockDim.x;
    bool zero_in = (sm.red[0] == 0.0);
    __syncthreads();
    // f_k into sm.c (reuse; h >= ... not enough for n entries) -> use offacc? sizes are npad/2; store f in Aim diag instead
    for (int k = tid; k < n; k += nt) {
        double lam = sm.Are[k + n * k];
        double sig = lam > 0.0 ? sqrt(lam) : 0.0;
        double f = (sig > tau) ? (1.0 - tau / sig) : 0.0;
        sm.Aim[k + n * k] = zero_in ? 0.0 : f;
    }
    __syncthreads();
    for (int t = tid; t < n * n; t += nt) {
        int i = t % n, j = t / n;
        double wr = 0.0, wi = 0.0;
        for (int k = 0; k < n; ++k) {
            double f = sm.Aim[k + n * k];
            double ar = sm.Ure[i + n * k], ai = sm.Uim[i + n * k];
            double br = sm.Ure[j + n * k], bi = -sm.Uim[j + n * k];   // conj(U[j,k])
            wr += f * (ar * br - ai * bi);
            wi += f * (ar * bi + ai * br);
        }
        store(i, j, wr, wi);
    }
}

}  // namespace jstsp
