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

// Round-robin ("circle method") pairing: npad players, step s in [0, npad-1).
__device__ __forceinline__ void rr_pair(int npad, int s, int k, int& p, int& q) {
    int a, b;
    if (k == 0) { a = npad - 1; b = s; }
    else { a = (s + k) % (npad - 1); b = (s - k + npad - 1) % (npad - 1); }
    p = a < b ? a : b; q = a < b ? b : a;
}

// In: Hermitian A (sm.Are/Aim).  Out: eigenvalues on the diagonal of A, eigenvectors in U.
// All threads of the block must call; uses __syncthreads().
__device__ inline void jacobi_hermitian_block(JacobiSmem& sm, int n, int max_sweeps = 24) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int npad = n + (n & 1), h = npad / 2;
    for (int i = tid; i < n * n; i += nt) { sm.Ure[i] = (i % n == i / n) ? 1.0 : 0.0; sm.Uim[i] = 0.0; }
    if (tid == 0) {
        double f = 0.0;
        for (int i = 0; i < n * n; ++i) f += sm.Are[i] * sm.Are[i] + sm.Aim[i] * sm.Aim[i];
        sm.red[0] = f; sm.red[2] = 0.0;
    }
    __syncthreads();
    if (n < 2 || sm.red[0] == 0.0) return;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
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
                    if (m2 > 0.0 && m2 > 1e-36 * fabs(app * aqq)) {
                        sm.offacc[tid] += m2;
                        double mag = sqrt(m2);
                        er = ar / mag; ei = ai / mag;
                        double tau = (aqq - app) / (2.0 * mag);
                        double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                sm.c[tid] = c; sm.s[tid] = s; sm.er[tid] = er; sm.ei[tid] = ei;
            }
            __syncthreads();
            // phase 1: every 2x2 block (k1,k2) <- J1^H * block * J2 ; U(:, {p,q}) <- U(:, {p,q}) * J
            for (int t = tid; t < h * h; t += nt) {
                int k1 = t / h, k2 = t % h;
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
                if (k1 == k2) {   // the annihilated pair: exact zeros off-diagonal, real diagonal
                    br[0][1] = bi[0][1] = br[1][0] = bi[1][0] = 0.0; bi[0][0] = 0.0; bi[1][1] = 0.0;
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
            sm.red[2] = (off <= 1e-31 * sm.red[0]) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (sm.red[2] != 0.0) break;
    }
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
