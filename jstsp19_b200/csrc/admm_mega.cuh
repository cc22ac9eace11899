// admm_mega.cuh - the whole proposed ADMM solve (proposed_algorithm.m:32-70, all Imax iterations) of one trial inside ONE
// persistent CTA: k_psi_mega.  Same algebra and operand images as admm_psi.cuh (structured dictionary B = (I (x) Dt') Psi_bar,
// bf16-exact Toeplitz pilots, three-term bf16 split of the small operands on tcgen05.mma.kind::f16), re-organised so that
// nothing but the streamed state leaves the SM:
//
//   grid = min(trials, SMs) CTAs x 480 threads, one CTA per SM (220 KiB shared memory, all 512 TMEM columns), each CTA takes
//   trials b = blockIdx.x, + gridDim.x, ... and runs the four phases of every iteration back to back:
//
//   F  for each 128-column chunk:  pass 1  Xs = Q_S e        (tensor core, accumulator in TMEM)                 (.m:58)
//                                  C, V2, Y = W Z, X, V1, K, XV += alpha G   (worker warps, state tiles via a TMA ring)
//                                  pass 2  T1'_l += (K - XV) e(. - l)^H  accumulated ACROSS the chunks in TMEM  (.m:47)
//                                  Gram of the next SVT input, accumulated across the chunks in registers      (.m:35)
//   R  Res_l = A'(T1'_l Dt), |Res|^2, operand image of G = (A Res) B      (workers, T1' read straight from TMEM)
//      ... while four dedicated warps run the fp64 Jacobi eigen-solve of the Gram matrix -> W of the next iteration
//   G  for each chunk:  G = Q_Res e (tensor core), |G|^2, G tile -> HBM
//   S  alpha = |Res|^2 / |G|^2, V += alpha Res, S = soft(V), operand image of the next pass 1                  (.m:48-56)
//
// Against the four-kernel form (k_fused_psi / k_psi_res / k_psi_g / k_psi_step + k_svt_weights on a side stream) this removes the
// T1' partials, the per-chunk Gram partials, 400 launches per solve and every kernel-boundary bubble; HBM sees only
// X, V1, V2, XV (read + write), subY, G and the pilot tiles.  Warp roles: 0-7 workers, 8 TMA loads of state and pilot tiles,
// 9 TMA bulk loads of the operand images (half-tap slots), 10 MMA issuer, 11-14 Jacobi.
//
// Preconditions (checked by the host, otherwise the four-kernel path runs): N = 16 rows, Nt = Gt = 64 with Dt the unitary DFT
// grid of wideband_mmwave_channel.m:9-10, L <= 4 (4 x 96 + 128 TMEM columns), M a multiple of 128, G <= 16.
#pragma once
#include "admm_psi.cuh"
#include "jacobi.cuh"

namespace jstsp {
namespace mega {

using namespace psi;
using tc::MC;
constexpr int MZP = MC + 4;                      // pitch of the planar Z tile: rows stay 16-byte aligned (LDS.128 in the Gram), consecutive rows 4 banks apart
constexpr int MZREG = 2 * N * MZP * 4;

constexpr int NWK = 8;                           // worker warps
constexpr int WTHREADS = 32 * NWK;
constexpr int W_LD = 8, W_Q = 9, W_MMA = 10, W_JAC = 11, NJW = 4, JT = 32 * NJW;
constexpr int MTHREADS = 32 * (W_JAC + NJW);     // 480
constexpr int NQ = 2;                            // half-tap operand slots in flight (pass 1)
constexpr int NIN = 4;                           // state-tile ring
constexpr int MEGA_MAXL = 4;
constexpr int KLBO = NRG * 128;                  // pass-2 small operand, MN-major: [8-column K group][row group][8 columns][8 rows] bf16, K groups KLBO apart
constexpr int LDJ = 17;                          // leading dimension of the 16 x 16 fp64 Jacobi matrices (bank spread)
constexpr int JN2 = LDJ * N;
constexpr int NSLOT = 2;                         // trials interleaved per CTA: the eigen-solve of one hides behind the phases of the other

constexpr int OFF_E = 0;                                   // two pilot tiles
constexpr int OFF_IN = OFF_E + 2 * TILE;                   // state-tile ring (SWIZZLE_128B, 1024-byte aligned)  } phase G: together the resident
constexpr int OFF_Q = OFF_IN + NIN * SLOT;                 // operand slots of pass 1                            } operand image of (A Res),
constexpr int OFF_Z = OFF_Q + NQ * QSLOT;                  // planar Z tile                                      } MEGA_MAXL taps x QTAP bytes
constexpr int OFF_KOP = OFF_Z + MZREG;                     // pass-2 small operand; phases R / S: the N x NT work matrices U, V
constexpr int KOPREG = (MC / 8) * KLBO;
constexpr int OFF_E2 = OFF_Z + (MEGA_MAXL * QTAP - NIN * SLOT - NQ * QSLOT);   // phase G only: third pilot tile behind the image tail, over the rest of Z, the
constexpr int E2PAD = OFF_E2 + TILE - (OFF_KOP + KOPREG);                     // pass-2 operand region and E2PAD bytes more (all idle between the phases R and S)
constexpr int OFF_W = OFF_KOP + KOPREG + E2PAD;            // W (fp32 complex 16 x 16) per trial slot
constexpr int OFF_A = OFF_W + NSLOT * WREG;                // A, A' per trial slot
constexpr int OFF_TW = OFF_A + NSLOT * 2 * N * N * 8;      // FFT twiddles
constexpr int OFF_JAC = OFF_TW + NT * 8;                   // Jacobi warps: A, U (fp64, padded), similarity scratch, rotation parameters
constexpr int JACREG = (4 * JN2 + 2 * N * N) * 8 + 512;
constexpr int OFF_MISC = OFF_JAC + JACREG;                 // small reductions
constexpr int OFF_BAR = OFF_MISC + 256;
constexpr size_t SMEM = (size_t)OFF_BAR + 512;
static_assert(OFF_IN % 1024 == 0 && OFF_Q % 1024 == 0, "tile alignment");
static_assert(2 * LDU * NT * 8 <= KOPREG, "work matrices do not fit the pass-2 operand region");
static_assert(MEGA_MAXL * QTAP <= NIN * SLOT + NQ * QSLOT + MZREG, "resident operand image of phase G does not fit");
static_assert(OFF_JAC % 16 == 0 && OFF_KOP % 16 == 0 && OFF_Z % 16 == 0 && OFF_E2 % 16 == 0 && E2PAD >= 0, "alignment");
static_assert(SMEM <= 232448, "shared memory");

__device__ __forceinline__ void wsync() { asm volatile("bar.sync 1, %0;" ::"n"(WTHREADS) : "memory"); }
__device__ __forceinline__ void jsync() { asm volatile("bar.sync 2, %0;" ::"n"(JT) : "memory"); }
struct WSync { __device__ __forceinline__ void operator()() const { wsync(); } };
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// this thread's 8 rows of one column (64 contiguous bytes): two 256-bit stores, i.e. whole 32-byte sectors per lane
__device__ __forceinline__ void stg8(cx<float>* __restrict__ g, const cx<float> (&v)[8], bool narrow = false) {
    if (narrow) {
#pragma unroll
        for (int u = 0; u < 4; ++u) reinterpret_cast<float4*>(g)[u] = make_float4(v[2 * u].re, v[2 * u].im, v[2 * u + 1].re, v[2 * u + 1].im);
        return;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(g + 4 * u), "f"(v[4 * u].re), "f"(v[4 * u].im), "f"(v[4 * u + 1].re), "f"(v[4 * u + 1].im),
                     "f"(v[4 * u + 2].re), "f"(v[4 * u + 2].im), "f"(v[4 * u + 3].re), "f"(v[4 * u + 3].im)
                     : "memory");
}
// warp-level release of a ring slot / accumulator: every lane has finished, lane 0 arrives (barrier count = NWK)
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(bar);
}

// Release of two state-ring slots this warp has read into registers.  The reads are generic-proxy accesses and the refill is an async-proxy
// (TMA) write: without the proxy fence the refill was observed to overtake loads still in flight (a tile read back as its successor).
__device__ __forceinline__ void ring_release(uint64_t* a, uint64_t* b, int lane) {
    tc::fence_async_smem();
    __syncwarp();
    if (lane == 0) { tc::mbar_arrive(a); tc::mbar_arrive(b); }
}

// ---- fp64 Hermitian Jacobi eigen-solve of the 16 x 16 Gram matrix on the four Jacobi warps (same rotations, pairing and stopping
// rules as jacobi_hermitian_block / jacobi_similarity_block of jacobi.cuh, specialised: named barrier, padded leading dimension) -----
struct Jac16 {
    double *Are, *Aim, *Ure, *Uim, *Tre, *Tim, *c, *s, *er, *ei, *offacc, *red;
    int *pp, *qq;
    __device__ void carve(unsigned char* work) {
        double* p = reinterpret_cast<double*>(work);
        Are = p; p += JN2; Aim = p; p += JN2; Ure = p; p += JN2; Uim = p; p += JN2; Tre = p; p += N * N; Tim = p; p += N * N;
        c = p; p += 8; s = p; p += 8; er = p; p += 8; ei = p; p += 8; offacc = p; p += 8; red = p; p += 8;
        pp = reinterpret_cast<int*>(p); qq = pp + 8;
    }
};
__device__ __forceinline__ double jac_sum(Jac16& sm, double v, int jt) {       // deterministic sum over the 128 Jacobi threads
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (jt % 32 == 0) sm.red[4 + jt / 32] = v;
    jsync();
    const double r = (sm.red[4] + sm.red[5]) + (sm.red[6] + sm.red[7]);
    jsync();
    return r;
}
// A <- Q^H A Q with Q = U (warm start)
__device__ inline void jac16_similarity(Jac16& sm, int jt) {
    for (int t = jt; t < N * N; t += JT) {               // T = A Q
        const int i = t % N, j = t / N;
        double re = 0.0, im = 0.0;
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
            const double ar = sm.Are[i + LDJ * k], ai = sm.Aim[i + LDJ * k], qr = sm.Ure[k + LDJ * j], qi = sm.Uim[k + LDJ * j];
            re += ar * qr - ai * qi; im += ar * qi + ai * qr;
        }
        sm.Tre[t] = re; sm.Tim[t] = im;
    }
    jsync();
    for (int t = jt; t < N * N; t += JT) {               // A = Q^H T
        const int i = t % N, j = t / N;
        double re = 0.0, im = 0.0;
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
            const double qr = sm.Ure[k + LDJ * i], qi = -sm.Uim[k + LDJ * i], tr = sm.Tre[k + N * j], ti = sm.Tim[k + N * j];
            re += qr * tr - qi * ti; im += qr * ti + qi * tr;
        }
        sm.Are[i + LDJ * j] = re; sm.Aim[i + LDJ * j] = im;
    }
    jsync();
    for (int t = jt; t < N * N; t += JT) {               // exact Hermitian symmetry
        const int i = t % N, j = t / N;
        if (i < j) {
            const double re = 0.5 * (sm.Are[i + LDJ * j] + sm.Are[j + LDJ * i]), im = 0.5 * (sm.Aim[i + LDJ * j] - sm.Aim[j + LDJ * i]);
            sm.Are[i + LDJ * j] = re; sm.Aim[i + LDJ * j] = im; sm.Are[j + LDJ * i] = re; sm.Aim[j + LDJ * i] = -im;
        } else if (i == j) sm.Aim[i + LDJ * j] = 0.0;
    }
    jsync();
}
// returns the Frobenius norm^2 of the input (0: nothing to do); eigenvalues end on the diagonal of A, eigenvectors in U
__device__ inline double jac16_solve(Jac16& sm, int jt, bool keep_U, double stop_rel2, int& sweeps) {
    constexpr int h = N / 2;
    if (!keep_U) for (int i = jt; i < N * N; i += JT) { sm.Ure[(i % N) + LDJ * (i / N)] = (i % N == i / N) ? 1.0 : 0.0; sm.Uim[(i % N) + LDJ * (i / N)] = 0.0; }
    double f = 0.0;
    for (int i = jt; i < N * N; i += JT) { const int a = (i % N) + LDJ * (i / N); f += sm.Are[a] * sm.Are[a] + sm.Aim[a] * sm.Aim[a]; }
    const double fro2 = jac_sum(sm, f, jt);
    if (fro2 == 0.0) return 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        if (keep_U || sweep > 0) {
            double o2 = 0.0;
            for (int i = jt; i < N * N; i += JT) if (i % N != i / N) { const int a = (i % N) + LDJ * (i / N); o2 += sm.Are[a] * sm.Are[a] + sm.Aim[a] * sm.Aim[a]; }
            o2 = jac_sum(sm, o2, jt);
            if (o2 <= stop_rel2 * sqrt(stop_rel2) * fro2) break;
        }
        if (jt < h) sm.offacc[jt] = 0.0;
        ++sweeps;
        for (int step = 0; step < N - 1; ++step) {
            if (jt < h) {                                // rotation parameters, one thread per pair (fp32 tangent, exactly unitary fp64 rotation)
                int p, q; rr_pair(N, step, jt, p, q);
                sm.pp[jt] = p; sm.qq[jt] = q;
                double c = 1.0, s = 0.0, er = 1.0, ei = 0.0;
                const double ar = sm.Are[p + LDJ * q], ai = sm.Aim[p + LDJ * q], app = sm.Are[p + LDJ * p], aqq = sm.Are[q + LDJ * q];
                const double m2 = ar * ar + ai * ai;
                if (m2 > 1e-290 && m2 > 1e-36 * fabs(app * aqq)) {
                    sm.offacc[jt] += m2;
                    const double rm = fast_rsqrt(m2);
                    er = ar * rm; ei = ai * rm;
                    const float tf = (float)((aqq - app) * 0.5 * rm);
                    const double t = (double)(copysignf(1.f, tf) / (fabsf(tf) + sqrtf(fmaf(tf, tf, 1.f))));
                    c = fast_rsqrt(1.0 + t * t);
                    s = t * c;
                }
                sm.c[jt] = c; sm.s[jt] = s; sm.er[jt] = er; sm.ei[jt] = ei;
            }
            jsync();
            if (jt < h * h) {                            // 2 x 2 block (k1, k2) <- J1^H block J2
                const int k1 = jt % h, k2 = jt / h;
                const int p1 = sm.pp[k1], q1 = sm.qq[k1], p2 = sm.pp[k2], q2 = sm.qq[k2];
                const double c1 = sm.c[k1], s1 = sm.s[k1], e1r = sm.er[k1], e1i = sm.ei[k1];
                const double c2 = sm.c[k2], s2 = sm.s[k2], e2r = sm.er[k2], e2i = sm.ei[k2];
                if (s1 != 0.0 || s2 != 0.0) {
                    double br[2][2], bi[2][2];
                    br[0][0] = sm.Are[p1 + LDJ * p2]; bi[0][0] = sm.Aim[p1 + LDJ * p2];
                    br[0][1] = sm.Are[p1 + LDJ * q2]; bi[0][1] = sm.Aim[p1 + LDJ * q2];
                    br[1][0] = sm.Are[q1 + LDJ * p2]; bi[1][0] = sm.Aim[q1 + LDJ * p2];
                    br[1][1] = sm.Are[q1 + LDJ * q2]; bi[1][1] = sm.Aim[q1 + LDJ * q2];
                    if (s1 != 0.0) {
#pragma unroll
                        for (int cc = 0; cc < 2; ++cc) {
                            const double xr = br[0][cc], xi = bi[0][cc], yr = br[1][cc], yi = bi[1][cc];
                            const double eyr = e1r * yr - e1i * yi, eyi = e1r * yi + e1i * yr;
                            br[0][cc] = c1 * xr - s1 * eyr; bi[0][cc] = c1 * xi - s1 * eyi;
                            br[1][cc] = s1 * xr + c1 * eyr; bi[1][cc] = s1 * xi + c1 * eyi;
                        }
                    }
                    if (s2 != 0.0) {
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const double xr = br[rr][0], xi = bi[rr][0], yr = br[rr][1], yi = bi[rr][1];
                            const double eyr = e2r * yr + e2i * yi, eyi = e2r * yi - e2i * yr;
                            br[rr][0] = c2 * xr - s2 * eyr; bi[rr][0] = c2 * xi - s2 * eyi;
                            br[rr][1] = s2 * xr + c2 * eyr; bi[rr][1] = s2 * xi + c2 * eyi;
                        }
                    }
                    if (k1 == k2) { br[1][0] = br[0][1]; bi[1][0] = -bi[0][1]; bi[0][0] = 0.0; bi[1][1] = 0.0; }
                    sm.Are[p1 + LDJ * p2] = br[0][0]; sm.Aim[p1 + LDJ * p2] = bi[0][0];
                    sm.Are[p1 + LDJ * q2] = br[0][1]; sm.Aim[p1 + LDJ * q2] = bi[0][1];
                    sm.Are[q1 + LDJ * p2] = br[1][0]; sm.Aim[q1 + LDJ * p2] = bi[1][0];
                    sm.Are[q1 + LDJ * q2] = br[1][1]; sm.Aim[q1 + LDJ * q2] = bi[1][1];
                }
            } else {                                     // U(:, {p,q}) <- U(:, {p,q}) J : 8 pairs x 16 rows over 64 threads
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int t = (jt - h * h) + rep * (JT - h * h), k = t / N, i = t % N;
                    const double s2 = sm.s[k];
                    if (s2 == 0.0) continue;
                    const int p = sm.pp[k], q = sm.qq[k];
                    const double c2 = sm.c[k], e2r = sm.er[k], e2i = sm.ei[k];
                    const double xr = sm.Ure[i + LDJ * p], xi = sm.Uim[i + LDJ * p], yr = sm.Ure[i + LDJ * q], yi = sm.Uim[i + LDJ * q];
                    const double eyr = e2r * yr + e2i * yi, eyi = e2r * yi - e2i * yr;
                    sm.Ure[i + LDJ * p] = c2 * xr - s2 * eyr; sm.Uim[i + LDJ * p] = c2 * xi - s2 * eyi;
                    sm.Ure[i + LDJ * q] = s2 * xr + c2 * eyr; sm.Uim[i + LDJ * q] = s2 * xi + c2 * eyi;
                }
            }
            jsync();
        }
        if (jt == 0) {
            double off = 0.0;
            for (int k = 0; k < h; ++k) off += sm.offacc[k];
            sm.red[2] = (off <= stop_rel2 * fro2) ? 1.0 : 0.0;
        }
        jsync();
        if (sm.red[2] != 0.0) break;
    }
    return fro2;
}
// W = U diag(max(0, 1 - tau / sigma)) U^H (svt.m:7), all-zero input -> zeros (svt.m:7-13)
__device__ inline void jac16_weights(Jac16& sm, int jt, double tau, double fro2, cx<float>* __restrict__ W) {
    if (jt < N) {
        const double lam = sm.Are[jt + LDJ * jt];
        const double sig = lam > 0.0 ? sqrt(lam) : 0.0;
        sm.Tre[jt] = (fro2 != 0.0 && sig > tau) ? (1.0 - tau / sig) : 0.0;
    }
    jsync();
    for (int t = jt; t < N * N; t += JT) {
        const int i = t % N, j = t / N;
        double wr = 0.0, wi = 0.0;
#pragma unroll 4
        for (int k = 0; k < N; ++k) {
            const double f = sm.Tre[k];
            const double ar = sm.Ure[i + LDJ * k], ai = sm.Uim[i + LDJ * k], br = sm.Ure[j + LDJ * k], bi = -sm.Uim[j + LDJ * k];
            wr += f * (ar * br - ai * bi); wi += f * (ar * bi + ai * br);
        }
        reinterpret_cast<float*>(W)[i + N * j] = (float)wr; reinterpret_cast<float*>(W)[N * N + i + N * j] = (float)wi;      // planar: re plane | im plane
    }
}

// Operand-image rows of row n and four antennas k0 .. k0 + 3 (kc = 2 k0 .. 2 k0 + 7: one 16-byte K group), all three split terms:
// rows (u, n, re) = [Qr, -Qi] and (u, n, im) = [Qi, Qr] are 32 contiguous bytes per split term.  store(byte offset in the tap image, re row, im row).
template <class Store>
__device__ __forceinline__ void put_q4(int n, int k0, const cx<float> (&q)[4], float scale, int rot, Store store) {
    unsigned short sr[4][3], si[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) { split3(scale * q[j].re, sr[j]); split3(scale * q[j].im, si[j]); }
    const int kc = 2 * k0;
    const uint32_t base = (uint32_t)(kc / 16) * QKS + (uint32_t)((kc % 16) / 8) * (NRG * 128);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int r = 32 * ((u + rot) % 3) + 2 * n;
        uint32_t a[4], bq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a[j] = (uint32_t)sr[j][u] | ((uint32_t)(si[j][u] ^ 0x8000u) << 16);      // (Qr, -Qi): the split terms of -x are those of x with the sign flipped
            bq[j] = (uint32_t)si[j][u] | ((uint32_t)sr[j][u] << 16);                 // (Qi,  Qr)
        }
        store(base + (uint32_t)(r / 8) * 128 + (uint32_t)(r % 8) * 16, make_uint4(a[0], a[1], a[2], a[3]), make_uint4(bq[0], bq[1], bq[2], bq[3]));
    }
}
__device__ __forceinline__ void stg_u4x2(unsigned char* g, uint4 a, uint4 b) {      // 32 contiguous, 32-byte aligned bytes
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(g), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// The MMA warp runs its loops with all 32 lanes (warp-uniform control flow and operands, so the descriptors live in uniform registers and
// no per-lane broadcast loop is generated); one elected lane issues.
__device__ __forceinline__ bool elect_one() {           // true in exactly one lane of the (converged) warp
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) { umma_bf16(d_tmem, a, b, idesc, accumulate); }
__device__ __forceinline__ void commit_elect(uint64_t* bar) { tc::umma_commit(bar); }
// accumulator of pass 1 / of G: 128 TMEM columns; even taps -> [0, 96) with operand rows [hi | mid | lo], odd taps -> [32, 128) with rows
// [mid | lo | hi].  The first MMA of tap 0 initialises [0, 96); the first k step of tap 1 is split so that [96, 128) is initialised too.
// RESIDENT: the whole image ([tap][QTAP]) sits in shared memory at slots_a (phase G); otherwise it streams through the NQ half-tap slots.
// Descriptors: the address field counts 16-byte units, so every operand address is the base descriptor plus a small constant.
template <bool RESIDENT>
__device__ __forceinline__ void issue_taps(uint32_t acc, uint32_t tile_a, uint32_t slots_a, uint64_t* q_full, uint64_t* q_empty, uint32_t& q_n, int L) {
    constexpr uint32_t id96 = instr_desc_bf16(128, NS, 0), id64 = instr_desc_bf16(128, 64, 0), id32 = instr_desc_bf16(128, 32, 0);
    const uint64_t a_base = tc::smem_desc(tile_a, RS, 128, 0), b_base = tc::smem_desc(slots_a, NRG * 128, 128, 0);
    for (int i = 0; i < 2 * L; ++i) {
        const int slot = q_n % NQ, l = i >> 1, hf = i & 1;
        if (!RESIDENT) { mbar_wait(&q_full[slot], (q_n / NQ) & 1); tc::tc_fence_after(); }
        const uint64_t ad0 = a_base + (uint64_t)(L - 1 - l) + (uint64_t)(hf * (KC / 32) * 2 * (RS / 16));
        const uint64_t bd0 = b_base + (uint64_t)((RESIDENT ? i : slot) * (QSLOT / 16));
        const uint32_t d = acc + 32 * (l & 1);
#pragma unroll
        for (int j = 0; j < KC / 32; ++j) {
            const uint64_t ad = ad0 + (uint64_t)(j * 2 * (RS / 16)), bd = bd0 + (uint64_t)(j * (QKS / 16));
            if (j == 0 && i == 2) {                                   // tap 1, first k step
                umma_elect(acc + 32, ad, bd, id64, 1u);
                umma_elect(acc + 96, ad, bd + (uint64_t)(8 * 128 / 16), id32, 0u);
            } else {
                umma_elect(d, ad, bd, id96, (i | j) ? 1u : 0u);
            }
        }
        if (!RESIDENT) { commit_elect(&q_empty[slot]); ++q_n; }
    }
}
// this thread's 8 rows of column m: lo, mid, then the hi blocks (smallest terms first)
__device__ __forceinline__ void read_acc(uint32_t acc, uint32_t lane_base, int n0, int L, float (&xr)[8], float (&xi)[8]) {
    float a[16];
#pragma unroll
    for (int r = 0; r < 8; ++r) { xr[r] = 0.f; xi[r] = 0.f; }
#pragma unroll
    for (int blk = 0; blk < 4; ++blk) {
        const int col = blk == 0 ? 64 : blk == 1 ? 32 : blk == 2 ? 0 : 96;
        if (blk == 3 && L < 2) break;
        tc::tmem_ld16(acc + lane_base + col + 2 * n0, a);
#pragma unroll
        for (int r = 0; r < 8; ++r) { xr[r] += a[2 * r]; xi[r] += a[2 * r + 1]; }
    }
}

// developer hook (tools/mega_probe.py): clock64 stamps of the first trial of every CTA, 16 iterations x 8 slots per CTA
// one copy of the 16 x 16 row-mixing product for its three call sites (code size: the kernel is instruction-fetch sensitive)
__device__ __noinline__ void apply_a_shared(const cx<float>* __restrict__ Mx, const cx<float>* __restrict__ in, cx<float>* __restrict__ out) { psi::apply_a(Mx, in, out, NT); }
// In-kernel instrumentation (phase stamps, hang beacons, data dumps) is compiled in only with -DJSTSP_MEGA_DEBUG=1 (`make debug-lib`, loaded through JSTSP_LIB
// by tools/mega_probe.py / mega_beacon.py / mega_dump.py): the product kernel carries none of it (ncu: 14 % of all warp stalls were instruction-fetch stalls).
#ifndef JSTSP_MEGA_DEBUG
#define JSTSP_MEGA_DEBUG 0
#endif
#if JSTSP_MEGA_DEBUG
#define MEGA_STAMP(slot, who)                                                                                                      \
    do {                                                                                                                           \
        if (p.dbg && p.dbg_kernel == 7 && (who) && b == (int)blockIdx.x && it < 16) p.dbg[((size_t)blockIdx.x * 16 + it) * 8 + (slot)] = clock64(); \
    } while (0)
// progress beacon for hang triage (tools/mega_beacon.py): the debug buffer is mapped host memory, CTA 0 reports where each role is
#define MEGA_BEACON(slot, value)                                                                                                   \
    do {                                                                                                                           \
        if (p.dbg && p.dbg_kernel == 10 && blockIdx.x == 0) { *reinterpret_cast<volatile long long*>(p.dbg + (slot)) = (long long)(value); __threadfence_system(); } \
    } while (0)
#define MEGA_STAMP3(kid, slot)                                                                                                     \
    do {                                                                                                                           \
        if (p.dbg && p.dbg_kernel == (kid) && tid == 0 && l == 1 && b == (int)blockIdx.x && it < 16) p.dbg[((size_t)blockIdx.x * 16 + it) * 8 + (slot)] = clock64(); \
    } while (0)
// data dump for race triage (tools/mega_dump.py): per trial 1024 floats - checksums of what iteration 1 consumed
#define MEGA_DUMP_ON (p.dbg && p.dbg_kernel == 14 && it == 1)
#define MEGA_DUMP_ADD(idx, val) do { if (MEGA_DUMP_ON) atomicAdd(reinterpret_cast<float*>(p.dbg) + (size_t)b * 1024 + (idx), (val)); } while (0)
#define MEGA_STAMP2(slot)                                                                                                          \
    do {                                                                                                                           \
        if (p.dbg && p.dbg_kernel == 8 && tid == 0 && c == 3 && b == (int)blockIdx.x && it < 16) p.dbg[((size_t)blockIdx.x * 16 + it) * 8 + (slot)] = clock64(); \
    } while (0)
// cycles thread 0 waits for a pair of state tiles: site 0 X,V1 / 1 V2,subY / 2 XV,G, per chunk, iteration 5 of the CTA's first trial
#define MEGA_WAIT_BEGIN long long w0_ = (p.dbg && p.dbg_kernel == 15) ? clock64() : 0
#define MEGA_WAIT_END(site)                                                                                                        \
    do {                                                                                                                           \
        if (p.dbg && p.dbg_kernel == 15 && tid == 0 && b == (int)blockIdx.x && it == 5 && c < 8) p.dbg[(size_t)blockIdx.x * 128 + c * 3 + (site)] = clock64() - w0_; \
    } while (0)
#else
#define MEGA_STAMP(slot, who) do { } while (0)
#define MEGA_BEACON(slot, value) do { } while (0)
#define MEGA_STAMP3(kid, slot) do { } while (0)
#define MEGA_DUMP_ON false
#define MEGA_DUMP_ADD(idx, val) do { } while (0)
#define MEGA_STAMP2(slot) do { } while (0)
#define MEGA_WAIT_BEGIN do { } while (0)
#define MEGA_WAIT_END(site) do { } while (0)
#endif
__global__ void __launch_bounds__(MTHREADS, 1) k_psi_mega(AdmmP<float> p, const __grid_constant__ Maps maps, In in, int nb) {
    constexpr int NH = N / 2;
    // launched before the host has seen the structure check's flags (run_admm, speculative pass): a failed check leaves everything untouched
    if (in.bad[0] != 0 || in.bad[1] != 0) return;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* etile = smem + OFF_E;
    unsigned char* inr = smem + OFF_IN;
    unsigned char* qsl = smem + OFF_Q;
    unsigned char* kop = smem + OFF_KOP;
    float* Zre = reinterpret_cast<float*>(smem + OFF_Z);
    float* Zim = Zre + N * MZP;
    cx<float>* Wsm = reinterpret_cast<cx<float>*>(smem + OFF_W);      // [slot][N * N]
    cx<float>* Asm = reinterpret_cast<cx<float>*>(smem + OFF_A);      // [slot][A | A'][N * N]
    cx<float>* tw = reinterpret_cast<cx<float>*>(smem + OFF_TW);
    double* misc = reinterpret_cast<double*>(smem + OFF_MISC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t *e_full = bars + 37, *e_empty = bars + 40, *r_done = bars + 43, *in_full = bars + 4, *in_empty = bars + 8, *q_full = bars + 12, *q_empty = bars + 16;
    uint64_t *acc_full = bars + 20, *acc_empty = bars + 22, *kop_full = bars + 24, *kop_empty = bars + 25, *t1_full = bars + 26;
    uint64_t *g_done = bars + 27, *img_ready = bars + 28, *g_done_q = bars + 29, *q_ready = bars + 30, *gram_ready = bars + 32, *w_ready = bars + 34;   // the last three: one per trial slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 44);

    const int tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    const int M = p.M, L = in.L, nch = M / MC, imax = p.imax, G = p.G;
    const size_t NM = (size_t)N * M;
    // this CTA's trials blockIdx.x, + gridDim.x, ... are taken two at a time; the work items of a pair are (iteration, slot) in that order
    const int ntr = (int)blockIdx.x < nb ? (nb - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
#define MEGA_ITEMS_BEGIN                                                                  \
    for (int pr = 0; pr < ntr; pr += NSLOT) {                                             \
        const int nw = (ntr - pr) < NSLOT ? (ntr - pr) : NSLOT;                           \
        for (int it = 0; it < imax; ++it)                                                 \
            for (int w = 0; w < nw; ++w) {                                                \
                const int b = (int)blockIdx.x + (pr + w) * (int)gridDim.x;
#define MEGA_ITEMS_END }}

    if (tid == 0) {
        for (int s = 0; s < 3; ++s) { mbar_init(&e_full[s], 1); mbar_init(&e_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], NWK); }
        mbar_init(r_done, NWK);
        for (int s = 0; s < NIN; ++s) { mbar_init(&in_full[s], 1); mbar_init(&in_empty[s], NWK); }
        for (int s = 0; s < NQ; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
        mbar_init(kop_full, NWK); mbar_init(kop_empty, 1); mbar_init(t1_full, 1);
        mbar_init(g_done, NWK); mbar_init(img_ready, NWK); mbar_init(g_done_q, NWK);
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&q_ready[s], NWK); mbar_init(&gram_ready[s], NWK); mbar_init(&w_ready[s], 1); }
        mbar_fence_init();
    }
    if (warp == W_LD) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = *tmem_slot;
    const uint32_t ACC0 = tm + 384, ACC1 = tm;               // pass 1 uses ACC0; G alternates ACC0 / ACC1 (T1' has been consumed by then)

    if (warp == W_LD) {
        // ===== TMA producer: pilot tiles and state tiles =====
        if (elect_one()) {
            uint32_t euse[3] = {0, 0, 0}, in_n = 0, gd_n = 0, rd_n = 0;
            // pilot tile of chunk c into buffer `buf` (phase F alternates 0 / 1, phase G cycles 0 / 1 / 2); a buffer is released by the last MMA that read it
            auto load_e = [&](int buf, int c, int b) {
                if (euse[buf] > 0) mbar_wait(&e_empty[buf], (euse[buf] - 1) & 1);
                ++euse[buf];
                unsigned char* dst = buf < 2 ? etile + buf * TILE : smem + OFF_E2;
                mbar_expect_tx(&e_full[buf], TILE);
                // [kc group][column][8 kc]: the ROWS columns of one group are RS contiguous bytes in the image - 16 bulk copies, no 16-byte tensor rows
                const unsigned short* src = in.E + ((size_t)(in.ld_Psi ? b : 0) * NKG * (size_t)(M + 8) + (size_t)c * MC) * 8;
#pragma unroll 4
                for (int kg = 0; kg < NKG; ++kg) tma_bulk_g2s(dst + kg * RS, src + (size_t)kg * (M + 8) * 8, RS, &e_full[buf]);
            };
            auto load_in = [&](const CUtensorMap* map, int c, int b) {
                const int s = in_n % NIN, use = in_n / NIN;
                if (use > 0) mbar_wait(&in_empty[s], (use - 1) & 1);
                mbar_expect_tx(&in_full[s], SLOT);
                tc::tma_3d(inr + s * SLOT, map, 0, c * MC, b, &in_full[s]);
                ++in_n;
            };
            MEGA_ITEMS_BEGIN
                const int sy_b = p.ld_subY ? b : 0;
                // Phase G of the previous work item is over: its operand image has left the ring, and the state stored two items ago (same trial) is
                // ordered before the TMA loads below - the proxy fence must sit HERE, in the thread that issues the async-proxy reads after acquiring
                // the barrier (a fence only on the writers' side was measured to let stale XV tiles through).
                if (pr > 0 || it > 0 || w > 0) { mbar_wait(g_done, gd_n & 1); ++gd_n; fence_proxy_async_all(); }
                for (int c = 0; c < nch; ++c) {
                    MEGA_BEACON(0, it * 1000 + 100 + c);
                    load_e(c & 1, c, b);
                    load_in(&maps.X, c, b); load_in(&maps.V1, c, b); load_in(&maps.V2, c, b); load_in(&maps.SY, c, sy_b);
                    load_in(&maps.XV, c, b); load_in(&maps.G, c, b);
                }
                for (int c = 0; c < nch; ++c) {
                    MEGA_BEACON(0, it * 1000 + 300 + c);
                    if (c == 2) { mbar_wait(r_done, rd_n & 1); ++rd_n; }     // the third buffer lies over the work matrices of phase R
                    load_e(c % 3, c, b);
                }
                if (nch <= 2) { mbar_wait(r_done, rd_n & 1); ++rd_n; }
                MEGA_BEACON(0, it * 1000 + 999);
            MEGA_ITEMS_END
        }
        __syncwarp();
    } else if (warp == W_Q) {
        // ===== TMA producer: operand images, half a tap per slot, once per chunk =====
        if (elect_one()) {
            uint32_t q_n = 0, qr_n[NSLOT] = {0, 0}, gq_n = 0;
            auto stream = [&](const unsigned char* img) {
                for (int c = 0; c < nch; ++c)
                    for (int i = 0; i < 2 * L; ++i) {
                        const int s = q_n % NQ, use = q_n / NQ;
                        if (use > 0) mbar_wait(&q_empty[s], (use - 1) & 1);
                        mbar_expect_tx(&q_full[s], QSLOT);
                        tma_bulk_g2s(qsl + s * QSLOT, img + (size_t)i * QSLOT, QSLOT, &q_full[s]);
                        ++q_n;
                    }
            };
            MEGA_ITEMS_BEGIN
                // the operand slots are part of the previous item's resident phase-G image: wait until that phase is over (the image this item
                // needs was finished one item earlier still, by the same trial's phase S)
                if (pr > 0 || it > 0 || w > 0) { mbar_wait(g_done_q, gq_n & 1); ++gq_n; }
                if (it > 0) { mbar_wait(&q_ready[w], qr_n[w] & 1); ++qr_n[w]; fence_proxy_async_all(); stream(in.QopS + (size_t)b * L * QTAP); }
            MEGA_ITEMS_END
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        // ===== MMA issuer: one elected lane (elect.sync, so that the compiler knows the branch holds a single thread) =====
        if (elect_one()) {
            constexpr uint32_t id2 = instr_desc_bf16(128, NS, 1, 1);
            const uint32_t e_a = smem_u32(etile), q_a = smem_u32(qsl), k_a = smem_u32(kop), img_a = smem_u32(inr);
            uint32_t euse[3] = {0, 0, 0}, q_n = 0, accn[2] = {0, 0}, kop_n = 0, ig_n = 0;
            const uint32_t e_addr[3] = {e_a, e_a + TILE, smem_u32(smem + OFF_E2)};
            const uint64_t k_base = tc::smem_desc(k_a, KLBO, 128, 0);
            MEGA_ITEMS_BEGIN
                {
                    (void)b;
                    // ---- phase F: P1(0), then P1(c), P2(c - 1) interleaved ----
                    for (int c = 0; c <= nch; ++c) {
                        if (c < nch && it > 0) {
                            const int eb = c & 1;
                            mbar_wait(&e_full[eb], euse[eb] & 1);
                            if (accn[0] > 0) mbar_wait(&acc_empty[0], (accn[0] - 1) & 1);
                            tc::tc_fence_after();
                            issue_taps<false>(ACC0, e_addr[eb], q_a, q_full, q_empty, q_n, L);
                            commit_elect(&acc_full[0]);
                            ++accn[0];
                        }
                        if (c > 0) {
                            const int eb = (c - 1) & 1;
                            mbar_wait(&e_full[eb], euse[eb] & 1);
                            mbar_wait(kop_full, kop_n & 1); ++kop_n;
                            tc::tc_fence_after();
                            // pass 2: A = pilot tile, MN-major (LBO = 8-column K group stride, SBO = kc group stride); B = K operand, MN-major
                            const uint64_t a_base = tc::smem_desc(e_addr[eb], 128, RS, 0);
                            for (int l = 0; l < L; ++l) {
                                const uint64_t ad0 = a_base + (uint64_t)(L - 1 - l);
                                const uint32_t d = tm + NS * l;
#pragma unroll
                                for (int ks = 0; ks < MC / 16; ++ks)
                                    umma_elect(d, ad0 + (uint64_t)(ks * (256 / 16)), k_base + (uint64_t)(ks * (2 * KLBO / 16)), id2, (c > 1 || ks) ? 1u : 0u);
                            }
                            commit_elect(kop_empty);
                            commit_elect(&e_empty[eb]);
                            ++euse[eb];
                        }
                    }
                    commit_elect(t1_full);
                    // ---- phase G: the operand image of (A Res) is resident (built by the workers over the idle state ring) ----
                    const bool mdbg = p.dbg && p.dbg_kernel == 13 && lane == 0 && b == (int)blockIdx.x && it < 16;
                    long long tw_img = 0, tw_e = 0, tw_acc = 0, t_iss = 0, tq = mdbg ? clock64() : 0;
                    mbar_wait(img_ready, ig_n & 1); ++ig_n;
                    if (mdbg) { tw_img = clock64() - tq; }
                    const long long tg0 = mdbg ? clock64() : 0;
                    for (int c = 0; c < nch; ++c) {
                        const int buf = c & 1, eb = c % 3;
                        if (mdbg) tq = clock64();
                        mbar_wait(&e_full[eb], euse[eb] & 1);
                        if (mdbg) { tw_e += clock64() - tq; tq = clock64(); }
                        if (accn[buf] > 0) mbar_wait(&acc_empty[buf], (accn[buf] - 1) & 1);
                        if (mdbg) { tw_acc += clock64() - tq; tq = clock64(); }
                        tc::tc_fence_after();
                        issue_taps<true>(buf ? ACC1 : ACC0, e_addr[eb], img_a, q_full, q_empty, q_n, L);
                        commit_elect(&acc_full[buf]);
                        commit_elect(&e_empty[eb]);
                        if (mdbg) t_iss += clock64() - tq;
                        ++accn[buf]; ++euse[eb];
                    }
                    if (mdbg) { long long* d = p.dbg + ((size_t)blockIdx.x * 16 + it) * 8; d[0] = tw_img; d[1] = tw_e; d[2] = tw_acc; d[3] = t_iss; d[4] = clock64() - tg0; }
                }
            MEGA_ITEMS_END
        }
        __syncwarp();
    } else if (warp >= W_JAC) {
        // ===== Jacobi warps: W of iteration it + 1 from the Gram matrix the workers finish in phase F of iteration it =====
        const int jt = tid - 32 * W_JAC;
        Jac16 js; js.carve(smem + OFF_JAC);
        uint32_t gr_n[NSLOT] = {0, 0};
        MEGA_ITEMS_BEGIN
            if (it + 1 < imax) {
                const double tau = p.tauY[b] / p.rho[b];
                mbar_wait(&gram_ready[w], gr_n[w] & 1); ++gr_n[w];
                MEGA_STAMP(6, jt == 0);
                // warm start from the previous eigenvectors of this trial; every 16th iteration (and the first of a trial) restarts cold
                const bool warm = it > 0 && ((it + 1) % 16) != 0 && !(JSTSP_MEGA_DEBUG && (in.t1_red & 4));
                const bool jdbg = p.dbg && p.dbg_kernel == 9 && jt == 0 && b == (int)blockIdx.x && it < 16;
                long long* jd = p.dbg + ((size_t)blockIdx.x * 16 + it) * 8;
                if (jdbg) jd[0] = clock64();
                // the Gram matrix and the eigenvectors travel through global memory (L2): shared memory holds one solve, whichever trial it belongs to
                const double2* gin = reinterpret_cast<const double2*>(p.gram + (size_t)b * p.nmc * 2 * N * N);
                double* up = p.Uprev + (size_t)b * 2 * N * N;
                for (int t = jt; t < N * N; t += JT) {
                    const double2 g = __ldcg(gin + t);
                    js.Are[(t % N) + LDJ * (t / N)] = g.x; js.Aim[(t % N) + LDJ * (t / N)] = g.y;
                    if (warm) { js.Ure[(t % N) + LDJ * (t / N)] = __ldcg(up + t); js.Uim[(t % N) + LDJ * (t / N)] = __ldcg(up + N * N + t); }
                }
                jsync();
                if (warm) jac16_similarity(js, jt);
                if (jdbg) jd[1] = clock64();
                int sweeps = 0;
                const double fro2 = jac16_solve(js, jt, warm, 1e-10, sweeps);
                if (jdbg) { jd[2] = clock64(); jd[4] = sweeps; }
                jac16_weights(js, jt, tau, fro2, Wsm + w * N * N);
                for (int t = jt; t < N * N; t += JT) { up[t] = js.Ure[(t % N) + LDJ * (t / N)]; up[N * N + t] = js.Uim[(t % N) + LDJ * (t / N)]; }
                jsync();
                if (jdbg) jd[3] = clock64();
                MEGA_STAMP(7, jt == 0);
                if (jt == 0) tc::mbar_arrive(&w_ready[w]);
            }
        MEGA_ITEMS_END
    } else {
        // ===== workers =====
        const int quad = warp % 4, half = warp / 4;
        const int m = quad * 32 + lane;                      // column of the chunk (pass 1, G) / kc row of the tile (pass 2)
        const int n0 = half * NH;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        cx<float>* U = reinterpret_cast<cx<float>*>(kop);     // phases R / S: N x NT work matrices (leading dimension LDU) over the idle pass-2 operand
        cx<float>* V = U + LDU * NT;
        uint32_t in_n = 0, accn[2] = {0, 0}, kop_n = 0, t1_n = 0, w_n[NSLOT] = {0, 0};
        const bool narrow = JSTSP_MEGA_DEBUG && (in.t1_red & 2) != 0;      // developer switch (instrumented build only): 128-bit stores
        if (tid < NT) { float sn, cs; sincospif(-2.0f * (float)tid / NT, &sn, &cs); tw[tid] = mk<float>(cs, sn); }
        auto in_slot = [&](uint32_t i) { return inr + (i % NIN) * SLOT; };
        auto in_wait = [&](uint32_t i) { mbar_wait(&in_full[i % NIN], (i / NIN) & 1); };
        float ap0 = 0.f, ap1 = 0.f;                           // alpha of the previous iteration, per trial slot

        for (int pr = 0; pr < ntr; pr += NSLOT) {
            const int nw = (ntr - pr) < NSLOT ? (ntr - pr) : NSLOT;
            wsync();                                          // the previous pair is done with A / W
            for (int w = 0; w < nw; ++w) {
                const int b = (int)blockIdx.x + (pr + w) * (int)gridDim.x;
                const cx<float>* A = p.A + (long long)b * p.ld_A;
                const cx<float> a = tid < N * G ? A[tid] : mk<float>(0.f, 0.f);
                cx<float>* As = Asm + w * 2 * N * N;
                As[tid] = a; As[N * N + (tid / N) + N * (tid % N)] = mk<float>(a.re, -a.im);
                Wsm[w * N * N + tid] = mk<float>(0.f, 0.f);   // svt of the all-zero first iterate is zero (svt.m:7-13)
            }
            ap0 = 0.f; ap1 = 0.f;
            wsync();
          for (int it = 0; it < imax; ++it)
            for (int w = 0; w < nw; ++w) {
                const int b = (int)blockIdx.x + (pr + w) * (int)gridDim.x;
                const float rho = (float)p.rho[b];
                const float irho = 1.0f / rho, kap = rho / (rho + 1.0f);
                const float d_off = 1.0f / (0.0f + 2.0f * rho), d_on = 1.0f / (1.0f + 2.0f * rho);
                const float thr = (float)(p.tauS[b] / p.rho[b]);
                const float sc = in.scale[b];
                const cx<float>* Ws = Wsm + w * N * N;
                const cx<float>* As = Asm + w * 2 * N * N;
                const cx<float>* AHs = As + N * N;
                const float alpha_prev = w ? ap1 : ap0;
                cx<float>* Xg = p.X + (size_t)b * NM;
                cx<float>* V1g = p.V1 + (size_t)b * NM;
                cx<float>* V2g = p.V2 + (size_t)b * NM;
                cx<float>* XVg = in.XV + (size_t)b * NM;
                cx<float>* Gg = in.Gm + (size_t)b * NM;
                const bool more = it + 1 < imax;
                if (it > 0) { mbar_wait(&w_ready[w], w_n[w] & 1); ++w_n[w]; }
                // ================= phase F =================
                MEGA_STAMP(0, tid == 0);
                double gram_r = 0.0, gram_i = 0.0;            // Gram entry (4 (combo % 4) + q / 4, 4 (combo / 4) + q % 4), combo = tid / 16, q = tid % 16
                for (int c = 0; c < nch; ++c) {
                    const int c0 = c * MC;
                    const unsigned ombits = (unsigned)in.omask[(size_t)b * M + c0 + m] >> n0;
                    cx<float> xo[NH], v1[NH];
                    if (tid == 0) MEGA_BEACON(1, it * 1000 + 100 + c);
                    MEGA_STAMP2(0);
                    // ---- X, V1 -> Z = X - V1/rho (SVT input, .m:35) ----
                    { MEGA_WAIT_BEGIN; in_wait(in_n); in_wait(in_n + 1); MEGA_WAIT_END(0); }
                    MEGA_STAMP2(1);
                    tile_read8(in_slot(in_n), m, half, xo); tile_read8(in_slot(in_n + 1), m, half, v1);
                    ring_release(&in_empty[in_n % NIN], &in_empty[(in_n + 1) % NIN], lane);
                    in_n += 2;
                    if (MEGA_DUMP_ON) { float sx = 0.f, sv = 0.f; for (int r = 0; r < NH; ++r) { sx += xo[r].re * xo[r].re + xo[r].im * xo[r].im; sv += v1[r].re * v1[r].re + v1[r].im * v1[r].im; } MEGA_DUMP_ADD(0 + c, sx); MEGA_DUMP_ADD(8 + c, sv); }
                    if (MEGA_DUMP_ON && c == 0 && tid < 256) { reinterpret_cast<float*>(p.dbg)[(size_t)b * 1024 + 512 + 2 * tid] = reinterpret_cast<const float*>(Ws)[tid]; reinterpret_cast<float*>(p.dbg)[(size_t)b * 1024 + 512 + 2 * tid + 1] = reinterpret_cast<const float*>(Ws)[N * N + tid]; if (tid == 0) { reinterpret_cast<float*>(p.dbg)[(size_t)b * 1024 + 100] = alpha_prev; reinterpret_cast<float*>(p.dbg)[(size_t)b * 1024 + 101] = rho; } }
                    wsync();                                  // the Gram of the previous chunk has read Z
#pragma unroll
                    for (int r = 0; r < NH; ++r) { Zre[(n0 + r) * MZP + m] = xo[r].re - irho * v1[r].re; Zim[(n0 + r) * MZP + m] = xo[r].im - irho * v1[r].im; }
                    wsync();
                    // ---- Y = W Z ; u = V1 + rho Y ----
                    {
                        // packed fp32x2 FMAs on row pairs (W is held planar, so (re_r, re_r+1) is one 64-bit operand): per k
                        //   YR += WR * zr ; YR += WI * (-zi) ; YI += WR * zi ; YI += WI * zr   - the operations and order of cmac(), two rows per instruction
                        float y_r[NH], y_i[NH];
                        {
                            const float* WR = reinterpret_cast<const float*>(Ws) + n0; const float* WI = WR + N * N;
                            uint64_t YR[NH / 2], YI[NH / 2];
#pragma unroll
                            for (int q = 0; q < NH / 2; ++q) { YR[q] = 0ull; YI[q] = 0ull; }
#pragma unroll 4
                            for (int k = 0; k < N; ++k) {
                                const float zr = Zre[k * MZP + m], zi = Zim[k * MZP + m];
                                const uint64_t ZR = pack2(zr, zr), ZI = pack2(zi, zi), ZN = pack2(-zi, -zi);
                                const ulonglong2 r0 = *reinterpret_cast<const ulonglong2*>(WR + N * k), r1 = *reinterpret_cast<const ulonglong2*>(WR + N * k + 4);
                                const ulonglong2 i0 = *reinterpret_cast<const ulonglong2*>(WI + N * k), i1 = *reinterpret_cast<const ulonglong2*>(WI + N * k + 4);
                                const uint64_t wr[4] = {r0.x, r0.y, r1.x, r1.y}, wi[4] = {i0.x, i0.y, i1.x, i1.y};
#pragma unroll
                                for (int q = 0; q < NH / 2; ++q) { ffma2(YR[q], wr[q], ZR); ffma2(YR[q], wi[q], ZN); ffma2(YI[q], wr[q], ZI); ffma2(YI[q], wi[q], ZR); }
                            }
#pragma unroll
                            for (int q = 0; q < NH / 2; ++q) { unpack2(YR[q], y_r[2 * q], y_r[2 * q + 1]); unpack2(YI[q], y_i[2 * q], y_i[2 * q + 1]); }
                        }
                        if (!more && p.Yout != nullptr) {
#pragma unroll
                            for (int hf = 0; hf < NH / 4; ++hf) {
                                cx<float> t[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) t[u] = mk<float>(y_r[4 * hf + u], y_i[4 * hf + u]);
                                st4c<float>(p.Yout + (long long)b * p.ld_Y + (size_t)(c0 + m) * N + n0 + 4 * hf, t);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < NH; ++r) { v1[r].re += rho * y_r[r]; v1[r].im += rho * y_i[r]; }
                    }
                    if (tid == 0) MEGA_BEACON(1, it * 1000 + 110 + c);
                    MEGA_STAMP2(2);
                    // ---- pass 1 result ----
                    float xs_r[NH], xs_i[NH];
                    if (it > 0) {
                        mbar_wait(&acc_full[0], accn[0] & 1); ++accn[0];
                        tc::tc_fence_after();
                        read_acc(ACC0, lane_base, n0, L, xs_r, xs_i);
                        tc::tc_fence_before();
                        warp_arrive(&acc_empty[0], lane);
                    } else {
#pragma unroll
                        for (int r = 0; r < NH; ++r) { xs_r[r] = 0.f; xs_i[r] = 0.f; }
                    }
                    MEGA_STAMP2(3);
                    if (MEGA_DUMP_ON) { float sx = 0.f; for (int r = 0; r < NH; ++r) sx += xs_r[r] * xs_r[r] + xs_i[r] * xs_i[r]; MEGA_DUMP_ADD(16 + c, sx); }
                    // ---- V2, subY -> C, V2, X, V1, K ----
                    cx<float> v2[NH], kt[NH];
                    {
                        cx<float> sy[NH];
                        { MEGA_WAIT_BEGIN; in_wait(in_n); in_wait(in_n + 1); MEGA_WAIT_END(1); }
                        tile_read8(in_slot(in_n), m, half, v2); tile_read8(in_slot(in_n + 1), m, half, sy);
                        if (MEGA_DUMP_ON) { float sx = 0.f, sv = 0.f; for (int r = 0; r < NH; ++r) { sx += v2[r].re * v2[r].re + v2[r].im * v2[r].im; sv += sy[r].re * sy[r].re + sy[r].im * sy[r].im; } MEGA_DUMP_ADD(24 + c, sx); MEGA_DUMP_ADD(32 + c, sv); }
                        ring_release(&in_empty[in_n % NIN], &in_empty[(in_n + 1) % NIN], lane);
                        in_n += 2;
#pragma unroll
                        for (int r = 0; r < NH; ++r) {
                            // C = rho/(rho+1) (X - Xs - V2/rho) ; V2 += rho (C - X + Xs)      (.m:61,65 of the previous iteration; all zero at i = 1)
                            const float wr = xo[r].re - xs_r[r], wi = xo[r].im - xs_i[r];
                            const float cr = kap * (wr - irho * v2[r].re), ci = kap * (wi - irho * v2[r].im);
                            v2[r].re += rho * (cr - wr); v2[r].im += rho * (ci - wi);
                            const float d = ((ombits >> r) & 1u) ? d_on : d_off;                                                          // iK1 (.m:20): 1 / (Omega + 2 rho), Omega in {0, 1}
                            const float xr = (v1[r].re + sy[r].re + v2[r].re + rho * cr + rho * xs_r[r]) * d;                             // .m:38-40
                            const float xi = (v1[r].im + sy[r].im + v2[r].im + rho * ci + rho * xs_i[r]) * d;
                            v1[r].re -= rho * xr; v1[r].im -= rho * xi;                                                                   // .m:64: V1 + rho (Y - X)
                            kt[r] = mk<float>(xr - irho * v2[r].re - cr, xi - irho * v2[r].im - ci);                                      // .m:43
                            xo[r] = mk<float>(xr, xi);
                        }
                    }
                    stg8(Xg + (size_t)(c0 + m) * N + n0, xo, narrow);
                    stg8(V1g + (size_t)(c0 + m) * N + n0, v1, narrow);
                    stg8(V2g + (size_t)(c0 + m) * N + n0, v2, narrow);
#pragma unroll
                    for (int r = 0; r < NH; ++r) { xo[r].re -= irho * v1[r].re; xo[r].im -= irho * v1[r].im; }      // zn = X - V1/rho
                    MEGA_STAMP2(4);
                    // ---- XV = A V B (carried: XV += alpha G of the previous iteration) -> K - XV, the operand of pass 2 (.m:47) ----
                    {
                        cx<float> xv[NH], g[NH];
                        { MEGA_WAIT_BEGIN; in_wait(in_n); in_wait(in_n + 1); MEGA_WAIT_END(2); }
                        tile_read8(in_slot(in_n), m, half, xv); tile_read8(in_slot(in_n + 1), m, half, g);
                        if (MEGA_DUMP_ON) { float sx = 0.f, sv = 0.f; for (int r = 0; r < NH; ++r) { sx += xv[r].re * xv[r].re + xv[r].im * xv[r].im; sv += g[r].re * g[r].re + g[r].im * g[r].im; } MEGA_DUMP_ADD(40 + c, sx); MEGA_DUMP_ADD(48 + c, sv); }
                        ring_release(&in_empty[in_n % NIN], &in_empty[(in_n + 1) % NIN], lane);
                        in_n += 2;
#pragma unroll
                        for (int r = 0; r < NH; ++r) {
                            xv[r].re = fmaf(alpha_prev, g[r].re, xv[r].re); xv[r].im = fmaf(alpha_prev, g[r].im, xv[r].im);
                            kt[r].re -= xv[r].re; kt[r].im -= xv[r].im;
                        }
                        stg8(XVg + (size_t)(c0 + m) * N + n0, xv, narrow);
                    }
                    if (tid == 0) MEGA_BEACON(1, it * 1000 + 120 + c);
                    MEGA_STAMP2(5);
                    // ---- pass-2 small operand: rows (split, n, c), k = m ----
                    if (kop_n > 0) mbar_wait(kop_empty, (kop_n - 1) & 1);
                    ++kop_n;
                    {   // MN-major: the 8 rows of a row group are contiguous, so this thread's 16 rows of one split term are two 16-byte stores
                        unsigned char* kb = kop + (size_t)(m / 8) * KLBO + (m % 8) * 16;
                        unsigned short sp[2 * NH][3];
#pragma unroll
                        for (int r = 0; r < NH; ++r) { split3(kt[r].re, sp[2 * r]); split3(kt[r].im, sp[2 * r + 1]); }
#pragma unroll
                        for (int u = 0; u < 3; ++u) {
                            const int rg = (32 * u + 2 * n0) / 8;
#pragma unroll
                            for (int hh = 0; hh < 2; ++hh) {
                                uint32_t w[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) w[j] = (uint32_t)sp[8 * hh + 2 * j][u] | ((uint32_t)sp[8 * hh + 2 * j + 1][u] << 16);
                                *reinterpret_cast<uint4*>(kb + (rg + hh) * 128) = make_uint4(w[0], w[1], w[2], w[3]);
                            }
                        }
                    }
                    tc::fence_async_smem();
                    warp_arrive(kop_full, lane);
                    if (tid == 0) MEGA_BEACON(1, it * 1000 + 130 + c);
                    MEGA_STAMP2(6);
                    // ---- Gram of the next SVT input ----
                    wsync();                                  // W Z of this chunk has read Z
#pragma unroll
                    for (int r = 0; r < NH; ++r) { Zre[(n0 + r) * MZP + m] = xo[r].re; Zim[(n0 + r) * MZP + m] = xo[r].im; }
                    wsync();
                    {
                        const int combo = tid / 16, slice = tid % 16, ib = combo % 4, jb = combo / 4;
                        float acc[32];
#pragma unroll
                        for (int e = 0; e < 32; ++e) acc[e] = 0.f;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {         // slice s owns columns [64 j + 4 s, + 4): a half-warp reads 256 contiguous bytes per row
                            const int cc = 64 * j + 4 * slice;
                            float4 xr[4], xi[4], yr[4], yi[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                xr[u] = *reinterpret_cast<const float4*>(Zre + (ib * 4 + u) * MZP + cc); xi[u] = *reinterpret_cast<const float4*>(Zim + (ib * 4 + u) * MZP + cc);
                                yr[u] = *reinterpret_cast<const float4*>(Zre + (jb * 4 + u) * MZP + cc); yi[u] = *reinterpret_cast<const float4*>(Zim + (jb * 4 + u) * MZP + cc);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int v = 0; v < 4; ++v) {
                                    float& ar = acc[2 * (4 * u + v)]; float& ai = acc[2 * (4 * u + v) + 1];
                                    cmac<float>(ar, ai, xr[u].x, xi[u].x, yr[v].x, -yi[v].x); cmac<float>(ar, ai, xr[u].y, xi[u].y, yr[v].y, -yi[v].y);
                                    cmac<float>(ar, ai, xr[u].z, xi[u].z, yr[v].z, -yi[v].z); cmac<float>(ar, ai, xr[u].w, xi[u].w, yr[v].w, -yi[v].w);
                                }
                        }
                        // transposed butterfly over the 16 slices: 32 -> 16 -> 8 -> 4 -> 2 values; lane q = tid % 16 ends with entry (u, v) = (q / 4, q % 4)
#pragma unroll
                        for (int i = 0; i < 16; ++i) { const bool up = lane & 8; const float snd = up ? acc[i] : acc[i + 16], kp = up ? acc[i + 16] : acc[i]; acc[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 8); }
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const bool up = lane & 4; const float snd = up ? acc[i] : acc[i + 8], kp = up ? acc[i + 8] : acc[i]; acc[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 4); }
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const bool up = lane & 2; const float snd = up ? acc[i] : acc[i + 4], kp = up ? acc[i + 4] : acc[i]; acc[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 2); }
#pragma unroll
                        for (int i = 0; i < 2; ++i) { const bool up = lane & 1; const float snd = up ? acc[i] : acc[i + 2], kp = up ? acc[i + 2] : acc[i]; acc[i] = kp + __shfl_xor_sync(0xffffffffu, snd, 1); }
                        gram_r += (double)acc[0]; gram_i += (double)acc[1];
                    }
                    MEGA_STAMP2(7);
                }
                MEGA_STAMP(1, tid == 0);
                __threadfence();                              // X, V1, V2, XV stores performed at L2 (a CTA-scope release stops at L1; TMA reads L2)
                fence_proxy_async_all();                      // X, V1, V2, XV stores -> the TMA loads of the next iteration
                if (more) {                                   // hand the Gram matrix to the Jacobi warps (through L2: they may still be busy with the other trial)
                    const int combo = tid / 16, q = tid % 16, gi = 4 * (combo % 4) + q / 4, gj = 4 * (combo / 4) + q % 4;
                    reinterpret_cast<double2*>(p.gram + (size_t)b * p.nmc * 2 * N * N)[gi + N * gj] = make_double2(gram_r, gram_i);
                    __threadfence();                          // the Jacobi warps read it from L2 (ld.cg): the store must have been performed there
                    warp_arrive(&gram_ready[w], lane);
                }
                // ================= phase R: Res_l = A'(scale T1'_l Dt), |Res|^2, operand image of G =================
                if (tid == 0) MEGA_BEACON(1, it * 1000 + 199);
                mbar_wait(t1_full, t1_n & 1); ++t1_n;
                tc::tc_fence_after();
                if (tid == 0) MEGA_BEACON(1, it * 1000 + 200);
                MEGA_STAMP(2, tid == 0);
                double rr = 0.0;
                for (int l = 0; l < L; ++l) {
                    MEGA_STAMP3(11, 0);
                    {
                        float acc[16], a1[16], a2[16];
                        const uint32_t D = tm + NS * l + lane_base + 2 * n0;
                        tc::tmem_ld16x3(D + 64, D + 32, D, acc, a1, a2);
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[j] = (acc[j] + a1[j]) + a2[j];          // lo + mid, then hi
                        // lane = kc: even lanes hold the e_re rows, odd lanes the e_im rows of the same antenna
                        //   T1'r = D[(k,re),(n,re)] + D[(k,im),(n,im)] ; T1'i = D[(k,re),(n,im)] - D[(k,im),(n,re)]
#pragma unroll
                        for (int r = 0; r < NH; ++r) {
                            const float vre = acc[2 * r], vim = acc[2 * r + 1];
                            const float other = __shfl_xor_sync(0xffffffffu, vim, 1);
                            reinterpret_cast<float*>(U)[2 * ((n0 + r) + LDU * (m >> 1)) + (m & 1)] = (lane & 1) ? (other - vre) : (vre + other);
                        }
                    }
                    tc::tc_fence_before();
                    wsync();
                    MEGA_STAMP3(11, 1);
                    apply_a_shared(AHs, U, V);                  // A' T1'_l
                    wsync();
                    apply_a_shared(As, V, U);                   // A A' T1'_l = (A Res_l) Dt' / scale^2 (Dt unitary)
                    wsync();
                    MEGA_STAMP3(11, 2);
                    {   // operand image of tap l, written where phase G reads it (the state ring is idle between the phases F)
                        const int n = tid % N, kq = tid / N;
                        cx<float> q[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) q[j] = U[n + LDU * (4 * kq + j)];
                        unsigned char* img = inr + (size_t)l * QTAP;
                        put_q4(n, 4 * kq, q, sc * sc, (l & 1) ? 2 : 0, [&](uint32_t off, uint4 a, uint4 bq) {
                            *reinterpret_cast<uint4*>(img + off) = a; *reinterpret_cast<uint4*>(img + off + 16) = bq;
                        });
                    }
                    wsync();                                  // U is free
                    MEGA_STAMP3(11, 3);
                    fft64<false, WSync>(V, U, tw, 0.125f * sc);   // Res_l = (A' T1'_l) Dt
                    wsync();
                    MEGA_STAMP3(11, 4);
                    cx<float>* Res = p.Res + (size_t)b * G * p.P + (size_t)G * NT * l;
                    for (int t = tid; t < N * NT; t += WTHREADS) {
                        const int r = t % N, cc = t / N;
                        const cx<float> v = U[r + LDU * cc];
                        if (r < G) { Res[r + (size_t)G * cc] = v; rr += (double)v.re * v.re + (double)v.im * v.im; }
                    }
                    wsync();                                  // U is free again
                    MEGA_STAMP3(11, 5);
                }
                tc::fence_async_smem();                       // image (generic stores) -> MMA operand reads (async proxy)
                warp_arrive(img_ready, lane);
                warp_arrive(r_done, lane);
                if (tid == 0) MEGA_BEACON(1, it * 1000 + 300);
                MEGA_STAMP(3, tid == 0);
                // Res / V of tap 0 for phase S are requested now (V is last iteration's, Res is complete): their L2 latency hides behind phase G
                const size_t off0 = (size_t)b * G * p.P;
                cx<float> rn[4], vn[4];                       // Res / V of the next tap travel while this tap is processed
                const int er = tid % N, ec = tid / N;          // this thread's elements of a tap: row er, columns ec + 16 j
                auto fetch = [&](int l) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int t = er + G * (ec + 16 * j);
                        if (er < G) { rn[j] = p.Res[off0 + (size_t)G * NT * l + t]; vn[j] = p.V[off0 + (size_t)G * NT * l + t]; }
                    }
                };
                fetch(0);
                // ================= phase G: G = (A Res) e per chunk, |G|^2 =================
                double gg = 0.0;
                for (int c = 0; c < nch; ++c) {
                    const int buf = c & 1;
                    mbar_wait(&acc_full[buf], accn[buf] & 1); ++accn[buf];
                    tc::tc_fence_after();
                    float gr[NH], gi[NH];
                    read_acc(buf ? ACC1 : ACC0, lane_base, n0, L, gr, gi);
                    tc::tc_fence_before();
                    warp_arrive(&acc_empty[buf], lane);
                    cx<float> g[NH];
#pragma unroll
                    for (int r = 0; r < NH; ++r) { g[r] = mk<float>(gr[r], gi[r]); gg += (double)gr[r] * gr[r] + (double)gi[r] * gi[r]; }
                    if (more) stg8(Gg + (size_t)(c * MC + m) * N + n0, g, narrow);
                }
                for (int o = 16; o > 0; o >>= 1) { gg += __shfl_down_sync(0xffffffffu, gg, o); rr += __shfl_down_sync(0xffffffffu, rr, o); }
                if (lane == 0) { misc[warp] = gg; misc[8 + warp] = rr; }
                __threadfence();                              // G stores performed at L2, where the TMA loads of the next iteration read them ...
                fence_proxy_async_all();                      // ... and ordered for the async proxy
                warp_arrive(g_done_q, lane);
                warp_arrive(g_done, lane);                    // all MMAs of phase G have completed (last acc_full): the ring is free for the next phase F
                wsync();
                {
                    double sg = 0.0, sr = 0.0;
#pragma unroll
                    for (int w = 0; w < NWK; ++w) { sg += misc[w]; sr += misc[8 + w]; }
                    gg = sg; rr = sr;
                }
                // ================= phase S: alpha, V, S, operand image of the next Xs =================
                if (tid == 0) MEGA_BEACON(1, it * 1000 + 400);
                MEGA_STAMP(4, tid == 0);
                const float alpha = gg > 0.0 ? (float)(rr / gg) : 0.f;          // res'res / (res' R res)  (.m:48)
                if (p.angles) {                               // Omega_S(indx_S(1 : min(10 + 5 i, G P))) = 1  (_angles.m:36), i = it + 1
                    int hi = 10 + 5 * (it + 1); if (hi > G * p.P) hi = G * p.P; if (hi > p.n_indx) hi = p.n_indx;
                    const int lo = it == 0 ? 0 : 10 + 5 * it;
                    const int* idx = p.indx + (long long)b * p.ld_indx;
                    unsigned char* mk_ = p.smask + (size_t)b * G * p.P;
                    for (int k = lo + tid; k < hi; k += WTHREADS) { const int v = idx[k]; if (v >= 1 && v <= G * p.P) mk_[v - 1] = 1; }
                    wsync();
                }
                {
                    for (int l = 0; l < L; ++l) {
                        MEGA_STAMP3(12, 0);
                        const size_t off = off0 + (size_t)G * NT * l;
                        const unsigned char* mask = p.angles ? p.smask + off : nullptr;
                        cx<float> rc[4], vc[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { rc[j] = rn[j]; vc[j] = vn[j]; }
                        if (l + 1 < L) fetch(l + 1);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int t = er + G * (ec + 16 * j);
                            cx<float> sv = mk<float>(0.f, 0.f);
                            if (er < G) {
                                const cx<float> v = mk<float>(vc[j].re + alpha * rc[j].re, vc[j].im + alpha * rc[j].im);
                                sv = mk<float>(soft1<float>(v.re, thr), soft1<float>(v.im, thr));
                                if (mask && !mask[t]) sv = mk<float>(0.f, 0.f);
                                p.V[off + t] = v;
                                if (!more) p.S[off + t] = sv;
                            }
                            if (more) U[er + LDU * (ec + 16 * j)] = sv;       // rows >= G of the padded work matrix are zero
                        }
                        if (!more) continue;
                        wsync();
                        MEGA_STAMP3(12, 1);
                        apply_a_shared(As, U, V);               // A S_l                           (.m:58, left factor)
                        wsync();
                        MEGA_STAMP3(12, 2);
                        fft64<true, WSync>(V, U, tw, 0.125f * sc);    // scale (A S_l) Dt'
                        wsync();
                        MEGA_STAMP3(12, 3);
                        {   // operand image of tap l of the next pass 1, 32-byte stores straight to the image in global memory
                            const int n = tid % N, kq = tid / N;
                            cx<float> q[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) q[j] = U[n + LDU * (4 * kq + j)];
                            unsigned char* img = in.QopS + ((size_t)b * L + l) * QTAP;
                            put_q4(n, 4 * kq, q, 1.0f, (l & 1) ? 2 : 0, [&](uint32_t o, uint4 a, uint4 bq) { stg_u4x2(img + o, a, bq); });
                        }
                        wsync();                              // U is free
                        MEGA_STAMP3(12, 4);
                    }
                }
                if (more) {
                    __threadfence();                          // image stores performed at L2 before the bulk copies of the next pass 1 read them
                    fence_proxy_async_all();
                    warp_arrive(&q_ready[w], lane);
                }
                if (w) ap1 = alpha; else ap0 = alpha;
                if (tid == 0) MEGA_BEACON(1, it * 1000 + 500);
                MEGA_STAMP(5, tid == 0);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == W_LD) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

}  // namespace mega
}  // namespace jstsp
