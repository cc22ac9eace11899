// ls.cu - the least-squares baseline of the reference's drivers, batched on the device:
//   S_ls = pinv(A) * Y * pinv(B)                        plot_errorVSsnr.m:83, plot_errorVSsnr_approx.m:61,67
//   Y * pinv(B)  (right-hand sides of the joint OMP)    plot_errorVSsnr.m:117
// A is N x G, Y is N x M, B is P x M, S is G x P.  pinv of a full-rank matrix is formed through the Hermitian positive definite Gram
// matrix of its short side (pinv(A) = inv(A'A) A' for G <= N, A' inv(A A') otherwise; likewise for B), inverted by Gauss-Jordan
// elimination in fp64-accumulated arithmetic of the call's type; every product is a batched complex GEMM.  Rank-deficient operands
// (where MATLAB's pinv truncates singular values) are outside this entry point: the elimination then produces non-finite values, which
// are counted per trial and returned like every solver's non-finite count.
#include "common.cuh"

namespace jstsp {

enum { OP_N = 0, OP_H = 1 };

// C (m x n) = op(A) (m x k) * op(B) (k x n), column-major, per-trial strides (0 = shared); 16 x 16 outputs per CTA
template <typename T>
__global__ void __launch_bounds__(256) k_bgemm(int m, int n, int k, const cx<T>* A, long long sA, int ldA, int opA, const cx<T>* B, long long sB, int ldB, int opB,
                                               cx<T>* C, long long sC, int ldC) {
    __shared__ cx<T> As[16][17], Bs[16][17];
    const int b = blockIdx.z, tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int i = blockIdx.x * 16 + tx, j = blockIdx.y * 16 + ty;
    const cx<T>* Ab = A + (long long)b * sA; const cx<T>* Bb = B + (long long)b * sB;
    T re = 0, im = 0;
    for (int k0 = 0; k0 < k; k0 += 16) {
        {   // As[tx][ty] = op(A)(i0 + tx, k0 + ty)
            const int r = blockIdx.x * 16 + tx, c = k0 + ty;
            cx<T> v = mk<T>(T(0), T(0));
            if (r < m && c < k) v = opA == OP_N ? Ab[r + (size_t)ldA * c] : conj(Ab[c + (size_t)ldA * r]);
            As[tx][ty] = v;
        }
        {   // Bs[tx][ty] = op(B)(k0 + tx, j0 + ty)
            const int r = k0 + tx, c = blockIdx.y * 16 + ty;
            cx<T> v = mk<T>(T(0), T(0));
            if (r < k && c < n) v = opB == OP_N ? Bb[r + (size_t)ldB * c] : conj(Bb[c + (size_t)ldB * r]);
            Bs[tx][ty] = v;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 16; ++q) cmac<T>(re, im, As[tx][q].re, As[tx][q].im, Bs[q][ty].re, Bs[q][ty].im);
        __syncthreads();
    }
    if (i < m && j < n) C[(long long)b * sC + i + (size_t)ldC * j] = mk<T>(re, im);
}

// in-place Gauss-Jordan inverse of an n x n Hermitian positive definite matrix in global memory, one CTA per matrix
template <typename T>
__global__ void __launch_bounds__(256) k_hpd_inv(cx<T>* mats, long long stride, int n) {
    cx<T>* a = mats + (long long)blockIdx.x * stride;
    extern __shared__ __align__(16) unsigned char smem[];
    cx<T>* colk = reinterpret_cast<cx<T>*>(smem);
    cx<T>* rowk = colk + n;
    __shared__ double pr, pi;
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) { colk[i] = a[i + (size_t)n * k]; rowk[i] = a[k + (size_t)n * i]; }
        __syncthreads();
        if (threadIdx.x == 0) { const cx<T> pv = colk[k]; const double d = (double)pv.re * pv.re + (double)pv.im * pv.im; pr = pv.re / d; pi = -pv.im / d; }
        __syncthreads();
        const cx<T> ip = mk<T>((T)pr, (T)pi);
        for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
            const int i = t % n, j = t / n;
            cx<T> v;
            if (i == k && j == k) v = ip;
            else if (i == k) v = rowk[j] * ip;
            else if (j == k) v = mk<T>(T(0), T(0)) - colk[i] * ip;
            else v = a[t] - colk[i] * (rowk[j] * ip);
            a[t] = v;
        }
    }
}

template <typename T>
__global__ void k_ls_nonfinite(const cx<T>* S, size_t per, int batch, int* flag) {
    const int b = blockIdx.x;
    int bad = 0;
    for (size_t t = threadIdx.x; t < per; t += blockDim.x) { const cx<T> v = S[(size_t)b * per + t]; if (!isfinite((double)v.re) || !isfinite((double)v.im)) bad = 1; }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0 && bad) atomicAdd(flag, 1);
}

template <typename T>
static void gemm(Handle* h, int m, int n, int k, const cx<T>* A, long long sA, int ldA, int opA, const cx<T>* B, long long sB, int ldB, int opB, cx<T>* C, long long sC, int ldC, int batch) {
    dim3 g(ceil_div(m, 16), ceil_div(n, 16), batch);
    JSTSP_LAUNCH(h, PK_OTHER, (k_bgemm<T><<<g, 256, 0, h->stream>>>(m, n, k, A, sA, ldA, opA, B, sB, ldB, opB, C, sC, ldC)));
}

template <typename T>
static int run_ls(Handle* h, int mem, int N, int M, int G, int P, int batch, const void* A_, long long ld_A, const void* B_, long long ld_B, const void* Y_, long long ld_Y,
                  void* S_, long long ld_S, void* YpB_, long long ld_YpB) {
    if (N <= 0 || M <= 0 || G <= 0 || P <= 0 || batch <= 0 || !B_ || !Y_ || (!S_ && !YpB_) || (S_ && !A_)) return fail(h, JSTSP_E_ARG, "bad argument");
    const bool host = mem == JSTSP_HOST;
    const size_t NG = (size_t)N * G, PM = (size_t)P * M, NM = (size_t)N * M, GP = (size_t)G * P, NP = (size_t)N * P;
    const bool sharedA = ld_A == 0, sharedB = ld_B == 0;
    const int ga = G <= N ? G : N, gb = P <= M ? P : M;            // sizes of the two Gram matrices
    if (!ld_Y) ld_Y = NM; if (!ld_S) ld_S = GP; if (!ld_YpB) ld_YpB = NP;
    const int nA = sharedA ? 1 : batch, nB = sharedB ? 1 : batch;
    const size_t esz = sizeof(cx<T>);
    size_t need = esz * ((size_t)nA * ga * ga + (size_t)nB * gb * gb + (size_t)batch * ((size_t)G * M + (size_t)N * M + (size_t)(G > N ? G : N) * (P > M ? P : M) + GP + NP)) + 8192;
    if (host) need += esz * (nA * NG + nB * PM + (size_t)batch * NM);
    int rc = ensure_workspace(h, need); if (rc) return rc;
    Arena ar(h->ws, h->ws_bytes);
    cx<T>* GA = ar.take<cx<T>>((size_t)nA * ga * ga);
    cx<T>* GB = ar.take<cx<T>>((size_t)nB * gb * gb);
    cx<T>* T1 = ar.take<cx<T>>((size_t)batch * G * M);            // pinv(A) Y
    cx<T>* T0 = ar.take<cx<T>>((size_t)batch * N * M);            // scratch of the wide-A route
    cx<T>* T2 = ar.take<cx<T>>((size_t)batch * (size_t)(G > N ? G : N) * (P > M ? P : M));
    cx<T>* dS = ar.take<cx<T>>((size_t)batch * GP);
    cx<T>* dYpB = ar.take<cx<T>>((size_t)batch * NP);
    const cx<T>* dA = (const cx<T>*)A_; const cx<T>* dB = (const cx<T>*)B_; const cx<T>* dY = (const cx<T>*)Y_;
    cudaStream_t st = h->stream;
    if (host) {
        auto up = [&](const void* src, size_t per, long long ld, int cnt) -> const cx<T>* {
            cx<T>* d = ar.take<cx<T>>(per * cnt);
            if (src) cudaMemcpy2DAsync(d, per * esz, src, (size_t)(ld ? ld : per) * esz, per * esz, cnt, cudaMemcpyHostToDevice, st);
            return d;
        };
        dA = A_ ? up(A_, NG, ld_A, nA) : nullptr; dB = up(B_, PM, ld_B, nB); dY = up(Y_, NM, ld_Y, batch);
        if (ld_A) ld_A = NG; if (ld_B) ld_B = PM; ld_Y = NM;
    }
    JSTSP_CUDA(h, cudaMemsetAsync(h->d_flag, 0, sizeof(int), st));
    const size_t smi = 2 * esz * (size_t)(ga > gb ? ga : gb);
    rc = set_smem(h, k_hpd_inv<T>, smi); if (rc) return rc;
    // ---- right factor: R(X) = X pinv(B), X with `rows` rows ----
    if (P <= M) gemm<T>(h, P, P, M, dB, ld_B, P, OP_N, dB, ld_B, P, OP_H, GB, (long long)gb * gb, gb, nB);        // B B'
    else gemm<T>(h, M, M, P, dB, ld_B, P, OP_H, dB, ld_B, P, OP_N, GB, (long long)gb * gb, gb, nB);               // B' B
    JSTSP_LAUNCH(h, PK_OTHER, (k_hpd_inv<T><<<nB, 256, smi, st>>>(GB, (long long)gb * gb, gb)));
    auto right = [&](const cx<T>* X, long long sX, int rows, cx<T>* out, long long sOut) {
        if (P <= M) {    // (X B') inv(B B')
            gemm<T>(h, rows, P, M, X, sX, rows, OP_N, dB, ld_B, P, OP_H, T2, (long long)rows * P, rows, batch);
            gemm<T>(h, rows, P, P, T2, (long long)rows * P, rows, OP_N, GB, sharedB ? 0 : (long long)gb * gb, gb, OP_N, out, sOut, rows, batch);
        } else {         // (X inv(B' B)) B'
            gemm<T>(h, rows, M, M, X, sX, rows, OP_N, GB, sharedB ? 0 : (long long)gb * gb, gb, OP_N, T2, (long long)rows * M, rows, batch);
            gemm<T>(h, rows, P, M, T2, (long long)rows * M, rows, OP_N, dB, ld_B, P, OP_H, out, sOut, rows, batch);
        }
    };
    if (YpB_) right(dY, ld_Y, N, host ? dYpB : (cx<T>*)YpB_, host ? (long long)NP : ld_YpB);
    if (S_) {
        // ---- left factor: pinv(A) Y ----
        if (G <= N) {
            gemm<T>(h, G, G, N, dA, ld_A, N, OP_H, dA, ld_A, N, OP_N, GA, (long long)ga * ga, ga, nA);            // A' A
            JSTSP_LAUNCH(h, PK_OTHER, (k_hpd_inv<T><<<nA, 256, smi, st>>>(GA, (long long)ga * ga, ga)));
            gemm<T>(h, G, M, N, dA, ld_A, N, OP_H, dY, ld_Y, N, OP_N, T0, (long long)G * M, G, batch);            // A' Y   (G x M fits the N x M scratch)
            gemm<T>(h, G, M, G, GA, sharedA ? 0 : (long long)ga * ga, ga, OP_N, T0, (long long)G * M, G, OP_N, T1, (long long)G * M, G, batch);
        } else {
            gemm<T>(h, N, N, G, dA, ld_A, N, OP_N, dA, ld_A, N, OP_H, GA, (long long)ga * ga, ga, nA);            // A A'
            JSTSP_LAUNCH(h, PK_OTHER, (k_hpd_inv<T><<<nA, 256, smi, st>>>(GA, (long long)ga * ga, ga)));
            gemm<T>(h, N, M, N, GA, sharedA ? 0 : (long long)ga * ga, ga, OP_N, dY, ld_Y, N, OP_N, T0, (long long)NM, N, batch);
            gemm<T>(h, G, M, N, dA, ld_A, N, OP_H, T0, (long long)NM, N, OP_N, T1, (long long)G * M, G, batch);
        }
        right(T1, (long long)G * M, G, host ? dS : (cx<T>*)S_, host ? (long long)GP : ld_S);
        JSTSP_LAUNCH(h, PK_OTHER, (k_ls_nonfinite<T><<<batch, 128, 0, st>>>(host ? dS : (const cx<T>*)S_, host ? GP : (size_t)ld_S, batch, h->d_flag)));
    }
    JSTSP_CUDA(h, cudaGetLastError());
    int bad = 0;
    if (host) {
        if (S_) JSTSP_CUDA(h, cudaMemcpy2DAsync(S_, (size_t)ld_S * esz, dS, GP * esz, GP * esz, batch, cudaMemcpyDeviceToHost, st));
        if (YpB_) JSTSP_CUDA(h, cudaMemcpy2DAsync(YpB_, (size_t)ld_YpB * esz, dYpB, NP * esz, NP * esz, batch, cudaMemcpyDeviceToHost, st));
        JSTSP_CUDA(h, cudaMemcpyAsync(&bad, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        JSTSP_CUDA(h, cudaStreamSynchronize(st));
    }
    return bad;
}

}  // namespace jstsp
using namespace jstsp;

extern "C" int jstsp_ls_estimate(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int batch,
                                 const void* A, long long ld_A, const void* B, long long ld_B, const void* Y, long long ld_Y,
                                 void* S, long long ld_S, void* YpinvB, long long ld_YpinvB) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_ls<float>(h, mem, N, M, G, P, batch, A, ld_A, B, ld_B, Y, ld_Y, S, ld_S, YpinvB, ld_YpinvB);
    if (dtype == JSTSP_F64) return run_ls<double>(h, mem, N, M, G, P, batch, A, ld_A, B, ld_B, Y, ld_Y, S, ld_S, YpinvB, ld_YpinvB);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}

/* plot_ee.m:69-77: power of the four receiver designs in watts (the axis label says mW; the constants are kept as written). */
extern "C" int jstsp_power_model(int Nr, int Mr, int Mr_e, double* power4) {
    if (!power4 || Nr <= 0 || Mr <= 0 || Mr_e <= 0) return JSTSP_E_ARG;
    const double Pcirc = 0, Psw = 0.005, Pps = 0.015, Plna = 0.02, Pps_zc = 0.06;
    power4[0] = Pcirc + (double)Nr * Nr * Plna + (double)Nr * (Nr + 1) * Pps_zc;                          /* digital beamforming          :74 */
    power4[1] = Pcirc + (double)Mr * Nr * Plna + (double)Nr * (Mr + 1) * Pps;                             /* conventional HBF, PS         :75 */
    power4[2] = Pcirc + (double)Mr * Nr * Plna + (double)Nr * (Mr + 1) * Pps_zc;                          /* conventional HBF, ZC         :76 */
    power4[3] = Pcirc + (double)Mr_e * Nr * Plna + (double)Mr_e * Psw + (double)Nr * (Mr_e + 1) * Pps;    /* proposed                     :77 */
    return JSTSP_OK;
}
