// vamp.cu - VAMP for the AWGN linear model with a Bernoulli-Gaussian prior, exactly as configured
// by benchmark_algorithms/vamp.m:1-55 and iterated by MPbased_solvers/VAMP/VampGlmEst.m:350-521.
//
// vamp.m real-embeds the complex problem (B = [Re A, -Im A; Im A, Re A], :3-4), takes a full
// svd(B) (:32-34) and runs VampGlmEst for exactly 100 iterations (VampGlmEst.m:509-511).  Only
// U f(d) U' enters the iteration, and the embedding of a complex matrix commutes with products, so
// the same iteration runs here in COMPLEX arithmetic on A with the m x m (or n x n) complex
// eigenbasis of A A^H (A^H A) - every singular value of A appears twice in the reference's `d`
// (SURVEY.md A.5).  The spectral basis is an INPUT (the gateway obtains it from the host's svd,
// like vamp.m:32 does); the 100 iterations - five dense mat-vecs plus fused element-wise stages
// per iteration - are the hot path and run on the GPU.
//   denoiser  : SparseScaEstim(CAwgnEstimIn(0, 1/beta), beta), beta = L/(2n)   (vamp.m:23-25;
//               SparseScaEstim.m:96-165, complex branch because r1init = eps*1i, vamp.m:45)
//   likelihood: CAwgnEstimOut(b, sigma)                                         (vamp.m:30; CAwgnEstimOut.m:97-108)
// The imaginary "dust" (~1e-16) that r1init = eps*1i seeds in the reference is not propagated.
#include "common.cuh"

namespace jstsp {

constexpr double kGamMin = 1e-8, kGamMax = 1e14;        // VampGlmOpt.m:7-8
constexpr double kEps = 2.220446049250313e-16;

template <typename T>
struct VampP {
    int m, n, nit; double damp;
    const cx<T>* A; long long ld_A;          // m x n
    const cx<T>* y; long long ld_y;          // m
    const cx<T>* basis; long long ld_b;      // m<=n: U (m x m) of A A^H ; m>n: V (n x n) of A^H A
    const T* d; long long ld_d;              // eigenvalues (length min-side: m if m<=n else n)
    const double *sigma, *Lnz;
    cx<T>* x_out; long long ld_x;
    // per-trial state
    cx<T> *r1, *p1, *x1, *z2, *r2, *p2, *t0, *t1, *t2, *t3;   // n,m,n,m,n,m, and 4 scratch vectors of max(m,n)
    T* inv;                                   // length K = (m<=n ? m : n)
    double* sc;                               // [b][16] scalars: gam1x gam1z gam2x gam2z alf ...
    int it;
};
enum { S_G1X = 0, S_G1Z, S_G2X, S_G2Z, S_ALF, S_G2ZOLD, S_G1XOLD };

__device__ __forceinline__ double clipg(double g) { return fmin(fmax(g, kGamMin), kGamMax); }

template <typename T>
__device__ double block_sum(double v) {
    __shared__ double red[32];
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) s += red[w];
    return s;
}

// first half of the iteration (VampGlmEst.m:357-401): denoiser, likelihood, extrinsics, inv, alf
template <typename T>
__global__ void __launch_bounds__(256) k_vamp_pre(VampP<T> p) {
    const int b = blockIdx.x, m = p.m, n = p.n, i = p.it;
    double* sc = p.sc + (size_t)b * 16;
    cx<T>* r1 = p.r1 + (size_t)b * n; cx<T>* x1 = p.x1 + (size_t)b * n; cx<T>* r2 = p.r2 + (size_t)b * n;
    cx<T>* p1 = p.p1 + (size_t)b * m; cx<T>* p2 = p.p2 + (size_t)b * m;
    const cx<T>* y = p.y + (long long)b * p.ld_y;
    const double gam1x = sc[S_G1X], gam1z = sc[S_G1Z];
    const double N2 = 2.0 * n, beta = p.Lnz[b] / N2, var0 = 1.0 / beta, wvar = p.sigma[b];
    const double rvar = fmax(1.0 / gam1x, kEps);
    const double lpi = log(M_PI), l1 = log(var0 + 1.0 / gam1x), l0 = log(rvar), lb = log(1.0 - beta) - log(beta);
    const double gain = var0 / (var0 + rvar);
    double sumv = 0.0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const cx<T> r = r1[k];
        double xo[2], comp[2] = {(double)r.re, (double)r.im};
#pragma unroll
        for (int q = 0; q < 2; ++q) {                       // each real-embedded component is its own scalar channel
            const double a2 = comp[q] * comp[q];
            const double ll1 = -(lpi + l1 + a2 / (var0 + 1.0 / gam1x));            // CAwgnEstimIn.m:181-184
            const double ll0 = -(lpi + l0 + a2 / rvar);                            // SparseScaEstim.m:101-103
            const double ea = fmin(fmax(ll0 - ll1 + lb, -500.0), 500.0);           // :108-110
            const double py1 = 1.0 / (1.0 + exp(ea)), py0 = 1.0 - py1;             // :111-112
            const double xh = gain * comp[q], xv = gain * rvar;                    // CAwgnEstimIn.m:100-102
            const double xx = py1 * xh;                                            // SparseScaEstim.m:161
            sumv += py1 * (xh * xh - xx * xx) + py1 * xv - py0 * xx * xx;          // :164-165
            xo[q] = xx;
        }
        if (i > 0) { const cx<T> o = x1[k]; xo[0] = p.damp * xo[0] + (1.0 - p.damp) * o.re; xo[1] = p.damp * xo[1] + (1.0 - p.damp) * o.im; }   // VampGlmEst.m:367
        x1[k] = mk<T>((T)xo[0], (T)xo[1]);
    }
    const double eta1x = 1.0 / (block_sum<T>(sumv) / N2);                          // :365
    double gam2x = eta1x - gam1x;                                                  // :369
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const cx<T> a = x1[k], r = r1[k];
        r2[k] = mk<T>((T)((a.re * eta1x - r.re * gam1x) / gam2x), (T)((a.im * eta1x - r.im * gam1x) / gam2x));   // :370 (unclipped gam2x)
    }
    gam2x = clipg(gam2x);                                                          // :379
    const double pvar = 1.0 / gam1z, gz = pvar / (pvar + wvar), zvar = wvar * gz;  // CAwgnEstimOut.m:102-108
    const double eta1z = 1.0 / zvar;                                               // mean of a constant vector (:382)
    double gam2z = eta1z - gam1z;                                                  // :383
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        const cx<T> pp = p1[k], yy = y[k];
        const double zr = gz * (yy.re - pp.re) + pp.re, zi = gz * (yy.im - pp.im) + pp.im;
        p2[k] = mk<T>((T)((zr * eta1z - pp.re * gam1z) / gam2z), (T)((zi * eta1z - pp.im * gam1z) / gam2z));   // :384
    }
    gam2z = clipg(gam2z);                                                          // :393
    if (i > 0) gam2z = p.damp * gam2z + (1.0 - p.damp) * sc[S_G2ZOLD];             // :395
    const int K = m <= n ? m : n;
    const T* d = p.d + (long long)b * p.ld_d;
    T* inv = p.inv + (size_t)b * K;
    double acc = 0.0;
    const double ratio = gam2x / gam2z;
    for (int k = threadIdx.x; k < K; k += blockDim.x) { const double iv = 1.0 / ((double)d[k] + ratio); inv[k] = (T)iv; acc += (double)d[k] * iv; }   // :400
    const double alf = block_sum<T>(acc) / (double)n - kEps;      // (2 * sum)/(2n) - eps: every eigenvalue counts twice in the embedding (:401)
    if (m > n) {                                                  // r2 * (gam2x / gam2z), the first term of :408
        const int mx = m;
        cx<T>* t3 = p.t3 + (size_t)b * mx;
        for (int k = threadIdx.x; k < n; k += blockDim.x) { const cx<T> r = r2[k]; t3[k] = mk<T>((T)(r.re * ratio), (T)(r.im * ratio)); }
    }
    if (threadIdx.x == 0) { sc[S_G2X] = gam2x; sc[S_G2Z] = gam2z; sc[S_ALF] = alf; sc[S_G2ZOLD] = gam2z; sc[S_G1XOLD] = gam1x; }
}

// dense mat-vec, one warp per output element:  out = op(Mat) (s1 .* s2 .* x)  [+ add | add - .]
//   trans = 0: out[i] = sum_j Mat[i,j] x[j]  (rows x cols);  trans = 1: out[j] = sum_i conj(Mat[i,j]) x[i]
template <typename T>
struct MvP {
    const cx<T>* Mat; long long ld_mat; int rows, cols, trans;
    const cx<T>* x; long long ld_x;
    cx<T>* out; long long ld_out;
    const cx<T>* add; long long ld_add; int sub_from;
    const T* s1; long long ld_s1; const T* s2; long long ld_s2;
};
template <typename T>
__global__ void __launch_bounds__(256) k_matvec(MvP<T> q) {
    const int b = blockIdx.y;
    const cx<T>* M = q.Mat + (long long)b * q.ld_mat;
    const cx<T>* xv = q.x + (long long)b * q.ld_x;
    const T* s1 = q.s1 ? q.s1 + (long long)b * q.ld_s1 : nullptr;
    const T* s2 = q.s2 ? q.s2 + (long long)b * q.ld_s2 : nullptr;
    if (!q.trans) {
        // 32 output rows per CTA: lanes along the rows (Mat(:, k) is read coalesced, x[k] is a broadcast), the eight warps take every eighth column and
        // their partial sums are added in warp order through shared memory - 8 x rows/32 CTAs per trial instead of one thread per row looping over all columns
        __shared__ T part[8][32][2];
        const int lane = threadIdx.x % 32, w = threadIdx.x / 32, i = blockIdx.x * 32 + lane;
        T re = 0, im = 0;
        if (i < q.rows) {
#pragma unroll 4
            for (int k = w; k < q.cols; k += 8) {
                const cx<T> a = M[i + (size_t)q.rows * k];
                cx<T> v = xv[k];
                if (s1) { T s = s1[k]; if (s2) s *= s2[k]; v = mk<T>(v.re * s, v.im * s); }
                cmac<T>(re, im, a.re, a.im, v.re, v.im);
            }
        }
        part[w][lane][0] = re; part[w][lane][1] = im;
        __syncthreads();
        if (w == 0 && i < q.rows) {
            re = 0; im = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { re += part[u][lane][0]; im += part[u][lane][1]; }
            if (q.add) { const cx<T> a = q.add[(long long)b * q.ld_add + i]; if (q.sub_from) { re = a.re - re; im = a.im - im; } else { re += a.re; im += a.im; } }
            q.out[(long long)b * q.ld_out + i] = mk<T>(re, im);
        }
        return;
    }
    // conjugate-transposed product: one warp per output, lanes along the (contiguous) column
    const int lane = threadIdx.x % 32, warp = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    if (warp >= q.cols) return;
    T re = 0, im = 0;
    for (int k = lane; k < q.rows; k += 32) {
        const cx<T> a = M[k + (size_t)q.rows * warp];
        cx<T> v = xv[k];
        if (s1) { T s = s1[k]; if (s2) s *= s2[k]; v = mk<T>(v.re * s, v.im * s); }
        cmac<T>(re, im, a.re, -a.im, v.re, v.im);
    }
    for (int s = 16; s > 0; s >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, s); im += __shfl_xor_sync(0xffffffffu, im, s); }
    if (lane == 0) {
        if (q.add) { const cx<T> a = q.add[(long long)b * q.ld_add + warp]; if (q.sub_from) { re = a.re - re; im = a.im - im; } else { re += a.re; im += a.im; } }
        q.out[(long long)b * q.ld_out + warp] = mk<T>(re, im);
    }
}

// second half (VampGlmEst.m:414-495): damping of z2, extrinsic r1 / p1, precisions
template <typename T>
__global__ void __launch_bounds__(256) k_vamp_post(VampP<T> p, const cx<T>* x2v, const cx<T>* resid, const cx<T>* udt, long long ld_t) {
    const int b = blockIdx.x, m = p.m, n = p.n, i = p.it;
    double* sc = p.sc + (size_t)b * 16;
    const double alf = sc[S_ALF], gam2x = sc[S_G2X], gam2z = sc[S_G2Z], del = (double)m / (double)n;
    cx<T>* r1 = p.r1 + (size_t)b * n; const cx<T>* r2 = p.r2 + (size_t)b * n; const cx<T>* x2 = x2v + (long long)b * ld_t;
    cx<T>* p1 = p.p1 + (size_t)b * m; const cx<T>* p2 = p.p2 + (size_t)b * m; cx<T>* z2 = p.z2 + (size_t)b * m;
    const cx<T>* rs = resid ? resid + (long long)b * ld_t : nullptr; const cx<T>* ud = udt + (long long)b * ld_t;      // p2 - A r2 ; U (d .* t)
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        r1[k] = mk<T>((T)((x2[k].re - r2[k].re * (1.0 - alf)) / alf), (T)((x2[k].im - r2[k].im * (1.0 - alf)) / alf));        // :467
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
        double zr = ud[k].re, zi = ud[k].im;                                                               // m > n: z2 = A x2 (:411)
        if (resid) { zr += (double)p2[k].re - rs[k].re; zi += (double)p2[k].im - rs[k].im; }               // z2 = A r2 + U (d .* t)  (:406)
        if (i > 0) { zr = p.damp * zr + (1.0 - p.damp) * z2[k].re; zi = p.damp * zi + (1.0 - p.damp) * z2[k].im; }            // :415
        z2[k] = mk<T>((T)zr, (T)zi);
        p1[k] = mk<T>((T)((del * zr - p2[k].re * alf) / (del - alf)), (T)((del * zi - p2[k].im * alf) / (del - alf)));        // :468
    }
    if (threadIdx.x == 0) {
        double g1x = clipg(gam2x * alf / (1.0 - alf));                                                                        // :472-481
        const double g1z = clipg(gam2z * (del - alf) / alf);                                                                  // :482-491
        if (i > 0) g1x = p.damp * g1x + (1.0 - p.damp) * sc[S_G1XOLD];                                                        // :494
        sc[S_G1X] = g1x; sc[S_G1Z] = g1z;
    }
}

template <typename T>
__global__ void k_vamp_init(VampP<T> p) {
    const int b = blockIdx.x;
    double* sc = p.sc + (size_t)b * 16;
    for (int k = threadIdx.x; k < p.n; k += blockDim.x) { p.r1[(size_t)b * p.n + k] = mk<T>(T(0), T(0)); p.x1[(size_t)b * p.n + k] = mk<T>(T(0), T(0)); }
    for (int k = threadIdx.x; k < p.m; k += blockDim.x) { p.p1[(size_t)b * p.m + k] = mk<T>(T(0), T(0)); p.z2[(size_t)b * p.m + k] = mk<T>(T(0), T(0)); }
    if (threadIdx.x == 0) { for (int k = 0; k < 16; ++k) sc[k] = 0.0; sc[S_G1X] = 1e-8; sc[S_G1Z] = 1e-8; }     // VampGlmOpt.m:25,27
}

template <typename T>
static int run_vamp(Handle* h, int mem, int m, int n, int batch, int nit, double damp, const void* y_, long long ld_y, const void* A_, long long ld_A,
                    const double* sigma_, const double* L_, const void* basis_, long long ld_b, const void* d_, long long ld_d, void* x_, long long ld_x) {
    if (m <= 0 || n <= 0 || batch <= 0 || nit < 1) return fail(h, JSTSP_E_ARG, "bad dimension");
    if (!y_ || !A_ || !sigma_ || !L_ || !basis_ || !d_ || !x_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    const int K = m <= n ? m : n, mx = m > n ? m : n;
    const size_t esz = sizeof(cx<T>);
    if (!ld_y) ld_y = m; if (!ld_x) ld_x = n;
    VampP<T> p{};
    for (int pass = 0; pass < 2; ++pass) {
        Arena ar(pass ? h->ws : nullptr, pass ? h->ws_bytes : 0);
        p = VampP<T>{};
        p.m = m; p.n = n; p.nit = nit; p.damp = damp;
        p.r1 = ar.take<cx<T>>((size_t)n * batch); p.x1 = ar.take<cx<T>>((size_t)n * batch); p.r2 = ar.take<cx<T>>((size_t)n * batch);
        p.p1 = ar.take<cx<T>>((size_t)m * batch); p.p2 = ar.take<cx<T>>((size_t)m * batch); p.z2 = ar.take<cx<T>>((size_t)m * batch);
        p.t0 = ar.take<cx<T>>((size_t)mx * batch); p.t1 = ar.take<cx<T>>((size_t)mx * batch); p.t2 = ar.take<cx<T>>((size_t)mx * batch); p.t3 = ar.take<cx<T>>((size_t)mx * batch);
        p.inv = ar.take<T>((size_t)K * batch);
        p.sc = ar.take<double>((size_t)16 * batch);
        if (host) {
            cx<T>* dA = ar.take<cx<T>>((size_t)m * n * (ld_A ? batch : 1));
            cx<T>* dy = ar.take<cx<T>>((size_t)m * batch);
            cx<T>* dB = ar.take<cx<T>>((size_t)K * K * (ld_b ? batch : 1));
            T* dd = ar.take<T>((size_t)K * (ld_d ? batch : 1));
            double* ds = ar.take<double>(batch); double* dl = ar.take<double>(batch);
            cx<T>* dx = ar.take<cx<T>>((size_t)n * batch);
            if (pass) {
                if (ld_A) JSTSP_CUDA(h, cudaMemcpy2DAsync(dA, (size_t)m * n * esz, A_, (size_t)ld_A * esz, (size_t)m * n * esz, batch, cudaMemcpyHostToDevice, st));
                else JSTSP_CUDA(h, cudaMemcpyAsync(dA, A_, (size_t)m * n * esz, cudaMemcpyHostToDevice, st));
                JSTSP_CUDA(h, cudaMemcpy2DAsync(dy, m * esz, y_, (size_t)ld_y * esz, m * esz, batch, cudaMemcpyHostToDevice, st));
                if (ld_b) JSTSP_CUDA(h, cudaMemcpy2DAsync(dB, (size_t)K * K * esz, basis_, (size_t)ld_b * esz, (size_t)K * K * esz, batch, cudaMemcpyHostToDevice, st));
                else JSTSP_CUDA(h, cudaMemcpyAsync(dB, basis_, (size_t)K * K * esz, cudaMemcpyHostToDevice, st));
                if (ld_d) JSTSP_CUDA(h, cudaMemcpy2DAsync(dd, K * sizeof(T), d_, (size_t)ld_d * sizeof(T), K * sizeof(T), batch, cudaMemcpyHostToDevice, st));
                else JSTSP_CUDA(h, cudaMemcpyAsync(dd, d_, K * sizeof(T), cudaMemcpyHostToDevice, st));
                JSTSP_CUDA(h, cudaMemcpyAsync(ds, sigma_, sizeof(double) * batch, cudaMemcpyHostToDevice, st));
                JSTSP_CUDA(h, cudaMemcpyAsync(dl, L_, sizeof(double) * batch, cudaMemcpyHostToDevice, st));
            }
            p.A = dA; p.ld_A = ld_A ? (long long)m * n : 0; p.y = dy; p.ld_y = m; p.basis = dB; p.ld_b = ld_b ? (long long)K * K : 0;
            p.d = dd; p.ld_d = ld_d ? K : 0; p.sigma = ds; p.Lnz = dl; p.x_out = dx; p.ld_x = n;
        } else {
            p.A = (const cx<T>*)A_; p.ld_A = ld_A; p.y = (const cx<T>*)y_; p.ld_y = ld_y; p.basis = (const cx<T>*)basis_; p.ld_b = ld_b;
            p.d = (const T*)d_; p.ld_d = ld_d; p.sigma = sigma_; p.Lnz = L_; p.x_out = (cx<T>*)x_; p.ld_x = ld_x;
        }
        if (!pass) { int rc = ensure_workspace(h, ar.off); if (rc) return rc; }
    }
    k_vamp_init<T><<<batch, 256, 0, st>>>(p); h->launches++;
    auto mv = [&](const cx<T>* Mat, long long ldm, int rows, int cols, int trans, const cx<T>* x, long long ldx, cx<T>* out, long long ldo,
                  const cx<T>* add, long long lda, int sub, const T* s1, long long lds1, const T* s2, long long lds2) {
        MvP<T> q{Mat, ldm, rows, cols, trans, x, ldx, out, ldo, add, lda, sub, s1, lds1, s2, lds2};
        dim3 g(trans ? (cols + 7) / 8 : (rows + 31) / 32, batch);
        JSTSP_LAUNCH(h, PK_OTHER, (k_matvec<T><<<g, 256, 0, st>>>(q)));
    };
    for (int it = 0; it < nit; ++it) {
        p.it = it;
        JSTSP_LAUNCH(h, PK_OTHER, (k_vamp_pre<T><<<batch, 256, 0, st>>>(p)));
        if (m > n) {
            // VampGlmEst.m:407-411 (M > N): basis = V (n x n) of A^H A, d its n eigenvalues
            mv(p.A, p.ld_A, m, n, 1, p.p2, m, p.t0, mx, p.t3, mx, 0, nullptr, 0, nullptr, 0);            // t0 = r2 gam2x/gam2z + A^H p2
            mv(p.basis, p.ld_b, n, n, 1, p.t0, mx, p.t1, mx, nullptr, 0, 0, nullptr, 0, nullptr, 0);     // t1 = V^H t0
            mv(p.basis, p.ld_b, n, n, 0, p.t1, mx, p.t3, mx, nullptr, 0, 0, p.inv, K, nullptr, 0);       // t3 = x2 = V (inv .* t1)
            mv(p.A, p.ld_A, m, n, 0, p.t3, mx, p.t2, mx, nullptr, 0, 0, nullptr, 0, nullptr, 0);         // t2 = z2 = A x2
            JSTSP_LAUNCH(h, PK_OTHER, (k_vamp_post<T><<<batch, 256, 0, st>>>(p, p.t3, (const cx<T>*)nullptr, p.t2, (long long)mx)));
            continue;
        }
        // VampGlmEst.m:402-406 (M <= N)
        mv(p.A, p.ld_A, m, n, 0, p.r2, n, p.t0, mx, p.p2, m, 1, nullptr, 0, nullptr, 0);             // t0 = p2 - A r2
        mv(p.basis, p.ld_b, m, m, 1, p.t0, mx, p.t1, mx, nullptr, 0, 0, nullptr, 0, nullptr, 0);     // t1 = U^H t0
        mv(p.basis, p.ld_b, m, m, 0, p.t1, mx, p.t2, mx, nullptr, 0, 0, p.inv, K, nullptr, 0);       // t2 = U (inv .* t1)
        mv(p.A, p.ld_A, m, n, 1, p.t2, mx, p.t3, mx, p.r2, n, 0, nullptr, 0, nullptr, 0);            // t3 = x2 = r2 + A^H t2
        mv(p.basis, p.ld_b, m, m, 0, p.t1, mx, p.t2, mx, nullptr, 0, 0, p.inv, K, p.d, p.ld_d);      // t2 = U (d .* inv .* t1)
        JSTSP_LAUNCH(h, PK_OTHER, (k_vamp_post<T><<<batch, 256, 0, st>>>(p, p.t3, p.t0, p.t2, (long long)mx)));
    }
    JSTSP_CUDA(h, cudaGetLastError());
    // x = x1(1:n) + 1j*x1(n+1:2n)  (vamp.m:54) == the complex x1 of the last iteration
    if (host) {
        JSTSP_CUDA(h, cudaMemcpy2DAsync(x_, (size_t)ld_x * esz, p.x1, n * esz, n * esz, batch, cudaMemcpyDeviceToHost, st));
        JSTSP_CUDA(h, cudaStreamSynchronize(st));
    } else {
        JSTSP_CUDA(h, cudaMemcpy2DAsync(x_, (size_t)ld_x * esz, p.x1, n * esz, n * esz, batch, cudaMemcpyDeviceToDevice, st));
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" int jstsp_vamp(jstsp_handle* h, int dtype, int mem, int m, int n, int batch, int nit, double damp,
                          const void* y, long long ld_y, const void* A, long long ld_A, const double* sigma, const double* L,
                          const void* basis, long long ld_basis, const void* d, long long ld_d, void* x, long long ld_x) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_vamp<float>(h, mem, m, n, batch, nit, damp, y, ld_y, A, ld_A, sigma, L, basis, ld_basis, d, ld_d, x, ld_x);
    if (dtype == JSTSP_F64) return run_vamp<double>(h, mem, m, n, batch, nit, damp, y, ld_y, A, ld_A, sigma, L, basis, ld_basis, d, ld_d, x, ld_x);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
