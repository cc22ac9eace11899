// somp.cu - simultaneous (joint / MMV) OMP, batched: one CTA per trial.
//
//   jstsp_somp replaces the reference's external call
//       spx.pursuit.joint.OrthogonalMatchingPursuit(A, K).solve(Y).Z
//   (plot_errorVSsnr.m:116-118, plot_errorVSdelays.m:114-115, plot_time_comparisions.m:101-102, ...),
//   where Y = Y_hbf * pinv(B) is N x S and A is the N x D receive dictionary.
//
// sparse-plex is not vendored by the reference and no version is pinned (README.md:9), so this is a
// documented replacement, not a parity target: the textbook row-l2 SOMP
//     t = 1..K:  d_t = first argmax_d || A(:,d)' R ||_2 ;  Z(support,:) = A(:,support) \ Y ;  R = Y - A(:,support) Z
// stopped early when the support reaches min(N, D) atoms, when the residual falls to res_tol*||Y||_F, or
// when an atom is picked twice (no new direction) - the situations in which the drivers' K = 100 on a
// 32-column dictionary (plot_errorVSsnr.m:20,116) would make the least-squares step degenerate.
// The growing least squares is an incremental modified Gram-Schmidt QR with fp64 accumulation, as in omp.cu.
#include "common.cuh"

namespace jstsp {

template <typename T>
struct SompP {
    int N, D, S, K, kmax;
    const cx<T>* A; long long ld_A;
    const cx<T>* Y; long long ld_Y;
    cx<T>* Z; long long ld_Z;            // D x S
    int* support;                        // K per trial, 1-based, 0 = unused
    int* n_iters;                        // per trial (may be null)
    cx<T>* r_out; long long ld_R;        // N x S (may be null)
    cx<T>* res;                          // [b][N*S]
    cx<T>* Q;                            // [b][N*kmax]
    cx<T>* Rt;                           // [b][kmax*kmax]
    cx<T>* W;                            // [b][kmax*S]   rows q_k' Y
    double* corr;                        // [b][D]
    double res_tol;
};

template <typename T>
__global__ void __launch_bounds__(256) k_somp(SompP<T> p) {
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, D = p.D, S = p.S, kmax = p.kmax;
    const cx<T>* A = p.A + (long long)b * p.ld_A;
    const cx<T>* Y = p.Y + (long long)b * p.ld_Y;
    cx<T>* R = p.res + (size_t)b * N * S;
    cx<T>* Q = p.Q + (size_t)b * N * kmax;
    cx<T>* Rt = p.Rt + (size_t)b * kmax * kmax;
    cx<T>* W = p.W + (size_t)b * kmax * S;
    double* corr = p.corr + (size_t)b * D;
    int* sup = p.support + (size_t)b * p.K;
    __shared__ double s_red[8];
    __shared__ double s_y2, s_nrm;
    __shared__ int s_pick, s_stop;
    __shared__ cx<double> s_c;
    extern __shared__ __align__(16) unsigned char smem[];
    cx<T>* qn = reinterpret_cast<cx<T>*>(smem);                                // N
    double y2 = 0.0;
    for (int i = tid; i < N * S; i += 256) { const cx<T> y = Y[i]; R[i] = y; y2 += (double)y.re * y.re + (double)y.im * y.im; }
    for (int i = tid; i < p.K; i += 256) sup[i] = 0;
    for (int o = 16; o > 0; o >>= 1) y2 += __shfl_xor_sync(0xffffffffu, y2, o);
    if (lane == 0) s_red[warp] = y2;
    __syncthreads();
    if (tid == 0) { double a = 0; for (int w = 0; w < 8; ++w) a += s_red[w]; s_y2 = a; s_stop = 0; }
    __syncthreads();
    int nu = 0;
    for (int t = 0; t < p.K && nu < kmax; ++t) {
        // ---- row-l2 correlation: corr[d] = sum_s |A(:,d)' R(:,s)|^2, one warp per atom, lanes along s ----
        for (int d = warp; d < D; d += 8) {
            const cx<T>* a = A + (size_t)N * d;
            double acc = 0.0;
            for (int s = lane; s < S; s += 32) {
                T re = 0, im = 0;
                const cx<T>* r = R + (size_t)N * s;
                for (int n = 0; n < N; ++n) { const cx<T> x = a[n], y = r[n]; cmac<T>(re, im, x.re, -x.im, y.re, y.im); }
                acc += (double)re * re + (double)im * im;
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) corr[d] = acc;
        }
        __syncthreads();
        if (tid == 0) {
            double gb = -1.0; int gi = 0;
            for (int d = 0; d < D; ++d) if (corr[d] > gb) { gb = corr[d]; gi = d; }      // first maximum
            int dup = 0;
            for (int k = 0; k < nu; ++k) if (sup[k] == gi + 1) dup = 1;
            s_pick = gi; if (dup) s_stop = 1;
        }
        __syncthreads();
        if (s_stop) break;
        const int pick = s_pick;
        // ---- modified Gram-Schmidt of A(:,pick) against the nu stored directions ----
        for (int i = tid; i < N; i += 256) qn[i] = A[(size_t)N * pick + i];
        __syncthreads();
        for (int k = 0; k < nu; ++k) {
            if (warp == 0) {
                double re = 0.0, im = 0.0;
                for (int i = lane; i < N; i += 32) { const cx<T> a = Q[(size_t)N * k + i], x = qn[i]; re += (double)a.re * x.re + (double)a.im * x.im; im += (double)a.re * x.im - (double)a.im * x.re; }
                for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
                if (lane == 0) { s_c.re = re; s_c.im = im; Rt[k + (size_t)kmax * nu] = mk<T>((T)re, (T)im); }
            }
            __syncthreads();
            const double cr = s_c.re, ci = s_c.im;
            for (int i = tid; i < N; i += 256) { const cx<T> a = Q[(size_t)N * k + i]; qn[i] = mk<T>(qn[i].re - (T)(cr * a.re - ci * a.im), qn[i].im - (T)(cr * a.im + ci * a.re)); }
            __syncthreads();
        }
        if (warp == 0) {
            double nn = 0.0;
            for (int i = lane; i < N; i += 32) nn += (double)qn[i].re * qn[i].re + (double)qn[i].im * qn[i].im;
            for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
            if (lane == 0) s_nrm = sqrt(nn);
        }
        __syncthreads();
        const double nrm = s_nrm;
        if (!(nrm > 0.0)) break;                                                         // uniform: s_nrm is shared
        const T inv = (T)(1.0 / nrm);
        for (int i = tid; i < N; i += 256) { const cx<T> qv = mk<T>(qn[i].re * inv, qn[i].im * inv); qn[i] = qv; Q[(size_t)N * nu + i] = qv; }
        __syncthreads();
        // ---- W(nu, s) = q' R(:,s) (== q' Y(:,s)); R(:,s) -= q W(nu, s); residual norm ----
        double r2 = 0.0;
        for (int s = tid; s < S; s += 256) {
            cx<T>* r = R + (size_t)N * s;
            double wr = 0.0, wi = 0.0;
            for (int n = 0; n < N; ++n) { const cx<T> qv = qn[n], x = r[n]; wr += (double)qv.re * x.re + (double)qv.im * x.im; wi += (double)qv.re * x.im - (double)qv.im * x.re; }
            W[nu + (size_t)kmax * s] = mk<T>((T)wr, (T)wi);
            for (int n = 0; n < N; ++n) {
                const cx<T> qv = qn[n];
                const cx<T> x = mk<T>(r[n].re - (T)(wr * qv.re - wi * qv.im), r[n].im - (T)(wr * qv.im + wi * qv.re));
                r[n] = x; r2 += (double)x.re * x.re + (double)x.im * x.im;
            }
        }
        for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        if (lane == 0) s_red[warp] = r2;
        __syncthreads();
        if (tid == 0) {
            double a = 0; for (int w = 0; w < 8; ++w) a += s_red[w];
            Rt[nu + (size_t)kmax * nu] = mk<T>((T)nrm, T(0));
            sup[nu] = pick + 1;
            if (a <= p.res_tol * p.res_tol * s_y2) s_stop = 1;
        }
        ++nu;
        __syncthreads();
        if (s_stop) break;
    }
    __syncthreads();
    // ---- Z(support,:) = Rt \ W, zeros elsewhere ----
    cx<T>* Z = p.Z + (long long)b * p.ld_Z;
    for (int i = tid; i < D * S; i += 256) Z[i] = mk<T>(T(0), T(0));
    __syncthreads();
    for (int s = tid; s < S; s += 256) {
        cx<T>* w = W + (size_t)kmax * s;
        for (int k = nu - 1; k >= 0; --k) {
            double sr = w[k].re, si = w[k].im;
            for (int j = k + 1; j < nu; ++j) {
                const cx<T> rr = Rt[k + (size_t)kmax * j], xj = w[j];
                sr -= (double)rr.re * xj.re - (double)rr.im * xj.im; si -= (double)rr.re * xj.im + (double)rr.im * xj.re;
            }
            const double d = Rt[k + (size_t)kmax * k].re;
            w[k] = mk<T>((T)(sr / d), (T)(si / d));
            Z[(sup[k] - 1) + (size_t)D * s] = w[k];
        }
    }
    if (p.n_iters && tid == 0) p.n_iters[b] = nu;
    if (p.r_out) { cx<T>* ro = p.r_out + (long long)b * p.ld_R; for (int i = tid; i < N * S; i += 256) ro[i] = R[i]; }
}

template <typename T>
static int run_somp(Handle* h, int mem, int N, int D, int S, int K, int batch, const void* A_, long long ld_A, const void* Y_, long long ld_Y,
                    void* Z_, long long ld_Z, int* sup_, int* nit_, void* r_, long long ld_R, double res_tol) {
    if (N <= 0 || D <= 0 || S <= 0 || K <= 0 || batch <= 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if (!A_ || !Y_ || !Z_ || !sup_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    const size_t esz = sizeof(cx<T>);
    const size_t ND = (size_t)N * D, NS = (size_t)N * S, DS = (size_t)D * S;
    if (ld_Y == 0) ld_Y = (long long)NS;
    if (ld_Z == 0) ld_Z = (long long)DS;
    if (ld_R == 0) ld_R = (long long)NS;
    int kmax = K; if (kmax > N) kmax = N; if (kmax > D) kmax = D;
    const size_t smem = esz * (size_t)N;
    int rc = set_smem(h, k_somp<T>, smem);
    if (rc) return rc;
    int chunk = batch;
    if (h->max_chunk > 0 && chunk > h->max_chunk) chunk = h->max_chunk;
    const bool sharedA = ld_A == 0;
    auto layout = [&](Arena& a, int nb, SompP<T>& q) {
        q.res = a.take<cx<T>>(NS * nb);
        q.Q = a.take<cx<T>>((size_t)N * kmax * nb);
        q.Rt = a.take<cx<T>>((size_t)kmax * kmax * nb);
        q.W = a.take<cx<T>>((size_t)kmax * S * nb);
        q.corr = a.take<double>((size_t)D * nb);
        if (host) {
            q.A = a.take<cx<T>>(sharedA ? ND : ND * nb);
            q.Y = a.take<cx<T>>(NS * nb);
            q.Z = a.take<cx<T>>(DS * nb);
            q.support = a.take<int>((size_t)K * nb);
            if (nit_) q.n_iters = a.take<int>(nb);
            if (r_) q.r_out = a.take<cx<T>>(NS * nb);
        }
    };
    size_t freeb = 0, totalb = 0;
    JSTSP_CUDA(h, cudaMemGetInfo(&freeb, &totalb));
    const size_t budget = (size_t)((freeb + h->ws_bytes) * 0.7);
    for (;;) {
        Arena probe(nullptr, 0); SompP<T> q{}; layout(probe, chunk, q);
        if (probe.off <= budget || chunk == 1) { rc = ensure_workspace(h, probe.off); if (rc) return rc; break; }
        chunk = (chunk + 1) / 2;
    }
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = (batch - b0) < chunk ? (batch - b0) : chunk;
        Arena ar(h->ws, h->ws_bytes);
        SompP<T> q{};
        q.N = N; q.D = D; q.S = S; q.K = K; q.kmax = kmax; q.res_tol = res_tol;
        layout(ar, nb, q);
        if (host) {
            if (sharedA) JSTSP_CUDA(h, cudaMemcpyAsync(const_cast<cx<T>*>(q.A), A_, ND * esz, cudaMemcpyHostToDevice, st));
            else JSTSP_CUDA(h, cudaMemcpy2DAsync(const_cast<cx<T>*>(q.A), ND * esz, (const char*)A_ + (size_t)b0 * ld_A * esz, (size_t)ld_A * esz, ND * esz, nb, cudaMemcpyHostToDevice, st));
            JSTSP_CUDA(h, cudaMemcpy2DAsync(const_cast<cx<T>*>(q.Y), NS * esz, (const char*)Y_ + (size_t)b0 * ld_Y * esz, (size_t)ld_Y * esz, NS * esz, nb, cudaMemcpyHostToDevice, st));
            q.ld_A = sharedA ? 0 : (long long)ND; q.ld_Y = (long long)NS; q.ld_Z = (long long)DS; q.ld_R = (long long)NS;
        } else {
            q.A = (const cx<T>*)A_ + (long long)b0 * ld_A; q.ld_A = ld_A;
            q.Y = (const cx<T>*)Y_ + (long long)b0 * ld_Y; q.ld_Y = ld_Y;
            q.Z = (cx<T>*)Z_ + (long long)b0 * ld_Z; q.ld_Z = ld_Z;
            q.support = sup_ + (size_t)b0 * K;
            q.n_iters = nit_ ? nit_ + b0 : nullptr;
            q.r_out = r_ ? (cx<T>*)r_ + (long long)b0 * ld_R : nullptr; q.ld_R = ld_R;
        }
        JSTSP_LAUNCH(h, PK_SOMP, (k_somp<T><<<nb, 256, smem, st>>>(q)));
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)Z_ + (size_t)b0 * ld_Z * esz, (size_t)ld_Z * esz, q.Z, DS * esz, DS * esz, nb, cudaMemcpyDeviceToHost, st));
            JSTSP_CUDA(h, cudaMemcpyAsync(sup_ + (size_t)b0 * K, q.support, sizeof(int) * (size_t)K * nb, cudaMemcpyDeviceToHost, st));
            if (nit_) JSTSP_CUDA(h, cudaMemcpyAsync(nit_ + b0, q.n_iters, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
            if (r_) JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)r_ + (size_t)b0 * ld_R * esz, (size_t)ld_R * esz, q.r_out, NS * esz, NS * esz, nb, cudaMemcpyDeviceToHost, st));
            JSTSP_CUDA(h, cudaStreamSynchronize(st));
        }
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" int jstsp_somp(jstsp_handle* h, int dtype, int mem, int N, int D, int S, int K, int batch,
                          const void* A, long long ld_A, const void* Y, long long ld_Y,
                          void* Z, long long ld_Z, int* support, int* n_iters, void* residual, long long ld_R, double res_tol) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_somp<float>(h, mem, N, D, S, K, batch, A, ld_A, Y, ld_Y, Z, ld_Z, support, n_iters, residual, ld_R, res_tol);
    if (dtype == JSTSP_F64) return run_somp<double>(h, mem, N, D, S, K, batch, A, ld_A, Y, ld_Y, Z, ld_Z, support, n_iters, residual, ld_R, res_tol);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
