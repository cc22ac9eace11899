// omp.cu - orthogonal matching pursuit, batched: one CTA per trial.
//
//   jstsp_omp replaces benchmark_algorithms/OMP.m:1-32
//     t = 1..m:  idx_t = first argmax_j |A(:,j)' r|            (OMP.m:17)
//                T = [T, A(:,idx_t)] ; x = pinv(T) v ; r = v - T x   (OMP.m:18-21)
//     x_hat(idx_t) = x(t), later duplicates overwrite           (OMP.m:27-31)
//
// The growing pinv (an SVD per iteration in the reference) is replaced by an incremental
// modified-Gram-Schmidt QR kept per trial: r <- r - q (q' r) is the same orthogonal projection,
// and x is recovered once at the end by back-substitution.  An index picked twice (only happens
// once r is at rounding level; the reference has no stopping rule) adds no new direction: pinv's
// minimum-norm solution then splits the coefficient evenly over the duplicate columns, which is
// reproduced explicitly.  The correlation A' r streams the dictionary once per iteration
// (coalesced, one warp per column) - this solver is HBM/L2-bandwidth bound.
//
// Tie handling: the arg-max compares |c_j|^2 with lowest-index-wins; every iteration also tracks
// the runner-up, and the number of iterations whose relative margin is below `margin_tol` is
// returned per trial so callers can tell when the support might differ from an fp64 evaluation.
#include "common.cuh"

namespace jstsp {

template <typename T>
struct OmpP {
    int measures, size_d, m;
    const cx<T>* A; long long ld_A;
    const cx<T>* v; long long ld_v;
    cx<T>* x_hat; long long ld_x;
    int* index_set; long long ld_idx;          // m per trial, 1-based
    cx<T>* target; long long ld_t;             // measures x m (may be null)
    int* ambiguous;                            // per trial (may be null)
    cx<T>* Q;                                  // workspace [b][measures*m]
    cx<T>* R;                                  // workspace [b][m*m]   (upper triangular, column-major)
    double margin_tol;
};

template <typename T>
__global__ void __launch_bounds__(256) k_omp(OmpP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid % 32, warp = tid / 32;
    const int Mm = p.measures, D = p.size_d, m = p.m;
    cx<T>* r = reinterpret_cast<cx<T>*>(smem);                 // residual (Mm)
    cx<T>* qn = r + Mm;                                        // candidate column (Mm)
    cx<T>* z = qn + Mm;                                        // Q' v coefficients (m)
    int* sel = reinterpret_cast<int*>(z + m);                  // unique slot of pick t (m)
    int* uniq_idx = sel + m;                                   // column index of unique slot (m)
    int* mult = uniq_idx + m;                                  // multiplicity of unique slot (m)
    __shared__ double s_best[8], s_second[8];
    __shared__ int s_bidx[8];
    __shared__ double s_red[8][2];
    __shared__ int s_pick, s_nuniq, s_amb;
    const cx<T>* A = p.A + (long long)b * p.ld_A;
    const cx<T>* v = p.v + (long long)b * p.ld_v;
    cx<T>* Q = p.Q + (size_t)b * Mm * m;
    cx<T>* R = p.R + (size_t)b * m * m;
    for (int i = tid; i < Mm; i += 256) r[i] = v[i];           // r = v (OMP.m:10)
    if (tid == 0) { s_nuniq = 0; s_amb = 0; }
    __syncthreads();
    for (int t = 0; t < m; ++t) {
        // ---- correlation + first arg-max: one warp per column, lanes along the measurements ----
        double best = -1.0, second = -1.0; int bidx = 0x7fffffff;
        for (int j = warp; j < D; j += 8) {
            const cx<T>* col = A + (size_t)j * Mm;
            T re = 0, im = 0;
            for (int i = lane; i < Mm; i += 32) { cx<T> a = col[i], x = r[i]; cmac<T>(re, im, a.re, -a.im, x.re, x.im); }
            for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
            const double mag = (double)re * re + (double)im * im;
            if (mag > best) { second = best; best = mag; bidx = j; }       // j ascending per warp: '>' keeps the first maximum
            else if (mag > second) second = mag;
        }
        if (lane == 0) { s_best[warp] = best; s_second[warp] = second; s_bidx[warp] = bidx; }
        __syncthreads();
        if (tid == 0) {
            double gb = -1.0, gs = -1.0; int gi = 0x7fffffff;
            for (int w = 0; w < 8; ++w) {
                const double wb = s_best[w];
                if (wb > gb || (wb == gb && s_bidx[w] < gi)) { if (gb > gs) gs = gb; gb = wb; gi = s_bidx[w]; }
                else if (wb > gs) gs = wb;
                if (s_second[w] > gs) gs = s_second[w];
            }
            if (gs >= 0.0 && gb - gs <= p.margin_tol * gb) s_amb++;
            s_pick = gi;
            p.index_set[(long long)b * p.ld_idx + t] = gi + 1;             // 1-based like MATLAB (OMP.m:17)
        }
        __syncthreads();
        const int pick = s_pick;
        if (p.target) { cx<T>* tg = p.target + (long long)b * p.ld_t + (size_t)t * Mm; for (int i = tid; i < Mm; i += 256) tg[i] = A[(size_t)pick * Mm + i]; }
        // ---- duplicate pick: no new direction ----
        int dup = -1;
        const int nu = s_nuniq;
        for (int k = 0; k < nu; ++k) if (uniq_idx[k] == pick) dup = k;
        if (dup >= 0) {
            if (tid == 0) { sel[t] = dup; mult[dup]++; }
            __syncthreads();
            continue;
        }
        // ---- modified Gram-Schmidt against the nu stored directions ----
        double an = 0.0;
        for (int i = tid; i < Mm; i += 256) { cx<T> a = A[(size_t)pick * Mm + i]; qn[i] = a; an += (double)a.re * a.re + (double)a.im * a.im; }
        for (int o = 16; o > 0; o >>= 1) an += __shfl_xor_sync(0xffffffffu, an, o);
        if (lane == 0) s_red[warp][0] = an;
        __syncthreads();
        double a2 = 0.0;
        for (int w = 0; w < 8; ++w) a2 += s_red[w][0];
        __syncthreads();
        for (int k = 0; k < nu; ++k) {
            const cx<T>* qk = Q + (size_t)k * Mm;
            double re = 0.0, im = 0.0;
            for (int i = tid; i < Mm; i += 256) { cx<T> a = qk[i], x = qn[i]; re += (double)a.re * x.re + (double)a.im * x.im; im += (double)a.re * x.im - (double)a.im * x.re; }
            for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
            if (lane == 0) { s_red[warp][0] = re; s_red[warp][1] = im; }
            __syncthreads();
            double cr = 0.0, ci = 0.0;
            for (int w = 0; w < 8; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; }
            if (tid == 0) R[k + (size_t)m * nu] = mk<T>((T)cr, (T)ci);
            for (int i = tid; i < Mm; i += 256) { cx<T> a = qk[i]; qn[i] = mk<T>(qn[i].re - (T)(cr * a.re - ci * a.im), qn[i].im - (T)(cr * a.im + ci * a.re)); }
            __syncthreads();
        }
        double nn = 0.0;
        for (int i = tid; i < Mm; i += 256) nn += (double)qn[i].re * qn[i].re + (double)qn[i].im * qn[i].im;
        for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
        if (lane == 0) s_red[warp][0] = nn;
        __syncthreads();
        double n2 = 0.0;
        for (int w = 0; w < 8; ++w) n2 += s_red[w][0];
        // a column that is numerically inside the span of the chosen ones carries no new direction
        const double dep_tol = sizeof(T) == 4 ? 1e-10 : 1e-26;
        const double nrm = n2 > dep_tol * a2 ? sqrt(n2) : 0.0;
        const T inv = nrm > 0.0 ? (T)(1.0 / nrm) : T(0);
        __syncthreads();
        // q = qn / ||qn|| ; z_nu = q' r (== q' v) ; r -= q z_nu
        double zr = 0.0, zi = 0.0;
        cx<T>* qs = Q + (size_t)nu * Mm;
        for (int i = tid; i < Mm; i += 256) {
            cx<T> q = mk<T>(qn[i].re * inv, qn[i].im * inv);
            qs[i] = q; qn[i] = q;
            cx<T> x = r[i];
            zr += (double)q.re * x.re + (double)q.im * x.im; zi += (double)q.re * x.im - (double)q.im * x.re;
        }
        for (int o = 16; o > 0; o >>= 1) { zr += __shfl_xor_sync(0xffffffffu, zr, o); zi += __shfl_xor_sync(0xffffffffu, zi, o); }
        if (lane == 0) { s_red[warp][0] = zr; s_red[warp][1] = zi; }
        __syncthreads();
        double cr = 0.0, ci = 0.0;
        for (int w = 0; w < 8; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; }
        for (int i = tid; i < Mm; i += 256) { cx<T> q = qn[i]; r[i] = mk<T>(r[i].re - (T)(cr * q.re - ci * q.im), r[i].im - (T)(cr * q.im + ci * q.re)); }
        if (tid == 0) {
            R[nu + (size_t)m * nu] = mk<T>((T)nrm, T(0));
            z[nu] = mk<T>((T)cr, (T)ci);
            sel[t] = nu; uniq_idx[nu] = pick; mult[nu] = 1;
            s_nuniq = nu + 1;
        }
        __syncthreads();
    }
    // ---- back-substitution R x = z over the unique directions (thread 0; m is small) ----
    if (tid == 0) {
        const int nu = s_nuniq;
        for (int k = nu - 1; k >= 0; --k) {
            double sr = z[k].re, si = z[k].im;
            for (int j = k + 1; j < nu; ++j) {
                cx<T> rr = R[k + (size_t)m * j], xj = z[j];
                sr -= (double)rr.re * xj.re - (double)rr.im * xj.im; si -= (double)rr.re * xj.im + (double)rr.im * xj.re;
            }
            const double d = R[k + (size_t)m * k].re;
            z[k] = d != 0.0 ? mk<T>((T)(sr / d), (T)(si / d)) : mk<T>(T(0), T(0));
        }
        if (p.ambiguous) p.ambiguous[b] = s_amb;
    }
    __syncthreads();
    cx<T>* xh = p.x_hat + (long long)b * p.ld_x;
    for (int j = tid; j < D; j += 256) xh[j] = mk<T>(T(0), T(0));            // x_hat = zeros (OMP.m:27)
    __syncthreads();
    for (int k = tid; k < s_nuniq; k += 256) {                               // OMP.m:29-31 with pinv's even split over duplicates
        const T s = T(1) / (T)mult[k];
        xh[uniq_idx[k]] = mk<T>(z[k].re * s, z[k].im * s);
    }
}

template <typename T>
static int run_omp(Handle* h, int mem, int measures, int size_d, int m, int batch, const void* A_, long long ld_A, const void* v_, long long ld_v,
                   void* x_, long long ld_x, int* idx_, void* target_, int* amb_, double margin_tol) {
    if (measures <= 0 || size_d <= 0 || m <= 0 || batch <= 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if (!A_ || !v_ || !x_ || !idx_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    const size_t esz = sizeof(cx<T>);
    if (ld_v == 0) ld_v = measures;
    if (ld_x == 0) ld_x = size_d;
    const size_t smem = esz * (2 * (size_t)measures + m) + sizeof(int) * 3 * (size_t)m + 16;
    int rc = set_smem(h, k_omp<T>, smem);
    if (rc) return rc;
    int chunk = batch;
    if (h->max_chunk > 0 && chunk > h->max_chunk) chunk = h->max_chunk;
    const size_t AD = (size_t)measures * size_d;
    auto layout = [&](Arena& a, int nb, OmpP<T>& q) {
        q.Q = a.take<cx<T>>((size_t)nb * measures * m);
        q.R = a.take<cx<T>>((size_t)nb * m * m);
        if (host) {
            q.A = a.take<cx<T>>(ld_A ? AD * nb : AD);
            q.v = a.take<cx<T>>((size_t)measures * nb);
            q.x_hat = a.take<cx<T>>((size_t)size_d * nb);
            q.index_set = a.take<int>((size_t)m * nb);
            if (target_) q.target = a.take<cx<T>>((size_t)measures * m * nb);
            if (amb_) q.ambiguous = a.take<int>(nb);
        }
    };
    size_t freeb = 0, totalb = 0;
    JSTSP_CUDA(h, cudaMemGetInfo(&freeb, &totalb));
    const size_t budget = (size_t)((freeb + h->ws_bytes) * 0.7);
    for (;;) {
        Arena probe(nullptr, 0); OmpP<T> q{}; layout(probe, chunk, q);
        if (probe.off <= budget || chunk == 1) { rc = ensure_workspace(h, probe.off); if (rc) return rc; break; }
        chunk = (chunk + 1) / 2;
    }
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = (batch - b0) < chunk ? (batch - b0) : chunk;
        Arena ar(h->ws, h->ws_bytes);
        OmpP<T> q{};
        q.measures = measures; q.size_d = size_d; q.m = m; q.margin_tol = margin_tol;
        layout(ar, nb, q);
        if (host) {
            if (ld_A == 0) JSTSP_CUDA(h, cudaMemcpyAsync(const_cast<cx<T>*>(q.A), A_, AD * esz, cudaMemcpyHostToDevice, st));
            else JSTSP_CUDA(h, cudaMemcpy2DAsync(const_cast<cx<T>*>(q.A), AD * esz, (const char*)A_ + (size_t)b0 * ld_A * esz, (size_t)ld_A * esz, AD * esz, nb, cudaMemcpyHostToDevice, st));
            JSTSP_CUDA(h, cudaMemcpy2DAsync(const_cast<cx<T>*>(q.v), measures * esz, (const char*)v_ + (size_t)b0 * ld_v * esz, (size_t)ld_v * esz, measures * esz, nb, cudaMemcpyHostToDevice, st));
            q.ld_A = ld_A ? (long long)AD : 0; q.ld_v = measures; q.ld_x = size_d; q.ld_idx = m; q.ld_t = (long long)measures * m;
        } else {
            q.A = (const cx<T>*)A_ + (long long)b0 * ld_A; q.ld_A = ld_A;
            q.v = (const cx<T>*)v_ + (long long)b0 * ld_v; q.ld_v = ld_v;
            q.x_hat = (cx<T>*)x_ + (long long)b0 * ld_x; q.ld_x = ld_x;
            q.index_set = idx_ + (size_t)b0 * m; q.ld_idx = m;
            q.target = target_ ? (cx<T>*)target_ + (size_t)b0 * measures * m : nullptr; q.ld_t = (long long)measures * m;
            q.ambiguous = amb_ ? amb_ + b0 : nullptr;
        }
        JSTSP_LAUNCH(h, PK_OMP, (k_omp<T><<<nb, 256, smem, st>>>(q)));
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)x_ + (size_t)b0 * ld_x * esz, (size_t)ld_x * esz, q.x_hat, size_d * esz, size_d * esz, nb, cudaMemcpyDeviceToHost, st));
            JSTSP_CUDA(h, cudaMemcpyAsync(idx_ + (size_t)b0 * m, q.index_set, sizeof(int) * (size_t)m * nb, cudaMemcpyDeviceToHost, st));
            if (target_) JSTSP_CUDA(h, cudaMemcpyAsync((char*)target_ + (size_t)b0 * measures * m * esz, q.target, esz * (size_t)measures * m * nb, cudaMemcpyDeviceToHost, st));
            if (amb_) JSTSP_CUDA(h, cudaMemcpyAsync(amb_ + b0, q.ambiguous, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
            JSTSP_CUDA(h, cudaStreamSynchronize(st));
        }
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" int jstsp_omp(jstsp_handle* h, int dtype, int mem, int measures, int size_d, int m, int batch,
                         const void* A, long long ld_A, const void* v, long long ld_v,
                         void* x_hat, long long ld_x, int* index_set, void* target_matrix, int* ambiguous, double margin_tol) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_omp<float>(h, mem, measures, size_d, m, batch, A, ld_A, v, ld_v, x_hat, ld_x, index_set, target_matrix, ambiguous, margin_tol);
    if (dtype == JSTSP_F64) return run_omp<double>(h, mem, measures, size_d, m, batch, A, ld_A, v, ld_v, x_hat, ld_x, index_set, target_matrix, ambiguous, margin_tol);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
