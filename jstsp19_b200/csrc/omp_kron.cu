// omp_kron.cu - OMP over the beamspace-delay Kronecker dictionary, never materialised.
//
//   jstsp_omp_kron(A, B, Y, m) == OMP(kron(B.', A), vec(Y), m)       (benchmark_algorithms/OMP.m:1-32)
//
// The reference's drivers hand the sparse solvers Phi = kron(B.', A), y = vec(Y)
// (plot_errorVSdelays.m:77-78, plot_errorVSframelength.m:78-79); at BASELINE config 2
// (Nt = Nr = 64, 4x oversampled grids) Phi would be 8192 x 262144 complex - 32 GiB in fp64 - so the
// factors are the input here.  Column j = g + G*p (0-based, column-major over the G x P unknown) of
// Phi is vec(A(:,g) * B(p,:)), hence
//     Phi' * r   = vec(A^H R B^H)            (OMP.m:17; R = reshape(r, N, M))
//     atom j     = A(:,g) B(p,:)             (OMP.m:18)
// and everything else (growing least squares, residual, scatter) is OMP.m unchanged.
//
// Per iteration, two launches:
//   k_kron_corr   grid (P/32, batch): T_c = R B_c^H (N x 32), C_c = A^H T_c (G x 32) out of shared
//                 memory, |C|^2 arg-max with the first-maximum rule and the runner-up -> one
//                 candidate per (chunk, trial).  C is never written.  FP32 FMA bound.
//   k_kron_update grid (batch): final arg-max, duplicate handling, modified Gram-Schmidt of the
//                 rank-one atom against the stored directions (fp64 accumulation), residual update.
// k_kron_finish back-substitutes once and scatters x_hat (OMP.m:27-31).
#include "common.cuh"

namespace jstsp {

constexpr int KC_PC = 32;      // delay-beam columns (p) per CTA
constexpr int KC_RB = 64;      // row block (n in phase 1, g in phase 2)
constexpr int KC_KT = 32;      // contraction tile

template <typename T>
struct KronP {
    int N, M, G, P, m, t, nchunk;
    const cx<T>* A;  long long ld_A;     // N x G
    const cx<T>* AH; long long ld_AH;    // G x N (conjugate transpose, built once per call)
    const cx<T>* B;  long long ld_B;     // P x M
    const cx<T>* Y;  long long ld_Y;     // N x M
    cx<T>* res;                          // [b][N*M] residual
    cx<T>* Q;                            // [b][m][N*M] orthonormal directions
    cx<T>* Rt;                           // [b][m*m] upper triangular
    cx<T>* z;                            // [b][m]
    cx<T>* scratch;                      // [b][N*M] candidate atom when it does not fit in shared memory
    int* state;                          // [b][2 + 3m]: nuniq, amb, sel[m], uniq_idx[m], mult[m]
    double* cand_val;                    // [b][nchunk][2] best, second
    int* cand_idx;                       // [b][nchunk]
    int atom_in_smem;
    double margin_tol;
    // outputs
    cx<T>* x_hat; long long ld_x;
    int* index_set;
    cx<T>* x_sel;
    cx<T>* r_out; long long ld_r;
    int* ambiguous;
};

// rows owned by a thread inside a 64-row block: {2q, 2q+1, 32+2q, 32+2q+1} - two 16-byte (fp32)
// shared-memory reads at a 16-byte stride across the quarter-warp, i.e. conflict-free.
__device__ __forceinline__ int kc_row(int q, int i) { return (i >> 1) * 32 + 2 * q + (i & 1); }

template <typename T>
__global__ void __launch_bounds__(256) k_kron_setup(KronP<T> p, int nA) {
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t NM = (size_t)p.N * p.M;
    const cx<T>* Y = p.Y + (long long)b * p.ld_Y;
    cx<T>* r = p.res + (size_t)b * NM;
    for (size_t i = tid; i < NM; i += 256) r[i] = Y[i];                       // r = v (OMP.m:10)
    int* st = p.state + (size_t)b * (2 + 3 * p.m);
    for (int i = tid; i < 2 + 3 * p.m; i += 256) st[i] = 0;
    if (b < nA) {                                                             // AH(g,n) = conj(A(n,g))
        const cx<T>* A = p.A + (long long)b * p.ld_A;
        cx<T>* AH = const_cast<cx<T>*>(p.AH) + (long long)b * p.ld_AH;
        for (int i = tid; i < p.N * p.G; i += 256) { const int g = i % p.G, n = i / p.G; AH[i] = conj(A[n + (size_t)p.N * g]); }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_kron_corr(KronP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int chunk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, M = p.M, G = p.G, P = p.P;
    const int Npad = ceil_div(N, KC_KT) * KC_KT;
    cx<T>* Tc = reinterpret_cast<cx<T>*>(smem);                   // [Npad][PC]
    cx<T>* Lt = Tc + (size_t)Npad * KC_PC;                        // [KT][RB]  left operand tile
    cx<T>* Bt = Lt + KC_KT * KC_RB;                               // [KT][PC]
    __shared__ double s_best[8], s_second[8];
    const cx<T>* R = p.res + (size_t)b * N * M;
    const cx<T>* B = p.B + (long long)b * p.ld_B;
    const cx<T>* AH = p.AH + (long long)b * p.ld_AH;
    const int p0 = chunk * KC_PC;
    const int q = tid & 15, tp = tid >> 4;                         // rows kc_row(q,0..3), columns 2tp, 2tp+1
    const cx<T> zero = mk<T>(T(0), T(0));

    // ---- phase 1: T_c(n, p) = sum_m R(n, m) conj(B(p0+p, m)) ----
    for (int n0 = 0; n0 < Npad; n0 += KC_RB) {
        T ar[4][2], ai[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ar[i][0] = ar[i][1] = ai[i][0] = ai[i][1] = T(0); }
        for (int m0 = 0; m0 < M; m0 += KC_KT) {
            __syncthreads();
            for (int e = tid; e < KC_KT * KC_RB; e += 256) {
                const int k = e / KC_RB, i = e % KC_RB, n = n0 + i, mm = m0 + k;
                Lt[e] = (n < N && mm < M) ? R[n + (size_t)N * mm] : zero;
            }
            for (int e = tid; e < KC_KT * KC_PC; e += 256) {
                const int k = e / KC_PC, j = e % KC_PC, pp = p0 + j, mm = m0 + k;
                Bt[e] = (pp < P && mm < M) ? B[pp + (size_t)P * mm] : zero;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < KC_KT; ++k) {
                cx<T> l[4], r2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) l[i] = Lt[k * KC_RB + kc_row(q, i)];
                r2[0] = Bt[k * KC_PC + 2 * tp]; r2[1] = Bt[k * KC_PC + 2 * tp + 1];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) cmac<T>(ar[i][j], ai[i][j], l[i].re, l[i].im, r2[j].re, -r2[j].im);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int n = n0 + kc_row(q, i);
            if (n < Npad) { Tc[(size_t)n * KC_PC + 2 * tp] = mk<T>(ar[i][0], ai[i][0]); Tc[(size_t)n * KC_PC + 2 * tp + 1] = mk<T>(ar[i][1], ai[i][1]); }
        }
    }
    // ---- phase 2: C_c(g, p) = sum_n AH(g, n) T_c(n, p); arg-max of |C|^2, lowest j = g + G p on ties ----
    double best = -1.0, second = -1.0; long long bidx = 0x7fffffffffffLL;
    for (int g0 = 0; g0 < G; g0 += KC_RB) {
        T ar[4][2], ai[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { ar[i][0] = ar[i][1] = ai[i][0] = ai[i][1] = T(0); }
        for (int n0 = 0; n0 < Npad; n0 += KC_KT) {
            __syncthreads();
            for (int e = tid; e < KC_KT * KC_RB; e += 256) {
                const int k = e / KC_RB, i = e % KC_RB, g = g0 + i, n = n0 + k;
                Lt[e] = (g < G && n < N) ? AH[g + (size_t)G * n] : zero;
            }
            __syncthreads();
#pragma unroll 8
            for (int k = 0; k < KC_KT; ++k) {
                cx<T> l[4], r2[2];
#pragma unroll
                for (int i = 0; i < 4; ++i) l[i] = Lt[k * KC_RB + kc_row(q, i)];
                r2[0] = Tc[(size_t)(n0 + k) * KC_PC + 2 * tp]; r2[1] = Tc[(size_t)(n0 + k) * KC_PC + 2 * tp + 1];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 2; ++j) cmac<T>(ar[i][j], ai[i][j], l[i].re, l[i].im, r2[j].re, r2[j].im);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int g = g0 + kc_row(q, i), pp = p0 + 2 * tp + j;
                if (g < G && pp < P) {
                    const double mag = (double)ar[i][j] * ar[i][j] + (double)ai[i][j] * ai[i][j];
                    const long long jj = g + (long long)G * pp;
                    if (mag > best || (mag == best && jj < bidx)) { second = best; best = mag; bidx = jj; }
                    else if (mag > second) second = mag;
                }
            }
    }
    // ---- reduce (best, second, index) over the CTA ----
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, second, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob > best || (ob == best && oi < bidx)) { second = best > os ? best : os; best = ob; bidx = oi; }
        else { const double c = ob > os ? ob : os; if (c > second) second = c; }
    }
    __shared__ long long s_lidx[8];
    if (lane == 0) { s_best[warp] = best; s_second[warp] = second; s_lidx[warp] = bidx; }
    __syncthreads();
    if (tid == 0) {
        double gb = -1.0, gs = -1.0; long long gi = 0x7fffffffffffLL;
        for (int w = 0; w < 8; ++w) {
            const double wb = s_best[w], wsd = s_second[w];
            if (wb > gb || (wb == gb && s_lidx[w] < gi)) { if (gb > gs) gs = gb; gb = wb; gi = s_lidx[w]; }
            else if (wb > gs) gs = wb;
            if (wsd > gs) gs = wsd;
        }
        p.cand_val[((size_t)b * p.nchunk + chunk) * 2 + 0] = gb;
        p.cand_val[((size_t)b * p.nchunk + chunk) * 2 + 1] = gs;
        p.cand_idx[(size_t)b * p.nchunk + chunk] = (int)gi;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_kron_update(KronP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, G = p.G, P = p.P, m = p.m, t = p.t;
    const size_t NM = (size_t)N * p.M;
    __shared__ double s_red[8][2];
    __shared__ int s_pick, s_dup;
    cx<T>* qn = p.atom_in_smem ? reinterpret_cast<cx<T>*>(smem) : p.scratch + (size_t)b * NM;
    cx<T>* r = p.res + (size_t)b * NM;
    cx<T>* Q = p.Q + (size_t)b * m * NM;
    cx<T>* Rt = p.Rt + (size_t)b * m * m;
    cx<T>* z = p.z + (size_t)b * m;
    int* st = p.state + (size_t)b * (2 + 3 * m);
    int* sel = st + 2; int* uniq_idx = sel + m; int* mult = uniq_idx + m;
    const int nu = st[0];
    if (tid == 0) {
        double gb = -1.0, gs = -1.0; int gi = 0x7fffffff;
        for (int c = 0; c < p.nchunk; ++c) {                                  // chunks ascend in p, i.e. in j
            const double wb = p.cand_val[((size_t)b * p.nchunk + c) * 2], wsd = p.cand_val[((size_t)b * p.nchunk + c) * 2 + 1];
            const int wi = p.cand_idx[(size_t)b * p.nchunk + c];
            if (wb > gb || (wb == gb && wi < gi)) { if (gb > gs) gs = gb; gb = wb; gi = wi; }
            else if (wb > gs) gs = wb;
            if (wsd > gs) gs = wsd;
        }
        if (gs >= 0.0 && gb - gs <= p.margin_tol * gb) st[1]++;
        p.index_set[(size_t)b * m + t] = gi + 1;                              // 1-based (OMP.m:17)
        int dup = -1;
        for (int k = 0; k < nu; ++k) if (uniq_idx[k] == gi) dup = k;
        if (dup >= 0) { sel[t] = dup; mult[dup]++; }
        s_pick = gi; s_dup = dup;
    }
    __syncthreads();
    if (s_dup >= 0) return;                                                   // no new direction (see omp.cu)
    const int pick = s_pick, g = pick % G, pp = pick / G;
    const cx<T>* A = p.A + (long long)b * p.ld_A + (size_t)N * g;
    const cx<T>* B = p.B + (long long)b * p.ld_B + pp;
    // atom = A(:,g) B(p,:)   (OMP.m:18 on the Kronecker column)
    double an = 0.0;
    for (size_t i = tid; i < NM; i += 256) {
        const int n = (int)(i % N), mm = (int)(i / N);
        const cx<T> a = A[n] * B[(size_t)P * mm];
        qn[i] = a; an += (double)a.re * a.re + (double)a.im * a.im;
    }
    for (int o = 16; o > 0; o >>= 1) an += __shfl_xor_sync(0xffffffffu, an, o);
    if (lane == 0) s_red[warp][0] = an;
    __syncthreads();
    double a2 = 0.0;
    for (int w = 0; w < 8; ++w) a2 += s_red[w][0];
    __syncthreads();
    for (int k = 0; k < nu; ++k) {
        const cx<T>* qk = Q + (size_t)k * NM;
        double re = 0.0, im = 0.0;
        for (size_t i = tid; i < NM; i += 256) { const cx<T> a = qk[i], x = qn[i]; re += (double)a.re * x.re + (double)a.im * x.im; im += (double)a.re * x.im - (double)a.im * x.re; }
        for (int o = 16; o > 0; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
        if (lane == 0) { s_red[warp][0] = re; s_red[warp][1] = im; }
        __syncthreads();
        double cr = 0.0, ci = 0.0;
        for (int w = 0; w < 8; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; }
        if (tid == 0) Rt[k + (size_t)m * nu] = mk<T>((T)cr, (T)ci);
        for (size_t i = tid; i < NM; i += 256) { const cx<T> a = qk[i]; qn[i] = mk<T>(qn[i].re - (T)(cr * a.re - ci * a.im), qn[i].im - (T)(cr * a.im + ci * a.re)); }
        __syncthreads();
    }
    double nn = 0.0;
    for (size_t i = tid; i < NM; i += 256) nn += (double)qn[i].re * qn[i].re + (double)qn[i].im * qn[i].im;
    for (int o = 16; o > 0; o >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, o);
    if (lane == 0) s_red[warp][0] = nn;
    __syncthreads();
    double n2 = 0.0;
    for (int w = 0; w < 8; ++w) n2 += s_red[w][0];
    const double dep_tol = sizeof(T) == 4 ? 1e-10 : 1e-26;
    const double nrm = n2 > dep_tol * a2 ? sqrt(n2) : 0.0;
    const T inv = nrm > 0.0 ? (T)(1.0 / nrm) : T(0);
    __syncthreads();
    double zr = 0.0, zi = 0.0;
    cx<T>* qs = Q + (size_t)nu * NM;
    for (size_t i = tid; i < NM; i += 256) {
        const cx<T> qv = mk<T>(qn[i].re * inv, qn[i].im * inv);
        qs[i] = qv; qn[i] = qv;
        const cx<T> x = r[i];
        zr += (double)qv.re * x.re + (double)qv.im * x.im; zi += (double)qv.re * x.im - (double)qv.im * x.re;
    }
    for (int o = 16; o > 0; o >>= 1) { zr += __shfl_xor_sync(0xffffffffu, zr, o); zi += __shfl_xor_sync(0xffffffffu, zi, o); }
    if (lane == 0) { s_red[warp][0] = zr; s_red[warp][1] = zi; }
    __syncthreads();
    double cr = 0.0, ci = 0.0;
    for (int w = 0; w < 8; ++w) { cr += s_red[w][0]; ci += s_red[w][1]; }
    for (size_t i = tid; i < NM; i += 256) { const cx<T> qv = qn[i]; r[i] = mk<T>(r[i].re - (T)(cr * qv.re - ci * qv.im), r[i].im - (T)(cr * qv.im + ci * qv.re)); }
    if (tid == 0) {
        Rt[nu + (size_t)m * nu] = mk<T>((T)nrm, T(0));
        z[nu] = mk<T>((T)cr, (T)ci);
        sel[t] = nu; uniq_idx[nu] = pick; mult[nu] = 1;
        st[0] = nu + 1;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_kron_finish(KronP<T> p) {
    const int b = blockIdx.x, tid = threadIdx.x, m = p.m;
    const size_t NM = (size_t)p.N * p.M;
    cx<T>* Rt = p.Rt + (size_t)b * m * m;
    cx<T>* z = p.z + (size_t)b * m;
    int* st = p.state + (size_t)b * (2 + 3 * m);
    int* sel = st + 2; int* uniq_idx = sel + m; int* mult = uniq_idx + m;
    const int nu = st[0];
    if (tid == 0) {                                                           // R x = z over the unique directions
        for (int k = nu - 1; k >= 0; --k) {
            double sr = z[k].re, si = z[k].im;
            for (int j = k + 1; j < nu; ++j) {
                const cx<T> rr = Rt[k + (size_t)m * j], xj = z[j];
                sr -= (double)rr.re * xj.re - (double)rr.im * xj.im; si -= (double)rr.re * xj.im + (double)rr.im * xj.re;
            }
            const double d = Rt[k + (size_t)m * k].re;
            z[k] = d != 0.0 ? mk<T>((T)(sr / d), (T)(si / d)) : mk<T>(T(0), T(0));
        }
        if (p.ambiguous) p.ambiguous[b] = st[1];
    }
    __syncthreads();
    if (p.x_sel) for (int k = tid; k < m; k += 256) { const int u = sel[k]; const T s = T(1) / (T)mult[u]; p.x_sel[(size_t)b * m + k] = mk<T>(z[u].re * s, z[u].im * s); }
    if (p.r_out) { cx<T>* ro = p.r_out + (long long)b * p.ld_r; const cx<T>* r = p.res + (size_t)b * NM; for (size_t i = tid; i < NM; i += 256) ro[i] = r[i]; }
    if (p.x_hat) {
        cx<T>* xh = p.x_hat + (long long)b * p.ld_x;
        const size_t D = (size_t)p.G * p.P;
        for (size_t j = tid; j < D; j += 256) xh[j] = mk<T>(T(0), T(0));       // OMP.m:27
        __syncthreads();
        for (int k = tid; k < nu; k += 256) { const T s = T(1) / (T)mult[k]; xh[uniq_idx[k]] = mk<T>(z[k].re * s, z[k].im * s); }   // OMP.m:29-31, pinv's even split over duplicates
    }
}

template <typename T>
static int run_omp_kron(Handle* h, int mem, int N, int M, int G, int P, int m, int batch,
                        const void* A_, long long ld_A, const void* B_, long long ld_B, const void* Y_, long long ld_Y,
                        void* x_, long long ld_x, int* idx_, void* xsel_, void* r_, long long ld_r, int* amb_, double margin_tol) {
    if (N <= 0 || M <= 0 || G <= 0 || P <= 0 || m <= 0 || batch <= 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if ((long long)G * P > 0x7fffffffLL) return fail(h, JSTSP_E_ARG, "G*P exceeds the int32 index range of index_set");
    if (!A_ || !B_ || !Y_ || !idx_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    const size_t esz = sizeof(cx<T>);
    const size_t NM = (size_t)N * M, NG = (size_t)N * G, PM = (size_t)P * M, GP = (size_t)G * P;
    if (ld_Y == 0) ld_Y = (long long)NM;
    if (ld_x == 0) ld_x = (long long)GP;
    if (ld_r == 0) ld_r = (long long)NM;
    const int Npad = ceil_div(N, KC_KT) * KC_KT;
    const size_t smem_corr = esz * ((size_t)Npad * KC_PC + KC_KT * KC_RB + KC_KT * KC_PC);
    int rc = set_smem(h, k_kron_corr<T>, smem_corr);
    if (rc) return rc;
    const int atom_in_smem = NM * esz <= 64 * 1024 ? 1 : 0;
    const size_t smem_upd = atom_in_smem ? NM * esz : 0;
    rc = set_smem(h, k_kron_update<T>, smem_upd);
    if (rc) return rc;
    const int nchunk = ceil_div(P, KC_PC);
    int chunk = batch;
    if (h->max_chunk > 0 && chunk > h->max_chunk) chunk = h->max_chunk;
    const bool sharedA = ld_A == 0, sharedB = ld_B == 0;
    auto layout = [&](Arena& a, int nb, KronP<T>& q) {
        q.res = a.take<cx<T>>(NM * nb);
        q.Q = a.take<cx<T>>(NM * nb * m);
        q.Rt = a.take<cx<T>>((size_t)nb * m * m);
        q.z = a.take<cx<T>>((size_t)nb * m);
        q.scratch = atom_in_smem ? nullptr : a.take<cx<T>>(NM * nb);
        q.state = a.take<int>((size_t)nb * (2 + 3 * m));
        q.cand_val = a.take<double>((size_t)nb * nchunk * 2);
        q.cand_idx = a.take<int>((size_t)nb * nchunk);
        q.AH = a.take<cx<T>>(sharedA ? NG : NG * nb);
        if (host) {
            q.A = a.take<cx<T>>(sharedA ? NG : NG * nb);
            q.B = a.take<cx<T>>(sharedB ? PM : PM * nb);
            q.Y = a.take<cx<T>>(NM * nb);
            q.index_set = a.take<int>((size_t)m * nb);
            if (x_) q.x_hat = a.take<cx<T>>(GP * nb);
            if (xsel_) q.x_sel = a.take<cx<T>>((size_t)m * nb);
            if (r_) q.r_out = a.take<cx<T>>(NM * nb);
            if (amb_) q.ambiguous = a.take<int>(nb);
        }
    };
    size_t freeb = 0, totalb = 0;
    JSTSP_CUDA(h, cudaMemGetInfo(&freeb, &totalb));
    const size_t budget = (size_t)((freeb + h->ws_bytes) * 0.7);
    for (;;) {
        Arena probe(nullptr, 0); KronP<T> q{}; layout(probe, chunk, q);
        if (probe.off <= budget || chunk == 1) { rc = ensure_workspace(h, probe.off); if (rc) return rc; break; }
        chunk = (chunk + 1) / 2;
    }
    for (int b0 = 0; b0 < batch; b0 += chunk) {
        const int nb = (batch - b0) < chunk ? (batch - b0) : chunk;
        Arena ar(h->ws, h->ws_bytes);
        KronP<T> q{};
        q.N = N; q.M = M; q.G = G; q.P = P; q.m = m; q.nchunk = nchunk; q.margin_tol = margin_tol; q.atom_in_smem = atom_in_smem;
        layout(ar, nb, q);
        q.ld_AH = sharedA ? 0 : (long long)NG;
        if (host) {
            auto up = [&](const cx<T>* dst, const void* src, size_t per, long long ld, bool shared) -> cudaError_t {
                if (shared) return cudaMemcpyAsync(const_cast<cx<T>*>(dst), src, per * esz, cudaMemcpyHostToDevice, st);
                return cudaMemcpy2DAsync(const_cast<cx<T>*>(dst), per * esz, (const char*)src + (size_t)b0 * ld * esz, (size_t)ld * esz, per * esz, nb, cudaMemcpyHostToDevice, st);
            };
            JSTSP_CUDA(h, up(q.A, A_, NG, ld_A, sharedA));
            JSTSP_CUDA(h, up(q.B, B_, PM, ld_B, sharedB));
            JSTSP_CUDA(h, up(q.Y, Y_, NM, ld_Y, false));
            q.ld_A = sharedA ? 0 : (long long)NG; q.ld_B = sharedB ? 0 : (long long)PM; q.ld_Y = (long long)NM;
            q.ld_x = (long long)GP; q.ld_r = (long long)NM;
        } else {
            q.A = (const cx<T>*)A_ + (long long)b0 * ld_A; q.ld_A = ld_A;
            q.B = (const cx<T>*)B_ + (long long)b0 * ld_B; q.ld_B = ld_B;
            q.Y = (const cx<T>*)Y_ + (long long)b0 * ld_Y; q.ld_Y = ld_Y;
            q.x_hat = x_ ? (cx<T>*)x_ + (long long)b0 * ld_x : nullptr; q.ld_x = ld_x;
            q.index_set = idx_ + (size_t)b0 * m;
            q.x_sel = xsel_ ? (cx<T>*)xsel_ + (size_t)b0 * m : nullptr;
            q.r_out = r_ ? (cx<T>*)r_ + (long long)b0 * ld_r : nullptr; q.ld_r = ld_r;
            q.ambiguous = amb_ ? amb_ + b0 : nullptr;
        }
        JSTSP_LAUNCH(h, PK_SETUP, (k_kron_setup<T><<<nb, 256, 0, st>>>(q, sharedA ? 1 : nb)));
        for (int t = 0; t < m; ++t) {
            q.t = t;
            JSTSP_LAUNCH(h, PK_OMP_CORR, (k_kron_corr<T><<<dim3(nchunk, nb), 256, smem_corr, st>>>(q)));
            JSTSP_LAUNCH(h, PK_OMP, (k_kron_update<T><<<nb, 256, smem_upd, st>>>(q)));
        }
        JSTSP_LAUNCH(h, PK_OMP, (k_kron_finish<T><<<nb, 256, 0, st>>>(q)));
        JSTSP_CUDA(h, cudaGetLastError());
        if (host) {
            JSTSP_CUDA(h, cudaMemcpyAsync(idx_ + (size_t)b0 * m, q.index_set, sizeof(int) * (size_t)m * nb, cudaMemcpyDeviceToHost, st));
            if (x_) JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)x_ + (size_t)b0 * ld_x * esz, (size_t)ld_x * esz, q.x_hat, GP * esz, GP * esz, nb, cudaMemcpyDeviceToHost, st));
            if (xsel_) JSTSP_CUDA(h, cudaMemcpyAsync((char*)xsel_ + (size_t)b0 * m * esz, q.x_sel, esz * (size_t)m * nb, cudaMemcpyDeviceToHost, st));
            if (r_) JSTSP_CUDA(h, cudaMemcpy2DAsync((char*)r_ + (size_t)b0 * ld_r * esz, (size_t)ld_r * esz, q.r_out, NM * esz, NM * esz, nb, cudaMemcpyDeviceToHost, st));
            if (amb_) JSTSP_CUDA(h, cudaMemcpyAsync(amb_ + b0, q.ambiguous, sizeof(int) * nb, cudaMemcpyDeviceToHost, st));
            JSTSP_CUDA(h, cudaStreamSynchronize(st));
        }
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" int jstsp_omp_kron(jstsp_handle* h, int dtype, int mem, int N, int M, int G, int P, int m, int batch,
                              const void* A, long long ld_A, const void* B, long long ld_B, const void* Y, long long ld_Y,
                              void* x_hat, long long ld_x, int* index_set, void* x_sel, void* residual, long long ld_r,
                              int* ambiguous, double margin_tol) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_omp_kron<float>(h, mem, N, M, G, P, m, batch, A, ld_A, B, ld_B, Y, ld_Y, x_hat, ld_x, index_set, x_sel, residual, ld_r, ambiguous, margin_tol);
    if (dtype == JSTSP_F64) return run_omp_kron<double>(h, mem, N, M, G, P, m, batch, A, ld_A, B, ld_B, Y, ld_Y, x_hat, ld_x, index_set, x_sel, residual, ld_r, ambiguous, margin_tol);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
