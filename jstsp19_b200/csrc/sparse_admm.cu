// sparse_admm.cu - l1-ADMM in beamspace, batched: one CTA per trial.
//
//   jstsp_sparse_admm replaces benchmark_algorithms/sparse_admm.m:1-36
//     s = soft(R + Z/rho, tau_s/rho)                                         (:21-22)
//     r = (A'A - rho I) \ (vec Z - rho s + A' vec OH),  A = kron(conj(Dt), Dr) (:15-16,26)
//     Z += rho (R - S)                                                        (:30)
//     conv(i) = norm(Dr S Dt' - Htrue)^2 / norm(Htrue)^2   (spectral norms)   (:32)
//   with the hard-coded rho = 0.01, tau_s = 1e-4 (:12-13) and the "- rho I" sign of :16.
//
// The dense (Mr Mt)^2 solve is done in the eigenbases of Dr'Dr and Dt'Dt:
//   A vec(S) = vec(Dr S Dt^H),  A'A vec(S) = vec(Dr'Dr S (Dt'Dt)^T),
//   R = Qr [ (Qr^H RHS conj(Qt)) ./ (lr_i lt_j - rho) ] Qt^T.
// Like the reference (reshape(s,Mr,Mt), Dr*S*Dt') this needs Gr == Mr and Gt == Mt.
// Matrices are at most 64 x 64, so one CTA owns a trial; state lives in a per-trial global
// workspace (L2 resident), the two eigen-decompositions and the spectral norms use the block
// Jacobi solver.
#include "common.cuh"
#include "jacobi.cuh"

namespace jstsp {

constexpr double kSpRho = 0.01;      // sparse_admm.m:12
constexpr double kSpTau = 0.0001;    // sparse_admm.m:13

template <typename T>
struct SpP {
    int Mr, Mt, imax;
    const cx<T>* Htrue; long long ld_H;
    const cx<T>* OH; long long ld_OH;
    const cx<T>* Dr; long long ld_Dr;
    const cx<T>* Dt; long long ld_Dt;
    cx<T>* S; long long ld_S;
    T* conv; long long ld_conv;
    cx<T>* ws; size_t ws_per_trial;          // complex scratch per trial
};

// C(m x n) = opA(A) * opB(B); op: 0 = as is, 1 = conjugate transpose, 2 = conjugate, 3 = transpose.
// A is (m x k) after op, B is (k x n) after op; lda/ldb are the leading dimensions of the stored arrays.
template <typename T>
__device__ void mm(cx<T>* C, const cx<T>* A, int opA, int lda, const cx<T>* B, int opB, int ldb, int m, int n, int k) {
    for (int t = threadIdx.x; t < m * n; t += blockDim.x) {
        const int i = t % m, j = t / m;
        T re = 0, im = 0;
        for (int q = 0; q < k; ++q) {
            cx<T> a = (opA == 0 || opA == 2) ? A[i + (size_t)lda * q] : A[q + (size_t)lda * i];
            if (opA == 1 || opA == 2) a.im = -a.im;
            cx<T> b = (opB == 0 || opB == 2) ? B[q + (size_t)ldb * j] : B[j + (size_t)ldb * q];
            if (opB == 1 || opB == 2) b.im = -b.im;
            cmac<T>(re, im, a.re, a.im, b.re, b.im);
        }
        C[t] = mk<T>(re, im);
    }
    __syncthreads();
}

// sigma_max^2 of E (m x n) through the Gram matrix on the smaller side
template <typename T>
__device__ double smax2(JacobiSmem& sm, const cx<T>* E, int m, int n) {
    const int g = m < n ? m : n;
    for (int t = threadIdx.x; t < g * g; t += blockDim.x) {
        const int i = t % g, j = t / g;
        double re = 0.0, im = 0.0;
        if (m <= n) for (int q = 0; q < n; ++q) { cx<T> a = E[i + (size_t)m * q], b = E[j + (size_t)m * q]; re += (double)a.re * b.re + (double)a.im * b.im; im += (double)a.im * b.re - (double)a.re * b.im; }
        else for (int q = 0; q < m; ++q) { cx<T> a = E[q + (size_t)m * i], b = E[q + (size_t)m * j]; re += (double)a.re * b.re + (double)a.im * b.im; im += (double)a.re * b.im - (double)a.im * b.re; }
        sm.Are[t] = re; sm.Aim[t] = im;
    }
    __syncthreads();
    jacobi_hermitian_block(sm, g);
    double mx = 0.0;
    for (int k = 0; k < g; ++k) mx = fmax(mx, sm.Are[k + g * k]);
    __syncthreads();
    return mx;
}

template <typename T>
__global__ void __launch_bounds__(256) k_sparse_admm(SpP<T> p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.x, Mr = p.Mr, Mt = p.Mt, MM = Mr * Mt;
    const int nmax = Mr > Mt ? Mr : Mt;
    JacobiSmem sm; sm.carve(smem, nmax);
    double* lr = reinterpret_cast<double*>(smem + JacobiSmem::bytes(nmax));
    double* lt = lr + Mr;
    const cx<T>* OH = p.OH + (long long)b * p.ld_OH;
    const cx<T>* Dr = p.Dr + (long long)b * p.ld_Dr;
    const cx<T>* Dt = p.Dt + (long long)b * p.ld_Dt;
    const cx<T>* Ht = p.Htrue ? p.Htrue + (long long)b * p.ld_H : nullptr;
    cx<T>* w = p.ws + (size_t)b * p.ws_per_trial;
    cx<T>* Z = w; cx<T>* R = Z + MM; cx<T>* S = R + MM; cx<T>* rhs = S + MM; cx<T>* tmp = rhs + MM; cx<T>* AhOH = tmp + MM;
    cx<T>* Qr = AhOH + MM; cx<T>* Qt = Qr + Mr * Mr; cx<T>* G = Qt + Mt * Mt;   // G: max(Mr,Mt)^2 scratch
    const T rho = (T)kSpRho, thr = (T)(kSpTau / kSpRho);
    // eigen-decompositions of Dr'Dr and Dt'Dt
    for (int side = 0; side < 2; ++side) {
        const int n = side == 0 ? Mr : Mt;
        const cx<T>* D = side == 0 ? Dr : Dt;
        for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
            const int i = t % n, j = t / n;
            double re = 0.0, im = 0.0;
            for (int q = 0; q < n; ++q) { cx<T> a = D[q + (size_t)n * i], c = D[q + (size_t)n * j]; re += (double)a.re * c.re + (double)a.im * c.im; im += (double)a.re * c.im - (double)a.im * c.re; }
            sm.Are[t] = re; sm.Aim[t] = im;
        }
        __syncthreads();
        jacobi_hermitian_block(sm, n);
        cx<T>* Q = side == 0 ? Qr : Qt; double* l = side == 0 ? lr : lt;
        for (int t = threadIdx.x; t < n * n; t += blockDim.x) Q[t] = mk<T>((T)sm.Ure[t], (T)sm.Uim[t]);
        for (int t = threadIdx.x; t < n; t += blockDim.x) l[t] = sm.Are[t + n * t];
        __syncthreads();
    }
    // A' vec(OH) = vec(Dr^H OH Dt)
    mm<T>(tmp, Dr, 1, Mr, OH, 0, Mr, Mr, Mt, Mr);
    mm<T>(AhOH, tmp, 0, Mr, Dt, 0, Mt, Mr, Mt, Mt);
    for (int t = threadIdx.x; t < MM; t += blockDim.x) { Z[t] = mk<T>(T(0), T(0)); R[t] = Z[t]; S[t] = Z[t]; }
    double hn = 1.0;
    if (p.conv && Ht) hn = smax2<T>(sm, Ht, Mr, Mt);
    __syncthreads();
    for (int it = 0; it < p.imax; ++it) {
        for (int t = threadIdx.x; t < MM; t += blockDim.x) {
            const T vr = R[t].re + Z[t].re / rho, vi = R[t].im + Z[t].im / rho;                 // :21
            const cx<T> s = mk<T>(soft1<T>(vr, thr), soft1<T>(vi, thr));                         // :22
            S[t] = s;
            rhs[t] = mk<T>(Z[t].re - rho * s.re + AhOH[t].re, Z[t].im - rho * s.im + AhOH[t].im);   // :26 right-hand side
        }
        __syncthreads();
        mm<T>(tmp, Qr, 1, Mr, rhs, 0, Mr, Mr, Mt, Mr);           // Qr^H RHS
        mm<T>(rhs, tmp, 0, Mr, Qt, 2, Mt, Mr, Mt, Mt);           // ... conj(Qt)
        for (int t = threadIdx.x; t < MM; t += blockDim.x) {
            const T d = (T)(lr[t % Mr] * lt[t / Mr] - kSpRho);    // eigenvalues of A'A - rho I   (:16)
            rhs[t] = mk<T>(rhs[t].re / d, rhs[t].im / d);
        }
        __syncthreads();
        mm<T>(tmp, Qr, 0, Mr, rhs, 0, Mr, Mr, Mt, Mr);
        mm<T>(R, tmp, 0, Mr, Qt, 3, Mt, Mr, Mt, Mt);             // ... Qt^T
        for (int t = threadIdx.x; t < MM; t += blockDim.x) Z[t] = mk<T>(Z[t].re + rho * (R[t].re - S[t].re), Z[t].im + rho * (R[t].im - S[t].im));   // :30
        __syncthreads();
        if (p.conv && Ht) {
            mm<T>(tmp, Dr, 0, Mr, S, 0, Mr, Mr, Mt, Mr);
            mm<T>(G, tmp, 0, Mr, Dt, 1, Mt, Mr, Mt, Mt);         // Dr S Dt'
            for (int t = threadIdx.x; t < MM; t += blockDim.x) G[t] = mk<T>(G[t].re - Ht[t].re, G[t].im - Ht[t].im);
            __syncthreads();
            const double e = smax2<T>(sm, G, Mr, Mt);
            if (threadIdx.x == 0) p.conv[(long long)b * p.ld_conv + it] = (T)(e / hn);          // :32
        }
    }
    cx<T>* out = p.S + (long long)b * p.ld_S;
    for (int t = threadIdx.x; t < MM; t += blockDim.x) out[t] = S[t];
}

template <typename T>
static int run_sparse_admm(Handle* h, int mem, int Mr, int Mt, int batch, int imax, const void* Ht_, long long ld_H, const void* OH_, long long ld_OH,
                           const void* Dr_, long long ld_Dr, const void* Dt_, long long ld_Dt, void* S_, long long ld_S, void* conv_, long long ld_conv) {
    if (Mr <= 0 || Mt <= 0 || batch <= 0 || imax < 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if (!OH_ || !Dr_ || !Dt_ || !S_) return fail(h, JSTSP_E_ARG, "NULL buffer");
    if (conv_ && !Ht_) return fail(h, JSTSP_E_ARG, "convergence_error needs Htrue");
    if (Mr > 64 || Mt > 64) return fail(h, JSTSP_E_UNSUPPORTED, "sparse_admm kernel covers Mr, Mt <= 64");
    const bool host = mem == JSTSP_HOST;
    cudaStream_t st = h->stream;
    const size_t esz = sizeof(cx<T>), MM = (size_t)Mr * Mt;
    const int nmax = Mr > Mt ? Mr : Mt;
    if (ld_S == 0) ld_S = (long long)MM;
    if (ld_conv == 0) ld_conv = imax;
    const size_t per = 6 * MM + (size_t)Mr * Mr + (size_t)Mt * Mt + (size_t)nmax * nmax + 8;
    const size_t smem = JacobiSmem::bytes(nmax) + sizeof(double) * (Mr + Mt) + 16;
    int rc = set_smem(h, k_sparse_admm<T>, smem);
    if (rc) return rc;
    Arena probe(nullptr, 0);
    auto layout = [&](Arena& a, SpP<T>& q) {
        q.ws = a.take<cx<T>>(per * batch); q.ws_per_trial = per;
        if (host) {
            q.OH = a.take<cx<T>>(MM * batch);
            q.Dr = a.take<cx<T>>((size_t)Mr * Mr * (ld_Dr ? batch : 1));
            q.Dt = a.take<cx<T>>((size_t)Mt * Mt * (ld_Dt ? batch : 1));
            if (Ht_) q.Htrue = a.take<cx<T>>(MM * batch);
            q.S = a.take<cx<T>>(MM * batch);
            if (conv_) q.conv = a.take<T>((size_t)imax * batch);
        }
    };
    SpP<T> q{};
    layout(probe, q);
    rc = ensure_workspace(h, probe.off);
    if (rc) return rc;
    Arena ar(h->ws, h->ws_bytes);
    q = SpP<T>{};
    q.Mr = Mr; q.Mt = Mt; q.imax = imax;
    layout(ar, q);
    if (host) {
        auto up = [&](const void* dst, const void* src, size_t per_, long long ld) -> cudaError_t {
            if (ld == 0) return cudaMemcpyAsync(const_cast<void*>(dst), src, per_ * esz, cudaMemcpyHostToDevice, st);
            return cudaMemcpy2DAsync(const_cast<void*>(dst), per_ * esz, src, (size_t)ld * esz, per_ * esz, batch, cudaMemcpyHostToDevice, st);
        };
        JSTSP_CUDA(h, up(q.OH, OH_, MM, ld_OH ? ld_OH : (long long)MM));
        JSTSP_CUDA(h, up(q.Dr, Dr_, (size_t)Mr * Mr, ld_Dr));
        JSTSP_CUDA(h, up(q.Dt, Dt_, (size_t)Mt * Mt, ld_Dt));
        if (Ht_) JSTSP_CUDA(h, up(q.Htrue, Ht_, MM, ld_H ? ld_H : (long long)MM));
        q.ld_OH = MM; q.ld_H = MM; q.ld_Dr = ld_Dr ? (long long)Mr * Mr : 0; q.ld_Dt = ld_Dt ? (long long)Mt * Mt : 0; q.ld_S = MM; q.ld_conv = imax;
    } else {
        q.OH = (const cx<T>*)OH_; q.ld_OH = ld_OH ? ld_OH : (long long)MM;
        q.Dr = (const cx<T>*)Dr_; q.ld_Dr = ld_Dr; q.Dt = (const cx<T>*)Dt_; q.ld_Dt = ld_Dt;
        q.Htrue = (const cx<T>*)Ht_; q.ld_H = ld_H ? ld_H : (long long)MM;
        q.S = (cx<T>*)S_; q.ld_S = ld_S; q.conv = (T*)conv_; q.ld_conv = ld_conv;
    }
    JSTSP_LAUNCH(h, PK_OTHER, (k_sparse_admm<T><<<batch, 256, smem, st>>>(q)));
    JSTSP_CUDA(h, cudaGetLastError());
    if (host) {
        JSTSP_CUDA(h, cudaMemcpy2DAsync(S_, (size_t)ld_S * esz, q.S, MM * esz, MM * esz, batch, cudaMemcpyDeviceToHost, st));
        if (conv_) JSTSP_CUDA(h, cudaMemcpy2DAsync(conv_, (size_t)ld_conv * sizeof(T), q.conv, imax * sizeof(T), imax * sizeof(T), batch, cudaMemcpyDeviceToHost, st));
        JSTSP_CUDA(h, cudaStreamSynchronize(st));
    }
    return JSTSP_OK;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" int jstsp_sparse_admm(jstsp_handle* h, int dtype, int mem, int Mr, int Mt, int batch, int imax,
                                 const void* Htrue, long long ld_H, const void* OH, long long ld_OH,
                                 const void* Dr, long long ld_Dr, const void* Dt, long long ld_Dt,
                                 void* S, long long ld_S, void* conv, long long ld_conv) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    if (dtype == JSTSP_F32) return run_sparse_admm<float>(h, mem, Mr, Mt, batch, imax, Htrue, ld_H, OH, ld_OH, Dr, ld_Dr, Dt, ld_Dt, S, ld_S, conv, ld_conv);
    if (dtype == JSTSP_F64) return run_sparse_admm<double>(h, mem, Mr, Mt, batch, imax, Htrue, ld_H, OH, ld_OH, Dr, ld_Dr, Dt, ld_Dt, S, ld_S, conv, ld_conv);
    return fail(h, JSTSP_E_ARG, "unknown dtype");
}
