// umma_probe.cu - developer probe for the tcgen05 (UMMA) building blocks of the fused ADMM kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_probe tools/umma_probe.cu && build/umma_probe
// Checks, against a CPU model, on one CTA:
//   (1) kind::tf32 SS MMA with K-major A (128 x K) / K-major B (64 x K), SWIZZLE_NONE core-matrix layout;
//   (2) the SAME shared-memory bytes viewed as an MN-major A operand (rows = the 16-byte direction);
//   (3) whether fp32 inputs are truncated or rounded to tf32 by the tensor core;
//   (4) the 32x32b TMEM load mapping (lane = row, column = n);
// and times back-to-back MMAs for N = 32 / 64 to see the issue rate of shared-memory-sourced tf32.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
    return d;                        // layout_type 0 = SWIZZLE_NONE
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Stage geometry under test (the fused kernel's): core matrix = 8 rows x 16 bytes (4 floats), 128 contiguous bytes.
//   pass-1 view: big operand [128 rows m][K = 32 floats], core matrix (mg, kg) at ((kg * 16 + mg) * 128) bytes   -> K-major A, SBO = 128, LBO = 2048
//   pass-2 view: big operand rows = the 16-byte direction: [128 rows q = (kg', j)][K = 32 m], core matrix (mg', kg') at ((kg' * 4 + mg') * 128)
//                                                                                                -> MN-major A, SBO = 512, LBO = 128
//   small operand [64 rows n][K = 32 floats], core matrix (ng, kg) at ((kg * 8 + ng) * 128)                    -> K-major B, SBO = 128, LBO = 1024
constexpr int KS = 32;
__global__ void __launch_bounds__(128) probe(const float* __restrict__ gA, const float* __restrict__ gB, float* __restrict__ gD, int mode, int nrep, long long* cyc) {
    __shared__ __align__(128) float sA[128 * KS];
    __shared__ __align__(128) float sB[64 * KS];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // operands: gA is the LOGICAL matrix A[row][k] (128 x 32, row-major), gB is B[n][k] (64 x 32, row-major)
    for (int e = tid; e < 128 * KS; e += 128) {
        const int row = e / KS, k = e % KS;
        int off;
        if (mode == 0) off = ((k / 4) * 16 + row / 8) * 32 + (row % 8) * 4 + (k % 4);          // K-major: 8 rows x 4 k per core matrix
        else           off = ((row / 4) * 4 + k / 8) * 32 + (k % 8) * 4 + (row % 4);           // MN-major: 8 k x 4 rows per core matrix
        sA[off] = gA[e];
    }
    for (int e = tid; e < 64 * KS; e += 128) {
        const int n = e / KS, k = e % KS;
        sB[((k / 4) * 8 + n / 8) * 32 + (n % 8) * 4 + (k % 4)] = gB[e];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 64, mode == 0 ? 0 : 1, 0);
        long long t0 = clock64();
        for (int rep = 0; rep < nrep; ++rep) {
#pragma unroll
            for (int ks = 0; ks < KS / 8; ++ks) {
                uint64_t da, db;
                if (mode == 0) da = make_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128);
                else           da = make_desc(smem_u32(sA) + ks * 128, 128, 512);
                db = make_desc(smem_u32(sB) + ks * 2 * 1024, 1024, 128);
                umma_tf32(tm, da, db, idesc, (rep | ks) ? 1u : 0u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (cyc) cyc[0] = t1 - t0;
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + half * 32, v);
        for (int j = 0; j < 32; ++j) gD[(size_t)tid * 64 + half * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64u));
}

// issue-rate probe: nrep x (4 k-steps) of M=128, N=nn MMAs on the same operands
__global__ void __launch_bounds__(128) rate(int mm, int nn, int nrep, long long* cyc) {
    __shared__ __align__(128) float sA[128 * KS];
    __shared__ __align__(128) float sB[224 * KS];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int e = tid; e < 128 * KS; e += 128) sA[e] = 1.0f;
    for (int e = tid; e < 224 * KS; e += 128) sB[e] = 0.5f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(mm, nn, 0, 0);
        long long t0 = clock64();
        for (int rep = 0; rep < nrep; ++rep) {
#pragma unroll
            for (int ks = 0; ks < KS / 8; ++ks) {
                uint64_t da = make_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128);
                uint64_t db = make_desc(smem_u32(sB) + ks * 2 * 3584, 3584, 128);      // 28 row groups per k group
                umma_tf32(tm, da, db, idesc, 1u);
            }
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        cyc[0] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float rn_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    std::vector<float> A(128 * KS), B(64 * KS), D(128 * 64);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    float *dA, *dB, *dD; long long* dc;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    int bad = 0;
    for (int mode = 0; mode < 2; ++mode) {
        CK(cudaMemset(dD, 0, D.size() * 4));
        probe<<<1, 128>>>(dA, dB, dD, mode, 1, dc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double e_tr = 0, e_rn = 0, e_full = 0, ref_max = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < 64; ++n) {
                double st = 0, sr = 0, sf = 0;
                for (int k = 0; k < KS; ++k) {
                    st += (double)trunc_tf32(A[r * KS + k]) * trunc_tf32(B[n * KS + k]);
                    sr += (double)rn_tf32(A[r * KS + k]) * rn_tf32(B[n * KS + k]);
                    sf += (double)A[r * KS + k] * B[n * KS + k];
                }
                double d = D[r * 64 + n];
                e_tr = fmax(e_tr, fabs(d - st)); e_rn = fmax(e_rn, fabs(d - sr)); e_full = fmax(e_full, fabs(d - sf)); ref_max = fmax(ref_max, fabs(sf));
            }
        printf("mode %d (%s A): max|D-trunc model| %.3e  max|D-rn model| %.3e  max|D-fp32| %.3e  (max|ref| %.2f)\n", mode, mode ? "MN-major" : "K-major", e_tr, e_rn, e_full, ref_max);
        if (e_tr > 2e-5 && e_rn > 2e-5) { printf("  MISMATCH in mode %d\n", mode); bad = 1; }
    }
    for (int mm : {128, 64})
        for (int nn : {16, 32, 64, 96, 128, 192, 224}) {
            const int nrep = 512;
            rate<<<1, 128>>>(mm, nn, nrep, dc);
            CK(cudaDeviceSynchronize());
            long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
            printf("rate: M=%d N=%d K=8 SS tf32: %.1f cycles/MMA\n", mm, nn, (double)c / (nrep * 4));
        }
    if (getenv("PROBE_MAP")) {
        // mapping dump for the MN-major view: B = selector of k, A = row index / k index
        for (int what = 0; what < 2; ++what) {
            for (int r = 0; r < 128; ++r) for (int k = 0; k < KS; ++k) A[r * KS + k] = what ? (float)k : (float)r;
            for (int n = 0; n < 64; ++n) for (int k = 0; k < KS; ++k) B[n * KS + k] = (n == k) ? 1.f : 0.f;
            CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
            probe<<<1, 128>>>(dA, dB, dD, 1, 1, dc);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            printf("MN-major map, D[r][n] should be %s:\n", what ? "n (the k index)" : "r");
            for (int r : {0, 1, 2, 3, 4, 5, 8, 9, 16, 33, 64, 127}) { printf(" r=%3d:", r); for (int n = 0; n < 34; ++n) printf(" %3.0f", D[r * 64 + n]); printf("\n"); }
        }
    }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
