// umma_probe3.cu - developer probe, third stage (bf16 lo tiles): the two views of the dictionary tile exactly as the fused ADMM kernel
// uses them, fed by TMA tensor copies from the dictionary's natural layout (row m = 2P contiguous floats).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_probe2 tools/umma_probe2.cu && build/umma_probe2
//   view 1 (X_s = (A S) B):   D1[m][n] = sum_k Bm[m][k] S[n][k]     A operand K-major, SWIZZLE_128B, box (32 floats, 128 m)
//   view 2 (T1  = K B^H):     D2[q][n] = sum_m Bm[m][q] Kq[n][m]    A operand MN-major, SWIZZLE_128B_ATOM_32B (the only MN-major
//                                                                   tf32 layout), box (32 floats, 32 m, 4 groups of 32 floats)
// The small operand is K-major SWIZZLE_NONE in both (written by ordinary stores).  Also times the MMA issue rate of both views.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;     // 0 none, 1 128B_BASE32B, 2 128B, 4 64B, 6 32B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


#include <cuda_bf16.h>
constexpr int MR = 128, KF = 128, NS = 32;     // Bm is MR x KF bf16, small operand NS x 128 bf16
constexpr int STAGE_BYTES = 8192;

// view 1: D[m][n] = sum_k Bm[m][k] S[n][k]   A K-major SWIZZLE_64B (box 32 bf16 x 128 rows), K = 16 per MMA
// view 2: D[q][n] = sum_m Bm[m][q] S[n][m]   A MN-major SWIZZLE_128B (box 64 bf16 x 32 rows x 2 groups)
__global__ void __launch_bounds__(128) probe3(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2, const __nv_bfloat16* __restrict__ gS,
                                              float* __restrict__ gD, int view) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* stage = smem;                                   // 4 x 8 KiB
    __nv_bfloat16* sS = reinterpret_cast<__nv_bfloat16*>(smem + 4 * STAGE_BYTES);  // 32 x 128, K-major SWIZZLE_NONE: core matrix 8 rows x 8 bf16
    __shared__ __align__(8) uint64_t full[4];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32;
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) mbar_init(&full[s], 1);
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    for (int e = tid; e < NS * 128; e += 128) {
        const int n = e / 128, k = e % 128;
        sS[((k / 8) * (NS / 8) + n / 8) * 64 + (n % 8) * 8 + (k % 8)] = gS[e];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            mbar_expect_tx(&full[s], STAGE_BYTES);
            if (view == 1) tma_2d(stage + s * STAGE_BYTES, &map1, 32 * s, 0, &full[s]);          // 32 bf16 of k x 128 rows m
            else           tma_3d(stage + s * STAGE_BYTES, &map2, 0, 32 * s, 0, &full[s]);       // 64 bf16 x 32 rows m x 2 groups
        }
        const uint32_t idesc = make_idesc(128, NS, view == 1 ? 0 : 1, 0);
        for (int s = 0; s < 4; ++s) mbar_wait(&full[s], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int s = 0; s < 4; ++s) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {                       // 32 k per stage, 16 per MMA
                const uint32_t sa = smem_u32(stage + s * STAGE_BYTES);
                uint64_t da = view == 1 ? make_desc(sa + ks * 32, 16, 512, 4)             // K-major SW64: 8-row groups 512 B apart, k-step = 32 B inside the 64-byte row
                                        : make_desc(sa + ks * 2048, 4096, 1024, 2);       // MN-major SW128: 64-element groups 4096 B apart, 8-row k groups 1024 B apart
                uint64_t db = make_desc(smem_u32(sS) + (s * 2 + ks) * 2 * (NS / 8) * 128, (NS / 8) * 128, 128, 0);
                umma_bf16(tm, da, db, idesc, (s | ks) ? 1u : 0u);
            }
        }
        umma_commit(&done);
    }
    __syncthreads();
    mbar_wait(&done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
    for (int j = 0; j < 32; ++j) gD[(size_t)tid * 32 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32u));
}

int main() {
    std::vector<__nv_bfloat16> Bm(MR * KF), S(NS * 128);
    std::vector<float> Bf(MR * KF), Sf(NS * 128), D(128 * 32);
    srand(3);
    for (size_t i = 0; i < Bm.size(); ++i) { Bm[i] = __float2bfloat16((float)rand() / RAND_MAX * 2.f - 1.f); Bf[i] = __bfloat162float(Bm[i]); }
    for (size_t i = 0; i < S.size(); ++i) { S[i] = __float2bfloat16((float)rand() / RAND_MAX * 2.f - 1.f); Sf[i] = __bfloat162float(S[i]); }
    __nv_bfloat16 *dB, *dS; float* dD;
    CK(cudaMalloc(&dB, Bm.size() * 2)); CK(cudaMalloc(&dS, S.size() * 2)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dB, Bm.data(), Bm.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dS, S.data(), S.size() * 2, cudaMemcpyHostToDevice));
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    { void* fn = nullptr; cudaDriverEntryPointQueryResult qr; CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr)); encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn); }
    CUtensorMap map1, map2;
    {
        cuuint64_t dims[2] = {KF, MR}; cuuint64_t strides[1] = {KF * 2}; cuuint32_t box[2] = {32, 128}; cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&map1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode map1 (2D bf16, SW64): %d\n", (int)r); if (r) return 1;
    }
    {
        cuuint64_t dims[3] = {64, MR, KF / 64}; cuuint64_t strides[2] = {KF * 2, 128}; cuuint32_t box[3] = {64, 32, 2}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = encode(&map2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dB, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode map2 (3D bf16, SW128): %d\n", (int)r); if (r) return 1;
    }
    const size_t smem = 4 * STAGE_BYTES + NS * 128 * 2 + 1024;
    CK(cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int bad = 0;
    for (int view = 1; view <= 2; ++view) {
        CK(cudaMemset(dD, 0, D.size() * 4));
        probe3<<<1, 128, smem>>>(map1, map2, dS, dD, view);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double err = 0, ref_max = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < NS; ++n) {
                double st = 0;
                for (int k = 0; k < 128; ++k) st += (double)(view == 1 ? Bf[r * KF + k] : Bf[k * KF + r]) * Sf[n * 128 + k];
                err = fmax(err, fabs(D[r * 32 + n] - st)); ref_max = fmax(ref_max, fabs(st));
            }
        printf("bf16 view %d: max|D - exact| %.3e (max|ref| %.2f)\n", view, err, ref_max);
        if (err > 5e-5) { bad = 1; printf("  MISMATCH view %d; D[0][0..7] =", view); for (int j = 0; j < 8; ++j) printf(" %.4f", D[j]); printf("\n"); }
    }
    printf(bad ? "PROBE3 FAILED\n" : "PROBE3 OK\n");
    return bad;
}
