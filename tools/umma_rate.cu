// umma_rate.cu - developer probe: tcgen05 kind::tf32 SS issue rate with FRESH operands (a different shared-memory tile per MMA)
// versus a re-used tile, for the dictionary tile on the M side (A operand) and on the N side (B operand).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_rate tools/umma_rate.cu && build/umma_rate
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// big: 160 KiB of shared memory holding K-major SWIZZLE_NONE tiles; the "big" operand walks through it (fresh) or stays put
__global__ void __launch_bounds__(128) rate(int mm, int nn, int big_on_n, int fresh, int nrep, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid / 32;
    float* f = reinterpret_cast<float*>(smem);
    for (int e = tid; e < 40 * 1024 + 2048; e += 128) f[e] = 1.0f;      // 160 KiB big region + 8 KiB small operand
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(mm, nn);
        const uint32_t big = smem_u32(smem), small = smem_u32(smem + 160 * 1024);
        const int big_rows = big_on_n ? nn : mm;
        const uint32_t tile_bytes = (uint32_t)big_rows * 32;            // one K = 8 slab: rows x 32 bytes, core matrices: 2 k-groups x (rows/8) x 128 B
        const int ntiles = fresh ? (160 * 1024) / tile_bytes : 1;
        long long t0 = clock64();
        for (int i = 0; i < nrep; ++i) {
            const uint32_t tb = big + (uint32_t)(i % ntiles) * tile_bytes;
            const uint64_t dbig = make_desc(tb, (uint32_t)(big_rows / 8) * 128, 128);
            const uint64_t dsm = make_desc(small, (uint32_t)((big_on_n ? mm : nn) / 8) * 128, 128);
            if (big_on_n) umma_tf32(tm, dsm, dbig, idesc); else umma_tf32(tm, dbig, dsm, idesc);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        mbar_wait(&bar, 0);
        cyc[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}
int main() {
    long long* dc; CK(cudaMalloc(&dc, 8 * 512));
    const size_t smem = 160 * 1024 + 8 * 1024;
    CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int nrep = 4096;
    struct Cfg { int mm, nn, big_on_n; } cfgs[] = {{128, 64, 0}, {128, 32, 0}, {64, 64, 0}, {128, 128, 0}, {128, 256, 0}, {64, 256, 1}, {128, 256, 1}, {64, 128, 1}, {128, 128, 1}};
    for (auto c : cfgs)
        for (int fresh = 0; fresh < 2; ++fresh)
            for (int grid : {1, 148}) {
                rate<<<grid, 128, smem>>>(c.mm, c.nn, c.big_on_n, fresh, nrep, dc);
                CK(cudaDeviceSynchronize());
                long long h[148]; CK(cudaMemcpy(h, dc, 8 * grid, cudaMemcpyDeviceToHost));
                double s = 0; for (int i = 0; i < grid; ++i) s += (double)h[i];
                printf("M=%3d N=%3d big tile on %s side, %s operands, grid %3d: %.1f cycles/MMA\n", c.mm, c.nn, c.big_on_n ? "N" : "M", fresh ? "fresh " : "reused", grid, s / grid / nrep);
            }
    return 0;
}
