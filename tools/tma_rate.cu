// tma_rate.cu - how fast does one SM's TMA unit move bulk copies of a given size?  (developer probe, B200)
// One CTA per SM, one thread issues `cp.async.bulk` copies of `bytes` from an L2-resident buffer into a shared-memory ring of `depth` slots
// and waits for each slot in order.  Prints cycles per copy and bytes per cycle per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_rate tools/tma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {      // non-blocking test in a loop
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {   // try_wait with a suspend-time hint
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__global__ void k(const unsigned char* src, size_t span, int bytes, int pieces, int depth, int n, long long* out, int mode) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bars[16];
    if (threadIdx.x == 0) {
        for (int s = 0; s < depth; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const unsigned char* base = src + (size_t)blockIdx.x * span;
        const int piece = bytes / pieces;
        long long t0 = clock64();
        for (int i = 0; i < n + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) { const uint32_t ph = ((i / depth) - 1) & 1; if (mode == 0) mbar_wait(&bars[s], ph); else if (mode == 1) mbar_spin(&bars[s], ph); else mbar_hint(&bars[s], ph, 100u); }
            if (i < n) {
                mbar_expect_tx(&bars[s], bytes);
                const unsigned char* g = base + ((size_t)i * bytes) % (span - bytes);
                for (int q = 0; q < pieces; ++q) bulk(sm + (size_t)s * bytes + q * piece, g + (size_t)q * piece, piece, &bars[s]);
            }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}
int main() {
    const size_t span = 512 * 1024;       // per SM: 148 x 512 KB = 76 MB, L2 resident after the first pass
    unsigned char* src; cudaMalloc(&src, span * 148); cudaMemset(src, 1, span * 148);
    long long* out; cudaMalloc(&out, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cfg[][3] = {{2176, 1, 8}, {2176, 1, 32}, {12288, 1, 2}, {12288, 1, 4}, {16384, 1, 4}, {16384, 1, 8}, {34816, 1, 2}, {34816, 16, 2}, {34816, 1, 4}, {34816, 16, 4}, {65536, 1, 2}, {65536, 1, 3}};
    for (auto& c : cfg) {
        const int bytes = c[0], pieces = c[1], depth = c[2], n = 400;
        for (int mode : {0, 1, 2}) {
            const int grid = 148;
            for (int rep = 0; rep < 2; ++rep) k<<<grid, 32, (size_t)bytes * depth>>>(src, span, bytes, pieces, depth, n, out, mode);
            long long h[148]; cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
            printf("bytes %6d in %2d pieces, depth %2d, wait mode %d (0 try_wait, 1 test_wait spin, 2 try_wait + 100 ns hint): %8.0f cycles per copy, %6.1f B/cycle/SM  (%s)\n", bytes, pieces, depth, mode, avg / n, (double)bytes * n / avg, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
