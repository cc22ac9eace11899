// tma_rate2.cu - TMA bulk-copy issue cost on one SM: K copies issued back to back by one thread (or by two threads of different warps), then one wait.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_rate2 tools/tma_rate2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// each issuing thread: `rounds` times { K copies to K barriers back to back ; wait for all K }
__global__ void k(const unsigned char* src, size_t span, int bytes, int K, int rounds, int issuers, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bars[64];
    const int w = threadIdx.x / 32;
    if (threadIdx.x == 0) { for (int s = 0; s < 64; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x % 32 == 0 && w < issuers) {
        const unsigned char* g = src + (size_t)blockIdx.x * span + (size_t)w * (span / 2);
        unsigned char* dst = sm + (size_t)w * K * bytes;
        uint64_t* b = bars + w * 32;
        long long t0 = clock64(), t_issue = 0;
        for (int r = 0; r < rounds; ++r) {
            long long a = clock64();
            for (int i = 0; i < K; ++i) { mbar_expect_tx(&b[i], bytes); bulk(dst + (size_t)i * bytes, g + (size_t)i * bytes, bytes, &b[i]); }
            t_issue += clock64() - a;
            for (int i = 0; i < K; ++i) mbar_wait(&b[i], r & 1);
        }
        out[blockIdx.x * 4 + w * 2] = clock64() - t0;
        out[blockIdx.x * 4 + w * 2 + 1] = t_issue;
    }
}
int main() {
    const size_t span = 512 * 1024;
    unsigned char* src; cudaMalloc(&src, span * 148); cudaMemset(src, 1, span * 148);
    long long* out; cudaMalloc(&out, 148 * 4 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cfg[][2] = {{2176, 1}, {2176, 4}, {2176, 16}, {16384, 1}, {16384, 2}, {16384, 4}, {12288, 4}, {34816, 1}, {34816, 2}};
    for (auto& c : cfg)
        for (int issuers : {1, 2}) {
            const int bytes = c[0], K = c[1], rounds = 200;
            if ((size_t)bytes * K * issuers > 190 * 1024) continue;
            for (int rep = 0; rep < 2; ++rep) k<<<148, 64, (size_t)bytes * K * issuers>>>(src, span, bytes, K, rounds, issuers, out);
            long long h[148 * 4]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double tot = 0, iss = 0; for (int i = 0; i < 148; ++i) { tot += h[4 * i]; iss += h[4 * i + 1]; } tot /= 148; iss /= 148;
            printf("bytes %6d x K %2d per round, %d issuing warp(s): round %7.0f cycles (%6.0f per copy), issue part %6.0f (%5.0f per copy), %6.1f B/cycle/SM  (%s)\n", bytes, K, issuers,
                   tot / rounds, tot / rounds / K, iss / rounds, iss / rounds / K, (double)bytes * K * issuers * rounds / tot, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
