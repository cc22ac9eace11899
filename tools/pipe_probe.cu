// Pipe-rate probe for B200 (sm_100a): FFMA vs packed fma.rn.f32x2 vs DFMA, plus a
// shared-memory-broadcast + FFMA2 mix shaped like the complex GEMM inner loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long da = *reinterpret_cast<unsigned long long*>(&d);
    unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a);
    unsigned long long bb = *reinterpret_cast<unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&da);
}

template<int ITERS> __global__ void k_ffma(float* out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ITERS> __global__ void k_ffma2(float* out, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i);
    float2 aa = make_float2(a, a), bb = make_float2(b, -b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) ffma2(acc[i], aa, bb);   // acc += aa*bb
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int ITERS> __global__ void k_dfma(float* out, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}
// complex-GEMM-like mix: per step one LDS.128 broadcast (ar,ar,ai,ai) feeds RB complex MACs (2 FFMA2 each)
template<int ITERS, int RB> __global__ void k_mix(float* out, float b0) {
    __shared__ float4 sh[256];
    sh[threadIdx.x] = make_float4(threadIdx.x * 1e-3f, threadIdx.x * 1e-3f, 1e-3f, 1e-3f);
    __syncthreads();
    float2 acc[RB]; float2 br[RB], bi[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) { acc[i] = make_float2(0, 0); br[i] = make_float2(b0 + i, b0 - i); bi[i] = make_float2(-(b0 - i), b0 + i); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float4 a = sh[(it * 8 + k) & 255];      // warp-uniform address => broadcast
#pragma unroll
            for (int i = 0; i < RB; ++i) { ffma2(acc[i], make_float2(a.x, a.y), br[i]); ffma2(acc[i], make_float2(a.z, a.w), bi[i]); }
        }
    }
    float s = 0; for (int i = 0; i < RB; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    const int ITERS = 4096; int blocks = p.multiProcessorCount * 8, threads = 256;
    float* out; CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
#define RUN(name, launch, fma_per_thread) \
    for (int w = 0; w < 3; ++w) { launch; } CK(cudaDeviceSynchronize()); \
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) { launch; } cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); \
    cudaEventElapsedTime(&ms, e0, e1); ms /= 5; \
    printf("%-28s %8.3f ms  %8.2f TFLOP/s  (%.1f fma/clk/SM @ %.0f MHz nominal)\n", name, ms, 2.0 * (double)(fma_per_thread) * blocks * threads / ms / 1e9, \
           (double)(fma_per_thread) * blocks * threads / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1e3);
    RUN("ffma 3-reg", (k_ffma<ITERS><<<blocks, threads>>>(out, 1.0001f, 0.5f)), (double)ITERS * 16)
    RUN("fma.rn.f32x2", (k_ffma2<ITERS><<<blocks, threads>>>(out, 1.0001f, 0.5f)), (double)ITERS * 32)
    RUN("dfma", (k_dfma<ITERS / 4><<<blocks, threads>>>(out, 1.0001, 0.5)), (double)ITERS / 4 * 16)
    RUN("mix lds128 + 8 cmac (ffma2)", (k_mix<ITERS / 8, 8><<<blocks, threads>>>(out, 0.5f)), (double)ITERS / 8 * 8 * 8 * 4)
    RUN("mix lds128 + 16 cmac (ffma2)", (k_mix<ITERS / 8, 16><<<blocks, threads>>>(out, 0.5f)), (double)ITERS / 8 * 8 * 16 * 4)
    RUN("mix lds128 + 4 cmac (ffma2)", (k_mix<ITERS / 8, 4><<<blocks, threads>>>(out, 0.5f)), (double)ITERS / 8 * 8 * 4 * 4)
    return 0;
}
