// rng.cu - the random draws of the Monte-Carlo trial loop, generated on the device by a counter-based generator.
// The reference draws with MATLAB's global stream (randn / rand / randsrc / randperm: wideband_mmwave_channel.m:19-22,
// plot_errorVSsnr.m:60,63-67, qam4mod.m:7-8, proposed_hbf.m:37); no seed is fixed anywhere in it, so only the distributions are
// part of its behaviour.  Here every number is Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11) of
//     counter = (block index within the stream, stream id, global trial index lo, hi),   key = (seed lo, seed hi)
// so a trial gets the same numbers whichever batch, shard or GPU count computes it, with no state to carry.
// Streams: 0 path-gain normals, 1 angle uniforms, 2 pilot symbols, 3 noise, 4 sampling order.
#include "common.cuh"

namespace jstsp {

__host__ __device__ inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0, hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct DrawP {
    uint32_t k0, k1; long long first; int batch;
    int Nr, Nt, L, Np, T;
    const double* sigma2;
    double *normals, *uniforms; void *pilots, *noise; int* perm;
};
__device__ __forceinline__ void draw4(const DrawP& p, int b, uint32_t stream, uint32_t idx, uint32_t (&r)[4]) {
    const unsigned long long t = (unsigned long long)(p.first + b);
    philox4x32_10(idx, stream, (uint32_t)t, (uint32_t)(t >> 32), p.k0, p.k1, r);
}
__device__ __forceinline__ double u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

// path gains and angles: one thread per (trial, tap, ray)    normals / uniforms [b][L][Np][2]
__global__ void k_draw_chan(DrawP p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, per = p.L * p.Np;
    if (i >= p.batch * per) return;
    const int b = i / per, j = i % per;
    uint32_t r[4];
    draw4(p, b, 0, j, r);
    const double rad = sqrt(-2.0 * log(u01(r[0])));
    double sn, cs; sincospi(2.0 * u01(r[1]), &sn, &cs);
    p.normals[2 * (size_t)i] = rad * cs; p.normals[2 * (size_t)i + 1] = rad * sn;
    draw4(p, b, 1, j, r);
    p.uniforms[2 * (size_t)i] = u01(r[0]); p.uniforms[2 * (size_t)i + 1] = u01(r[1]);
}
// 4-QAM pilots (qam4mod.m:7-8: randsrc over [1+1j, -1+1j, 1-1j, -1-1j] / sqrt 2): 2 bits per symbol, 64 symbols per block; one thread per
// 4 symbols (the block is recomputed by its 16 threads: 60 integer instructions against a coalesced 32-byte store).  [b][T][Nt]
template <typename T>
__global__ void k_draw_pilots(DrawP p) {
    const size_t per = (size_t)p.T * p.Nt;
    const size_t e0 = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (e0 >= per) return;
    const int b = blockIdx.y;
    uint32_t r[4];
    draw4(p, b, 2, (uint32_t)(e0 / 64), r);
    cx<T>* o = reinterpret_cast<cx<T>*>(p.pilots) + (size_t)b * per;
    const T a = (T)0.70710678118654752440;
    const uint32_t word = r[(e0 % 64) / 16];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (e0 + q < per) { const uint32_t s = (word >> (2 * ((e0 + q) % 16))) & 3u; o[e0 + q] = mk<T>((s & 1u) ? -a : a, (s & 2u) ? -a : a); }
}
// noise N = sqrt(sigma2 / 2) (randn + j randn) (plot_errorVSsnr.m:60): one thread per two samples (Box-Muller on both halves of a block).  [b][T][Nr]
template <typename T>
__global__ void k_draw_noise(DrawP p) {
    const size_t per = (size_t)p.T * p.Nr, calls = (per + 1) / 2;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= calls) return;
    const int b = blockIdx.y;
    uint32_t r[4];
    draw4(p, b, 3, (uint32_t)i, r);
    cx<T>* o = reinterpret_cast<cx<T>*>(p.noise) + (size_t)b * per;
    const T sc = (T)sqrt(p.sigma2[b] * 0.5);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        T re, im;
        if (sizeof(T) == 4) {
            const float u1 = ((float)(r[2 * w] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(r[2 * w + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float rad = sqrtf(-2.0f * logf(u1));
            float sn, cs; sincospif(2.0f * u2, &sn, &cs);
            re = (T)(rad * cs); im = (T)(rad * sn);
        } else {
            const double rad = sqrt(-2.0 * log(u01(r[2 * w])));
            double sn, cs; sincospi(2.0 * u01(r[2 * w + 1]), &sn, &cs);
            re = (T)(rad * cs); im = (T)(rad * sn);
        }
        const size_t e = 2 * i + w;
        if (e < per) o[e] = mk<T>(sc * re, sc * im);
    }
}
// sampling order of each training instant = randperm(Nr) (proposed_hbf.m:37): rows sorted by one 32-bit key each (ties by row).  [b][T][Nr], 1-based
__global__ void k_draw_perm(DrawP p) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (m >= p.T) return;
    const int Nr = p.Nr, cpc = (Nr + 3) / 4;
    unsigned long long key[64];
    for (int j = 0; j < cpc; ++j) {
        uint32_t r[4];
        draw4(p, b, 4, (uint32_t)(m * cpc + j), r);
#pragma unroll
        for (int w = 0; w < 4; ++w) if (4 * j + w < Nr) key[4 * j + w] = ((unsigned long long)r[w] << 8) | (unsigned)(4 * j + w);
    }
    for (int a = 1; a < Nr; ++a) {                      // insertion sort, Nr <= 64
        const unsigned long long v = key[a];
        int c = a - 1;
        while (c >= 0 && key[c] > v) { key[c + 1] = key[c]; --c; }
        key[c + 1] = v;
    }
    int* o = p.perm + ((size_t)b * p.T + m) * Nr;
    for (int a = 0; a < Nr; ++a) o[a] = (int)(key[a] & 0xFFu) + 1;
}

}  // namespace jstsp

using namespace jstsp;

extern "C" void jstsp_philox4x32_10(const unsigned* ctr, const unsigned* key, unsigned* out) {
    uint32_t r[4];
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], r);
    for (int i = 0; i < 4; ++i) out[i] = r[i];
}

extern "C" int jstsp_draw_trials(jstsp_handle* h, int dtype, unsigned long long seed, long long first_trial, int batch, int Nr, int Nt, int L, int Np, int T,
                                 const double* sigma2, double* normals, double* uniforms, void* pilots, void* noise, int* perm) {
    if (!h) return JSTSP_E_ARG;
    if (batch <= 0 || Nr <= 0 || Nt <= 0 || L <= 0 || Np <= 0 || T <= 0 || first_trial < 0) return fail(h, JSTSP_E_ARG, "non-positive dimension");
    if (Nr > 64) return fail(h, JSTSP_E_UNSUPPORTED, "jstsp_draw_trials: the sampling order is drawn for up to 64 rows");
    if (noise && !sigma2) return fail(h, JSTSP_E_ARG, "noise needs the per-trial variances");
    if (dtype != JSTSP_F32 && dtype != JSTSP_F64) return fail(h, JSTSP_E_ARG, "unknown dtype");
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    DrawP p{(uint32_t)seed, (uint32_t)(seed >> 32), first_trial, batch, Nr, Nt, L, Np, T, sigma2, normals, uniforms, pilots, noise, perm};
    cudaStream_t st = h->stream;
    if (normals && uniforms) JSTSP_LAUNCH(h, PK_OTHER, (k_draw_chan<<<ceil_div(batch * L * Np, 128), 128, 0, st>>>(p)));
    if (pilots) {
        dim3 g((unsigned)ceil_div_ll(ceil_div_ll((long long)T * Nt, 4), 256), batch);
        if (dtype == JSTSP_F32) JSTSP_LAUNCH(h, PK_OTHER, (k_draw_pilots<float><<<g, 256, 0, st>>>(p)));
        else JSTSP_LAUNCH(h, PK_OTHER, (k_draw_pilots<double><<<g, 256, 0, st>>>(p)));
    }
    if (noise) {
        dim3 g((unsigned)ceil_div_ll(ceil_div_ll((long long)T * Nr, 2), 256), batch);
        if (dtype == JSTSP_F32) JSTSP_LAUNCH(h, PK_OTHER, (k_draw_noise<float><<<g, 256, 0, st>>>(p)));
        else JSTSP_LAUNCH(h, PK_OTHER, (k_draw_noise<double><<<g, 256, 0, st>>>(p)));
    }
    if (perm) { dim3 g(ceil_div(T, 128), batch); JSTSP_LAUNCH(h, PK_OTHER, (k_draw_perm<<<g, 128, 0, st>>>(p))); }
    JSTSP_CUDA(h, cudaGetLastError());
    return JSTSP_OK;
}
