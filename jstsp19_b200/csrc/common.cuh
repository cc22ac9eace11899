// common.cuh - shared device/host helpers for libjstsp_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jstsp_b200.h"

namespace jstsp {

// ---------------------------------------------------------------------------
// complex scalar (interleaved, MATLAB -R2018a layout)
// ---------------------------------------------------------------------------
template <typename T>
struct __align__(2 * sizeof(T)) cx {
    T re, im;
};
template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T re, T im) { cx<T> r; r.re = re; r.im = im; return r; }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator*(cx<T> a, cx<T> b) { return mk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
template <typename T> __host__ __device__ __forceinline__ cx<T> operator*(T s, cx<T> a) { return mk<T>(s * a.re, s * a.im); }
template <typename T> __host__ __device__ __forceinline__ cx<T> conj(cx<T> a) { return mk<T>(a.re, -a.im); }
template <typename T> __host__ __device__ __forceinline__ T abs2(cx<T> a) { return a.re * a.re + a.im * a.im; }

// acc += a * b   /   acc += a * conj(b)
template <typename T> __device__ __forceinline__ void cmac(T& ar, T& ai, T xr, T xi, T yr, T yi) {
    ar = fma(xr, yr, ar); ar = fma(-xi, yi, ar);
    ai = fma(xr, yi, ai); ai = fma(xi, yr, ai);
}

// soft threshold with MATLAB sign(0) = 0 (proposed_algorithm.m:56)
template <typename T> __device__ __forceinline__ T soft1(T x, T thr) {
    T a = fabs(x) - thr;
    a = a > T(0) ? a : T(0);
    return x > T(0) ? a : (x < T(0) ? -a : T(0));
}

constexpr int kWarp = 32;

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------
// host-side handle
// ---------------------------------------------------------------------------
// Optional per-kernel-class device timing (CUDA events on the launching stream), switched on
// by jstsp_profile(); bench.py reads it for the live roofline numbers.
enum { PK_XUPD_T1 = 0, PK_RES, PK_Q, PK_VUPD, PK_XS, PK_EIG, PK_SETUP, PK_SVT_STEP, PK_OMP, PK_OTHER, PK_FUSED_TC, PK_EXPAND, PK_FUSED_PSI, PK_PSI_AUX, PK_PSI_G, PK_PSI_STEP, PK_OMP_CORR, PK_SOMP, PK_OMP_CORR_TC, PK_PSI_MEGA, PK_LG_STATE, PK_LG_PASS1, PK_LG_PASS2, PK_LG_SMALL, PK_COUNT };
struct Prof {
    bool on = false;
    struct Rec { int slot; cudaEvent_t a, b; };
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::vector<Rec> recs;
    double total_ms[PK_COUNT] = {};
    long long count[PK_COUNT] = {};
};

struct Handle {
    Prof prof;
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;      // stream all work is enqueued on
    cudaStream_t side = nullptr;        // side stream for the eigen-solves (forked/joined by events)
    bool own_stream = false;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_check = nullptr;     // the structure-check flags of the current pass have landed in flags_host
    int* flags_host = nullptr;          // pinned, 2 ints
    cudaStream_t copy = nullptr;        // H2D stream of the HOST-buffer path (next pass's inputs travel while this pass computes)
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    std::string err;
    std::string bad_launch;             // first kernel launch the runtime rejected (slot and source line), reported with the next error
    long long launches = 0;
    int max_chunk = 0;
    int last_variant = 0;               // structured tcgen05 path: 1 = one persistent kernel per pass (admm_mega.cuh), 0 = four kernels per iteration
    int last_path = 0;                  // structured entry: 1 = dense kernels on the materialised B, 2 = Psi-domain tcgen05 kernel (last pass)
    // simple grow-only workspace
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int* d_flag = nullptr;              // device-side non-finite counter
    long long* dbg = nullptr;           // optional device buffer for in-kernel phase timestamps (tools/phase_probe.py)
};

}  // namespace jstsp

struct jstsp_handle : public jstsp::Handle {};

namespace jstsp {

inline int fail(Handle* h, int code, const std::string& msg) {
    if (h) {
        h->err = msg;
        if (!h->bad_launch.empty()) { h->err += " [first rejected launch: " + h->bad_launch + "]"; h->bad_launch.clear(); }
    }
    return code;
}

#define JSTSP_CUDA(h, expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            return ::jstsp::fail((h), JSTSP_E_CUDA,                                               \
                                 std::string(#expr) + ": " + cudaGetErrorString(_e));             \
        }                                                                                         \
    } while (0)

inline cudaEvent_t prof_event(Handle* h) {
    Prof& p = h->prof;
    if (p.used == p.pool.size()) { cudaEvent_t e; cudaEventCreate(&e); p.pool.push_back(e); }
    return p.pool[p.used++];
}
inline void prof_begin(Handle* h, int slot) {
    if (!h->prof.on) return;
    Prof::Rec r{slot, prof_event(h), prof_event(h)};
    cudaEventRecord(r.a, h->stream);
    h->prof.recs.push_back(r);
}
inline void prof_end(Handle* h) {
    if (!h->prof.on) return;
    cudaEventRecord(h->prof.recs.back().b, h->stream);
}
inline void prof_collect(Handle* h) {
    Prof& p = h->prof;
    if (p.recs.empty()) return;
    cudaStreamSynchronize(h->stream);
    for (auto& r : p.recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { p.total_ms[r.slot] += ms; p.count[r.slot]++; }
    }
    p.recs.clear();
    p.used = 0;
}
// launch wrapper: counts the launch and (when profiling) brackets it with events
#define JSTSP_LAUNCH(h, slot, ...)          \
    do {                                    \
        ::jstsp::prof_begin((h), (slot));   \
        __VA_ARGS__;                        \
        if (cudaPeekAtLastError() != cudaSuccess && (h)->bad_launch.empty())                             \
            (h)->bad_launch = std::string(#slot) + " at " + __FILE__ + ":" + std::to_string(__LINE__);   \
        ::jstsp::prof_end((h));             \
        (h)->launches++;                    \
    } while (0)

// Bump allocator over the handle's workspace.
struct Arena {
    char* base;
    size_t cap, off;
    Arena(void* b, size_t c) : base(static_cast<char*>(b)), cap(c), off(0) {}
    template <typename U> U* take(size_t n) {
        size_t bytes = (n * sizeof(U) + 255) & ~size_t(255);
        char* p = base ? base + off : nullptr;
        off += bytes;
        return reinterpret_cast<U*>(p);
    }
};

inline int ensure_workspace(Handle* h, size_t bytes) {
    if (bytes <= h->ws_bytes) return JSTSP_OK;
    if (h->ws) {
        cudaStreamSynchronize(h->stream);
        cudaFree(h->ws);
        h->ws = nullptr; h->ws_bytes = 0;
    }
    cudaError_t e = cudaMalloc(&h->ws, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(h, JSTSP_E_NOMEM, "workspace cudaMalloc failed: " + std::to_string(bytes) + " bytes"); }
    h->ws_bytes = bytes;
    return JSTSP_OK;
}

template <typename K>
inline int set_smem_named(Handle* h, const char* what, K kernel, size_t bytes) {
    if (bytes > h->smem_optin)
        return fail(h, JSTSP_E_UNSUPPORTED, std::string("kernel needs more shared memory than the device offers: ") + what + " wants " + std::to_string(bytes) + " bytes");
    if (bytes > 32 * 1024) {                 // dynamic + static shared memory may cross the 48 KiB default together
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fail(h, JSTSP_E_CUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    }
    return JSTSP_OK;
}
#define set_smem(h, ...) set_smem_named((h), #__VA_ARGS__, __VA_ARGS__)   // variadic: kernel template-ids contain commas

}  // namespace jstsp
