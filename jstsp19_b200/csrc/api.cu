// api.cu - handle management for libjstsp_b200 (see include/jstsp_b200.h).
#include "common.cuh"

using namespace jstsp;

extern "C" const char* jstsp_version(void) { return "jstsp19_b200 0.1.0 (sm_100a)"; }

extern "C" int jstsp_create(jstsp_handle** out, int device) {
    if (!out) return JSTSP_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) { cudaGetLastError(); return JSTSP_E_CUDA; }   // no CPU fallback by design
    if (device < 0 || device >= ndev) return JSTSP_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return JSTSP_E_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return JSTSP_E_CUDA;
    if (prop.major < 10) return JSTSP_E_CUDA;   // kernels are compiled for sm_100a only
    jstsp_handle* h = new jstsp_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return JSTSP_E_CUDA; }
    h->own_stream = true;
    {   // the eigen-solves on the side stream are short, latency-bound and on the critical path of the next iteration: their CTAs go first
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi) != cudaSuccess) { delete h; return JSTSP_E_CUDA; }
    }
    if (cudaStreamCreateWithFlags(&h->copy, cudaStreamNonBlocking) != cudaSuccess) { delete h; return JSTSP_E_CUDA; }
    for (int k = 0; k < 2; ++k) { cudaEventCreateWithFlags(&h->ev_in[k], cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_done[k], cudaEventDisableTiming); }
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_check, cudaEventDisableTiming);
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->flags_host), 2 * sizeof(int), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); delete h; return JSTSP_E_CUDA; }
    cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (cudaMalloc(&h->d_flag, sizeof(int)) != cudaSuccess) { delete h; return JSTSP_E_CUDA; }
    *out = h;
    return JSTSP_OK;
}

extern "C" void jstsp_destroy(jstsp_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->ws) cudaFree(h->ws);
    if (h->d_flag) cudaFree(h->d_flag);
    for (auto e : h->prof.pool) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_check) cudaEventDestroy(h->ev_check);
    if (h->flags_host) cudaFreeHost(h->flags_host);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->copy) cudaStreamDestroy(h->copy);
    for (int k = 0; k < 2; ++k) { if (h->ev_in[k]) cudaEventDestroy(h->ev_in[k]); if (h->ev_done[k]) cudaEventDestroy(h->ev_done[k]); }
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char* jstsp_last_error(const jstsp_handle* h) { return h ? h->err.c_str() : "NULL handle"; }

extern "C" int jstsp_set_stream(jstsp_handle* h, void* cuda_stream) {
    if (!h) return JSTSP_E_ARG;
    cudaStream_t ns = static_cast<cudaStream_t>(cuda_stream);
    if (ns == h->stream) return JSTSP_OK;
    // The handle's workspace (ADMM state, Jacobi warm start, flags) is shared by every call: work queued on the old stream may still be
    // using it, so the new stream waits for the old one before anything else is enqueued.
    if (h->stream) {
        if (cudaEventRecord(h->ev_fork, h->stream) == cudaSuccess) cudaStreamWaitEvent(ns, h->ev_fork, 0);
        else cudaGetLastError();                  // the caller destroyed the old stream: its work has completed
    }
    if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->stream = ns;
    h->own_stream = false;
    return JSTSP_OK;
}

extern "C" long long jstsp_nonfinite_count(jstsp_handle* h) {
    if (!h || !h->d_flag) return -1;
    int bad = 0;
    if (cudaSetDevice(h->device) != cudaSuccess) return -1;
    if (cudaMemcpyAsync(&bad, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -1;
    return bad;
}

extern "C" int jstsp_synchronize(jstsp_handle* h) {
    if (!h) return JSTSP_E_ARG;
    JSTSP_CUDA(h, cudaSetDevice(h->device));
    JSTSP_CUDA(h, cudaStreamSynchronize(h->stream));
    return JSTSP_OK;
}

extern "C" long long jstsp_launch_count(const jstsp_handle* h) { return h ? h->launches : 0; }

extern "C" int jstsp_set_chunk(jstsp_handle* h, int max_trials_per_pass) {
    if (!h || max_trials_per_pass < 0) return JSTSP_E_ARG;
    h->max_chunk = max_trials_per_pass;
    return JSTSP_OK;
}

static const char* kProfNames[PK_COUNT] = {"xupd_t1", "res", "q", "vupd", "xs", "eig", "setup", "svt_step", "omp", "other", "fused_tc", "expand_as", "fused_psi", "psi_res", "psi_g", "psi_step", "omp_kron_corr", "somp", "omp_kron_corr_tc", "psi_mega", "lg_state", "lg_pass1", "lg_pass2", "lg_small"};

extern "C" int jstsp_profile(jstsp_handle* h, int enable) {
    if (!h) return JSTSP_E_ARG;
    prof_collect(h);
    h->prof.on = enable != 0;
    // the event pool is sized here, not at the first profiled launches: creating timing events inside a region that is being timed cost its
    // first step ~10 ms (bench.py trace)
    if (enable) while (h->prof.pool.size() < 1024) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) break; h->prof.pool.push_back(e); }
    if (enable == 2) { for (int i = 0; i < PK_COUNT; ++i) { h->prof.total_ms[i] = 0; h->prof.count[i] = 0; } }
    return JSTSP_OK;
}

extern "C" int jstsp_profile_read(jstsp_handle* h, int slot, double* total_ms, long long* launches, const char** name) {
    if (!h || slot < 0) return JSTSP_E_ARG;
    if (slot >= PK_COUNT) return 1;   // past the last slot
    prof_collect(h);
    if (total_ms) *total_ms = h->prof.total_ms[slot];
    if (launches) *launches = h->prof.count[slot];
    if (name) *name = kProfNames[slot];
    return JSTSP_OK;
}

// Developer hook (not part of the drop-in surface): device buffer that receives in-kernel
// clock64() phase timestamps, 8 slots per CTA, for tools/phase_probe.py.  NULL disables.
extern "C" int jstsp_debug_buffer(jstsp_handle* h, void* device_buffer) {
    if (!h) return JSTSP_E_ARG;
    h->dbg = static_cast<long long*>(device_buffer);
    return JSTSP_OK;
}
