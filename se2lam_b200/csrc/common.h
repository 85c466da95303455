// Shared host-side helpers of libse2gpu (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/se2gpu.h"

namespace se2gpu {

extern thread_local std::string g_last_error;
extern std::atomic<unsigned long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define SE2_CUDA(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::se2gpu::fail(SE2GPU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                                   \
    } while (0)

// every kernel launch of the library goes through this so that se2gpu_launch_count() is exact
#define SE2_LAUNCH(kernel, grid, block, smem, stream, ...)      \
    do {                                                        \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        ::se2gpu::g_launches.fetch_add(1, std::memory_order_relaxed); \
    } while (0)

inline int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) return fail(SE2GPU_ERR_NO_DEVICE, "no CUDA device available (%s)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(SE2GPU_ERR_INVALID, "device %d out of range (%d devices)", device, n);
    SE2_CUDA(cudaSetDevice(device));
    return SE2GPU_OK;
}

// Accumulates per-group kernel time with CUDA events recorded on the launching stream.
struct Profiler {
    static constexpr int MAXG = 16, MAXEV = 4096;
    bool on = false;
    int nev = 0;
    cudaEvent_t ev[MAXEV][2];
    int grp[MAXEV];
    bool created = false;
    double ms[MAXG] = {0};
    int launches[MAXG] = {0};
    void enable(bool e) {
        if (e && !created) { for (int i = 0; i < MAXEV; ++i) { cudaEventCreate(&ev[i][0]); cudaEventCreate(&ev[i][1]); } created = true; }
        on = e; nev = 0;
        for (int g = 0; g < MAXG; ++g) { ms[g] = 0; launches[g] = 0; }
    }
    void flush() {
        for (int i = 0; i < nev; ++i) {
            cudaEventSynchronize(ev[i][1]);
            float t = 0;
            cudaEventElapsedTime(&t, ev[i][0], ev[i][1]);
            ms[grp[i]] += t; launches[grp[i]]++;
        }
        nev = 0;
    }
    inline void begin(int g, cudaStream_t s) { if (on) { if (nev == MAXEV) flush(); grp[nev] = g; cudaEventRecord(ev[nev][0], s); } }
    inline void end(cudaStream_t s) { if (on) { cudaEventRecord(ev[nev][1], s); ++nev; } }
    ~Profiler() { if (created) for (int i = 0; i < MAXEV; ++i) { cudaEventDestroy(ev[i][0]); cudaEventDestroy(ev[i][1]); } }
};

// NVTX ranges around every kernel group and host entry point (visible in Nsight Systems / ncu --nvtx; a no-op when no tool is attached)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define SE2_NVTX_CAT2(a, b) a##b
#define SE2_NVTX_CAT(a, b) SE2_NVTX_CAT2(a, b)
#define SE2_NVTX(name) ::se2gpu::NvtxRange SE2_NVTX_CAT(_se2_nvtx_, __LINE__)(name)

template <class T>
inline cudaError_t dev_alloc(T** p, size_t count) {
    return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
}

}  // namespace se2gpu
