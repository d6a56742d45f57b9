// sr_common.h -- host-side plumbing shared by the translation units of libspectrobot.so
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include "../../include/spectrobot.h"

namespace sr {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define SR_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return sr::fail(SR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,       \
                            cudaGetErrorString(e__));                                     \
    } while (0)

// every kernel launch in the library goes through this so that sr_kernel_launch_count() is exact
#define SR_LAUNCH(kernel, grid, block, smem, stream, ...)                                 \
    do {                                                                                  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                       \
        sr::g_launches.fetch_add(1, std::memory_order_relaxed);                           \
        SR_CUDA(cudaGetLastError());                                                      \
    } while (0)

// Optional per-kernel timing (sr_prof_enable): CUDA events recorded on the launching stream right
// before and after selected launches, summed by sr_prof_summary.  bench.py uses it for the live
// roofline of the dominant kernel; off by default (no events, no overhead).
extern std::atomic<int> g_prof_on;
void prof_begin(int kind, double work, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    bool on;
    ProfScope(int kind, double work, cudaStream_t s) : st(s), on(g_prof_on.load() != 0) {
        if (on) prof_begin(kind, work, st);
    }
    ~ProfScope() { if (on) prof_end(st); }
};

template <typename T>
struct DevBuf {  // RAII device buffer
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc(&p, count * sizeof(T));
    }
    cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s = 0) {
        cudaError_t e = ensure(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

// Stream-ordered scratch (cudaMallocAsync): freed blocks stay in the device's default pool instead
// of going back to the driver at every synchronisation, so per-call scratch costs no cudaMalloc.
inline void pool_keep() {
    static std::atomic<unsigned long long> done{0};   // one bit per device (a process may use several)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    if (done.load() & (1ull << dev)) return;
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    done.fetch_or(1ull << dev);
}

template <typename T>
struct PoolBuf {  // RAII stream-ordered device buffer
    T* p = nullptr;
    cudaStream_t st = 0;
    PoolBuf() = default;
    PoolBuf(const PoolBuf&) = delete;
    PoolBuf& operator=(const PoolBuf&) = delete;
    ~PoolBuf() { if (p) cudaFreeAsync(p, st); }
    cudaError_t alloc(size_t count, cudaStream_t s) {
        pool_keep();
        st = s;
        return cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), s);
    }
    cudaError_t upload(const T* h, size_t count, cudaStream_t s) {
        cudaError_t e = alloc(count, s);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

}  // namespace sr
