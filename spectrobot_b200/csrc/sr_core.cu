// sr_core.cu -- library globals, device selection, FP64 peak micro-benchmark
#include "sr_common.h"
#include <mutex>
#include <vector>

namespace sr {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};

namespace {
struct ProfRec { int kind; double work; cudaEvent_t a, b; };
std::mutex g_prof_mtx;
std::vector<ProfRec> g_prof_recs;
thread_local ProfRec g_prof_open{-1, 0.0, nullptr, nullptr};
}  // namespace

void prof_begin(int kind, double work, cudaStream_t st) {
    ProfRec r{kind, work, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    g_prof_open = r;
}
void prof_end(cudaStream_t st) {
    if (g_prof_open.kind < 0) return;
    cudaEventRecord(g_prof_open.b, st);
    std::lock_guard<std::mutex> g(g_prof_mtx);
    g_prof_recs.push_back(g_prof_open);
    g_prof_open.kind = -1;
}
}  // namespace sr

namespace {

// Dependent-chain FP64 FMA benchmark: 8 independent chains per thread, `iters` rounds.
// Gives the sustained DFMA rate the K1/K2 roofline is quoted against (the FP64 peak
// must be measured, not taken from the spec sheet).
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double* sink) {
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;  // keeps the chains alive
}

}  // namespace

extern "C" {

int sr_version(void) { return 100; }

const char* sr_last_error(void) { return sr::g_err; }

long long sr_kernel_launch_count(void) { return sr::g_launches.load(); }

int sr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int sr_set_device(int device) {
    SR_CUDA(cudaSetDevice(device));
    return SR_OK;
}

int sr_prof_enable(int on) {
    std::lock_guard<std::mutex> g(sr::g_prof_mtx);
    for (auto& r : sr::g_prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    sr::g_prof_recs.clear();
    sr::g_prof_on.store(on ? 1 : 0);
    return SR_OK;
}

int sr_prof_summary(int kind, long long* launches, double* ms, double* work) {
    if (!launches || !ms || !work) return sr::fail(SR_ERR_ARG, "sr_prof_summary: bad argument");
    SR_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> g(sr::g_prof_mtx);
    *launches = 0;
    *ms = 0.0;
    *work = 0.0;
    for (auto& r : sr::g_prof_recs) {
        if (r.kind != kind) continue;
        float t = 0.f;
        SR_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
        *launches += 1;
        *ms += t;
        *work += r.work;
    }
    return SR_OK;
}

int sr_fp64_peak(int iters, double* flops_per_s) {
    if (iters < 1 || !flops_per_s) return sr::fail(SR_ERR_ARG, "sr_fp64_peak: bad argument");
    int dev = 0, sms = 0;
    SR_CUDA(cudaGetDevice(&dev));
    SR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sr::DevBuf<double> sink;
    SR_CUDA(sink.alloc(1));
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    SR_CUDA(cudaEventCreate(&e0));
    SR_CUDA(cudaEventCreate(&e1));
    SR_LAUNCH(k_fp64_peak, blocks, threads, 0, 0, iters / 8 + 1, 1.0, sink.p);  // warm-up
    SR_CUDA(cudaEventRecord(e0, 0));
    SR_LAUNCH(k_fp64_peak, blocks, threads, 0, 0, iters, 1.0, sink.p);
    SR_CUDA(cudaEventRecord(e1, 0));
    SR_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    SR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
    *flops_per_s = flops / (ms * 1e-3);
    return SR_OK;
}

}  // extern "C"
