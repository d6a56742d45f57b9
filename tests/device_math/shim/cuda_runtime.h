// TEST INFRASTRUCTURE.  A stand-in for <cuda_runtime.h> that lets g++ compile the product's
// device-function header (spectrobot_b200/csrc/sr_device.cuh) for the HOST, so that the arithmetic
// the kernels are built from can be checked on a box without a GPU (tests/test_device_math.py).
// Nothing in the product includes this file; the product has no CPU path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__

static inline float __double2float_rn(double x) { return (float)x; }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    std::memcpy(&d, &u, 8);
    return d;
}
static inline int __double2loint(double x) {
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return (int)(uint32_t)u;
}
static inline int __double2hiint(double x) {
    uint64_t u;
    std::memcpy(&u, &x, 8);
    return (int)(uint32_t)(u >> 32);
}
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }

// MUFU.RCP64H (rcp.approx.ftz.f64): reads the upper 32 bits of the operand and returns about 20
// good bits with a zero lower word.  The table itself is not public; what the callers rely on is
// only that error level (one or two Newton steps follow), which this model reproduces.
static inline double sr_host_rcp_approx(double x) {
    uint64_t u;
    std::memcpy(&u, &x, 8);
    u &= 0xFFFFFFFF00000000ull;
    double xt;
    std::memcpy(&xt, &u, 8);
    double r = 1.0 / xt;
    std::memcpy(&u, &r, 8);
    u &= 0xFFFFFFFF00000000ull;
    std::memcpy(&r, &u, 8);
    return r;
}
// the one inline-PTX statement of the header: asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#define asm(...) r = sr_host_rcp_approx(x)
