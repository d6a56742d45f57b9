// sr_device.cuh -- device-side Humlicek-w4 pieces of humliv_bb (lineshape.f:226-569)
//
// What must be reproduced to stay within 1e-6 of the reference:
//   * core points: the complex argument is built with Fortran CMPLX() without KIND, i.e. both
//     parts are rounded to float32 before the complex*16 arithmetic (lineshape.f:529);
//   * the region-3/4 coefficients are real*4 literals widened to double (lineshape.f:539-558);
//   * complex division with Smith's range reduction (gfortran -fcx-fortran-rules);
//   * region 1/2 are real FP64 rationals in x^2 (lineshape.f:456-477, 492-522).
#pragma once
#include <cuda_runtime.h>

namespace srdev {

struct cplx { double re, im; };

__device__ __forceinline__ cplx c_mul(cplx a, cplx b) {
    cplx r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
__device__ __forceinline__ cplx c_add_r(double r, cplx a) { a.re += r; return a; }
__device__ __forceinline__ cplx c_r_sub(double r, cplx a) { a.re = r - a.re; a.im = -a.im; return a; }
// real part of a/b with Smith's algorithm (only the real part is ever used, lineshape.f:546,560)
__device__ __forceinline__ double c_div_re(cplx a, cplx b) {
    if (fabs(b.re) < fabs(b.im)) {
        double ratio = b.re / b.im;
        double div = b.re * ratio + b.im;
        return (a.re * ratio + a.im) / div;
    } else {
        double ratio = b.im / b.re;
        double div = b.im * ratio + b.re;
        return (a.im * ratio + a.re) / div;
    }
}

#define SRF(x) ((double)x##f)  // real*4 literal widened to double

// regions 3 and 4 (lineshape.f:527-561); rx >= 0 and ry in FP64, rounded to float32 inside
static __device__ __noinline__ double humliv_core(double rx, double ry) {
    double r2 = 0.195 * rx - 0.176;
    cplx c2;
    c2.re = (double)__double2float_rn(ry);
    c2.im = (double)__double2float_rn(-rx);
    if (ry < r2) {  // region 4
        cplx c1 = c_mul(c2, c2);
        cplx num;
        num.re = c1.re * SRF(.56419);
        num.im = c1.im * SRF(.56419);
        num = c_r_sub(SRF(1.320522), num);
        num = c_r_sub(SRF(35.76683), c_mul(c1, num));
        num = c_r_sub(SRF(219.0313), c_mul(c1, num));
        num = c_r_sub(SRF(1540.787), c_mul(c1, num));
        num = c_r_sub(SRF(3321.9905), c_mul(c1, num));
        num = c_r_sub(SRF(36183.31), c_mul(c1, num));
        num = c_mul(c2, num);
        cplx den = c_r_sub(SRF(1.841439), c1);
        den = c_r_sub(SRF(61.57037), c_mul(c1, den));
        den = c_r_sub(SRF(364.2191), c_mul(c1, den));
        den = c_r_sub(SRF(2186.181), c_mul(c1, den));
        den = c_r_sub(SRF(9022.228), c_mul(c1, den));
        den = c_r_sub(SRF(24322.84), c_mul(c1, den));
        den = c_r_sub(SRF(32066.6), c_mul(c1, den));
        return exp(c1.re) * cos(c1.im) - c_div_re(num, den);
    } else {  // region 3
        cplx num;
        num.re = c2.re * SRF(.5642236);
        num.im = c2.im * SRF(.5642236);
        num = c_add_r(SRF(3.778987), num);
        num = c_add_r(SRF(11.96482), c_mul(c2, num));
        num = c_add_r(SRF(20.20933), c_mul(c2, num));
        num = c_add_r(SRF(16.4955), c_mul(c2, num));
        cplx den = c_add_r(SRF(6.699398), c2);
        den = c_add_r(SRF(21.69274), c_mul(c2, den));
        den = c_add_r(SRF(39.27121), c_mul(c2, den));
        den = c_add_r(SRF(38.82363), c_mul(c2, den));
        den = c_add_r(SRF(16.4955), c_mul(c2, den));
        return c_div_re(num, den);
    }
}
// region 2 (lineshape.f:492-522), x2 = x*x
static __device__ __noinline__ double humliv_reg2(double x2, double ry) {
    double ry2 = ry * ry;
    double a = ry * (1.0578555 + ry2 * (4.6545642 + ry2 * (3.1030428 + 0.5641896 * ry2)));
    double b = ry * (2.9619954 + ry2 * (0.5641896 + 1.6925688 * ry2));
    double c = ry * (-2.5388532 + ry2 * 1.6925688);
    double d = ry * 0.5641896;
    double e = 0.5625 + ry2 * (4.5 + ry2 * (10.5 + ry2 * (6. + ry2)));
    double f = -4.5 + ry2 * (9. + ry2 * (6. + 4. * ry2));
    double g = 10.5 + ry2 * (-6. + 6. * ry2);
    double h = 4. * ry2 - 6.;
    return (a + x2 * (b + x2 * (c + d * x2))) / (e + x2 * (f + x2 * (g + x2 * (h + x2))));
}

// ---- variants for the batched centre kernel (k_core_eval) ------------------------------------
// Same formulas; the Horner steps are written as FMA pairs (r - c*n and r + c*n in 4 instructions
// instead of 6), the quotient's real part as (n.d)/(d.d) with one division instead of Smith's two,
// and the region-2 coefficients are hoisted out of the point loop.  Differences from the forms
// above are rounding-level (<= a few 1e-16 of the largest term).
struct reg2_coef { double a, b, c, d, e, f, g, h; };
__device__ __forceinline__ reg2_coef humliv_reg2_coefs(double ry) {
    const double ry2 = ry * ry;
    reg2_coef k;
    k.a = ry * (1.0578555 + ry2 * (4.6545642 + ry2 * (3.1030428 + 0.5641896 * ry2)));
    k.b = ry * (2.9619954 + ry2 * (0.5641896 + 1.6925688 * ry2));
    k.c = ry * (-2.5388532 + ry2 * 1.6925688);
    k.d = ry * 0.5641896;
    k.e = 0.5625 + ry2 * (4.5 + ry2 * (10.5 + ry2 * (6. + ry2)));
    k.f = -4.5 + ry2 * (9. + ry2 * (6. + 4. * ry2));
    k.g = 10.5 + ry2 * (-6. + 6. * ry2);
    k.h = 4. * ry2 - 6.;
    return k;
}
__device__ __forceinline__ double humliv_reg2_eval(const reg2_coef& k, double x2) {
    return (k.a + x2 * (k.b + x2 * (k.c + k.d * x2))) /
           (k.e + x2 * (k.f + x2 * (k.g + x2 * (k.h + x2))));
}
// n <- r - c*n
__device__ __forceinline__ cplx h_rsub(double r, cplx c, cplx n) {
    cplx o;
    o.re = fma(-c.re, n.re, fma(c.im, n.im, r));
    o.im = -fma(c.re, n.im, c.im * n.re);
    return o;
}
// n <- r + c*n
__device__ __forceinline__ cplx h_radd(double r, cplx c, cplx n) {
    cplx o;
    o.re = fma(c.re, n.re, fma(-c.im, n.im, r));
    o.im = fma(c.re, n.im, c.im * n.re);
    return o;
}
__device__ __forceinline__ double c_div_re_direct(cplx a, cplx b) {
    return fma(a.re, b.re, a.im * b.im) / fma(b.re, b.re, b.im * b.im);
}
__device__ __forceinline__ double humliv_core_fast(double rx, double ry) {
    const double r2 = 0.195 * rx - 0.176;
    cplx c2;
    c2.re = (double)__double2float_rn(ry);
    c2.im = (double)__double2float_rn(-rx);
    if (ry < r2) {  // region 4
        const cplx c1 = c_mul(c2, c2);
        cplx num;
        num.re = c1.re * SRF(.56419);
        num.im = c1.im * SRF(.56419);
        num = c_r_sub(SRF(1.320522), num);
        num = h_rsub(SRF(35.76683), c1, num);
        num = h_rsub(SRF(219.0313), c1, num);
        num = h_rsub(SRF(1540.787), c1, num);
        num = h_rsub(SRF(3321.9905), c1, num);
        num = h_rsub(SRF(36183.31), c1, num);
        num = c_mul(c2, num);
        cplx den = c_r_sub(SRF(1.841439), c1);
        den = h_rsub(SRF(61.57037), c1, den);
        den = h_rsub(SRF(364.2191), c1, den);
        den = h_rsub(SRF(2186.181), c1, den);
        den = h_rsub(SRF(9022.228), c1, den);
        den = h_rsub(SRF(24322.84), c1, den);
        den = h_rsub(SRF(32066.6), c1, den);
        return exp(c1.re) * cos(c1.im) - c_div_re_direct(num, den);
    } else {  // region 3
        cplx num;
        num.re = c2.re * SRF(.5642236);
        num.im = c2.im * SRF(.5642236);
        num = c_add_r(SRF(3.778987), num);
        num = h_radd(SRF(11.96482), c2, num);
        num = h_radd(SRF(20.20933), c2, num);
        num = h_radd(SRF(16.4955), c2, num);
        cplx den = c_add_r(SRF(6.699398), c2);
        den = h_radd(SRF(21.69274), c2, den);
        den = h_radd(SRF(39.27121), c2, den);
        den = h_radd(SRF(38.82363), c2, den);
        den = h_radd(SRF(16.4955), c2, den);
        return c_div_re_direct(num, den);
    }
}

#undef SRF

// region 1 exactly as written in the Fortran (lineshape.f:456-477); used by the Tier-1 drop-in
__device__ __forceinline__ double humliv_reg1(double x2, double ry) {
    double ry2 = ry * ry;
    double a = ry * (1.1283792 + 2.2567584 * ry2);
    double b = 2.2567584 * ry;
    double c = (1. + 2. * ry2) * (1. + 2. * ry2);
    double d = -4. + 8. * ry2;
    return (a + x2 * b) / (c + x2 * (d + 4. * x2));
}

// Region 1 rewritten for the hot loop.  With q = 0.5 + ry^2 and s = x^2 the Fortran rational is
//   K = (a + s b)/(c + s(d + 4 s)) = (b/4) * w / (w^2 - 2 s),  w = q + s
//     = (b/4) * (u + 1) / (u^2 + 2 ry^2),                     u = s + ry^2 - 0.5
// (a/b = q, c = 4 q^2, d = 8 q - 8).  Returns K/(b/4) given u and c2 = 2 ry^2.
// The reciprocal is MUFU.RCP64H (rcp.approx.ftz.f64, ~2^-20 relative) plus one Newton step
// folded into the product: w*r0*(1+e), e = 1 - den*r0  ->  relative error ~1e-12.
__device__ __forceinline__ double rcp_approx(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}
__device__ __forceinline__ double humliv_reg1_fast(double u, double c2) {
    double den = fma(u, u, c2);
    double w = u + 1.0;
    double r0 = rcp_approx(den);
    double e = fma(-den, r0, 1.0);
    double t = w * r0;
    return fma(t, e, t);
}

// same with w = u + 1 supplied by the caller (computed in parallel with u)
__device__ __forceinline__ double humliv_reg1_uw(double u, double w, double c2) {
    double den = fma(u, u, c2);
    double r0 = rcp_approx(den);
    double e = fma(-den, r0, 1.0);
    double t = w * r0;
    return fma(t, e, t);
}

// same in 5 FP64 instructions + MUFU: (u + 1) / den = u r + r with r = r0 (2 - den r0).  One
// instruction fewer than the (u, w) form at the price of one more level in the dependency chain -
// the tile kernel is bound by instruction issue, not by latency.
__device__ __forceinline__ double humliv_reg1_u(double u, double c2) {
    const double den = fma(u, u, c2);
    const double r0 = rcp_approx(den);
    const double e = fma(-den, r0, 1.0);
    const double r = fma(r0, e, r0);
    return fma(u, r, r);
}

__device__ __forceinline__ long long f_nint(double x) { return llround(x); }  // Fortran NINT

// exp(x) and expm1(x) from ONE range reduction and ONE polynomial (layer update of the LOS
// recursion needs both: the transmission exp(-tau) must stay accurate when tau is large, the
// emission weight 1-exp(-tau) when tau is tiny).
//   x = n ln2 + r, |r| <= ln2/2;  p = expm1(r) (Taylor to r^13, |err| < 5e-18 |r|)
//   exp(x) = 2^n (1 + p);  expm1(x) = 2^n p + (2^n - 1)
__device__ __forceinline__ void exp_pair(double x, double& ex, double& em) {
    const double n = rint(x * 1.4426950408889634);
    if (!(n > -1000.0 && n < 1000.0)) {   // |x| > ~693 or NaN: rare, take the library path
        ex = exp(x);
        em = expm1(x);
        return;
    }
    double r = fma(n, -6.93147180369123816490e-01, x);   // ln2 hi
    r = fma(n, -1.90821492927058770002e-10, r);          // ln2 lo
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p * r, r, r);                                // expm1(r)
    const int ni = (int)n;
    const double s = __hiloint2double((ni + 1023) << 20, 0);   // 2^n
    ex = fma(s, p, s);
    em = (ni == 0) ? p : fma(s, p, s - 1.0);
}

// exp(-t) and phi(t) = (1 - e^{-t})/t (1 at t = 0) of the layer update with the emission given as J
// (I <- I e^{-t} + J phi(t)), in ~18 FP64 instructions on the common path
// (|t| < ln2/2) instead of exp_pair + an IEEE division:
//   x = -t = n ln2 + r (round-to-nearest by the 1.5*2^52 trick, no FRND/F2I on the XU pipe),
//   g(r) = expm1(r)/r = sum_k r^k/(k+1)!  (degree 13, |err| < 2e-17 for |r| <= ln2/2),
//   e^x = 2^n (1 + r g);   phi = expm1(x)/x = g when n = 0, else (e^x - 1)/x with the quotient
//   from MUFU.RCP64H + two Newton steps (no cancellation: |x| >= ln2/2 there).
// Relative error of both results <= 4e-16 (checked against 60-digit arithmetic).
__device__ __forceinline__ void exp_phi(double t, double& ex, double& phi) {
    const double x = -t;
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double z = fma(x, 1.4426950408889634, MAGIC);
    const int ni = __double2loint(z);                        // n as an integer
    const double n = z - MAGIC;
    if (!(n > -1000.0 && n < 1000.0)) {                      // |t| > ~693 or NaN: library path
        ex = exp(x);
        phi = (t == 0.0) ? 1.0 : -expm1(x) / t;
        return;
    }
    double r = fma(n, -6.93147180369123816490e-01, x);       // ln2 hi
    r = fma(n, -1.90821492927058770002e-10, r);              // ln2 lo
    double g = 1.0 / 87178291200.0;                          // 1/14!
    g = fma(g, r, 1.0 / 6227020800.0);
    g = fma(g, r, 1.0 / 479001600.0);
    g = fma(g, r, 1.0 / 39916800.0);
    g = fma(g, r, 1.0 / 3628800.0);
    g = fma(g, r, 1.0 / 362880.0);
    g = fma(g, r, 1.0 / 40320.0);
    g = fma(g, r, 1.0 / 5040.0);
    g = fma(g, r, 1.0 / 720.0);
    g = fma(g, r, 1.0 / 120.0);
    g = fma(g, r, 1.0 / 24.0);
    g = fma(g, r, 1.0 / 6.0);
    g = fma(g, r, 0.5);
    g = fma(g, r, 1.0);                                      // expm1(r)/r
    ex = fma(r, g, 1.0);                                     // e^r
    phi = g;
    if (ni != 0) {
        const double s = __hiloint2double((ni + 1023) << 20, 0);   // 2^n
        ex *= s;
        double r0 = rcp_approx(x);
        r0 = fma(fma(-x, r0, 1.0), r0, r0);
        r0 = fma(fma(-x, r0, 1.0), r0, r0);
        phi = (ex - 1.0) * r0;
    }
}
__device__ __forceinline__ double layer_update_j(double I, double t, double J, bool solo) {
    double ex, phi;
    exp_phi(t, ex, phi);
    return solo ? I * ex : fma(I, ex, J * phi);
}

// Short forms of the update for layers stored as float32, chosen per warp and step by the caller
// (k_los_layers_f32): 85 % of the warps of a limb batch see |t| < 1e-2 on all their points.
//   tier 0  |t| < 1e-2 : g = expm1(x)/x to degree 4 (|err| < 1.4e-13), no range reduction: 7 FP64 ops
//   tier 1  |t| < ln2/2: degree 9 (|err| < 3e-11), no range reduction: 12 FP64 ops
__device__ __forceinline__ double layer_update_j_small(double I, double t, double J, bool solo) {
    const double x = -t;
    double g = fma(1.0 / 120.0, x, 1.0 / 24.0);
    g = fma(g, x, 1.0 / 6.0);
    g = fma(g, x, 0.5);
    g = fma(g, x, 1.0);
    const double ex = fma(x, g, 1.0);
    return solo ? I * ex : fma(I, ex, J * g);
}
__device__ __forceinline__ double layer_update_j_medium(double I, double t, double J, bool solo) {
    const double x = -t;
    double g = 1.0 / 3628800.0;
    g = fma(g, x, 1.0 / 362880.0);
    g = fma(g, x, 1.0 / 40320.0);
    g = fma(g, x, 1.0 / 5040.0);
    g = fma(g, x, 1.0 / 720.0);
    g = fma(g, x, 1.0 / 120.0);
    g = fma(g, x, 1.0 / 24.0);
    g = fma(g, x, 1.0 / 6.0);
    g = fma(g, x, 0.5);
    g = fma(g, x, 1.0);
    const double ex = fma(x, g, 1.0);
    return solo ? I * ex : fma(I, ex, J * g);
}

// The same update for layers that were stored as float32 (k_los_layers_f32): t and J carry a
// relative rounding of 6e-8 already, so g(r) is cut at degree 9 (|err| < 3e-11 for |r| <= ln2/2)
// and the quotient takes one Newton step (1e-13): 5 FP64 instructions fewer per update, and the
// error added by the arithmetic stays three orders of magnitude below that of the inputs.
__device__ __forceinline__ double layer_update_j_f32in(double I, double t, double J, bool solo) {
    const double x = -t;
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double z = fma(x, 1.4426950408889634, MAGIC);
    const int ni = __double2loint(z);
    const double n = z - MAGIC;
    if (!(n > -1000.0 && n < 1000.0)) return layer_update_j(I, t, J, solo);
    double r = fma(n, -6.93147180369123816490e-01, x);
    r = fma(n, -1.90821492927058770002e-10, r);
    double g = 1.0 / 3628800.0;                              // 1/10!
    g = fma(g, r, 1.0 / 362880.0);
    g = fma(g, r, 1.0 / 40320.0);
    g = fma(g, r, 1.0 / 5040.0);
    g = fma(g, r, 1.0 / 720.0);
    g = fma(g, r, 1.0 / 120.0);
    g = fma(g, r, 1.0 / 24.0);
    g = fma(g, r, 1.0 / 6.0);
    g = fma(g, r, 0.5);
    g = fma(g, r, 1.0);                                      // expm1(r)/r
    double ex = fma(r, g, 1.0);
    double phi = g;
    if (ni != 0) {
        ex *= __hiloint2double((ni + 1023) << 20, 0);
        double r0 = rcp_approx(x);
        r0 = fma(fma(-x, r0, 1.0), r0, r0);
        phi = (ex - 1.0) * r0;
    }
    return solo ? I * ex : fma(I, ex, J * phi);
}

// ---------------------------------------------------------------------------------------------
// One segment of the Curtis-Godson integrals curgod_fort_1..4 (curgods.f:2-98): number density
// piecewise exponential, vmr (and f in variant 3) piecewise linear, nd*f exponential in variant 4.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double curgod_seg1(double nd0, double nd1, double dx) {   // :14-19
    const double fu = nd1 / nd0;
    const double D = log(fu) / dx;
    return (nd1 - nd0) / D;
}
__device__ __forceinline__ double curgod_seg2(double nd0, double nd1, double v0, double v1,
                                              double dx) {                             // :35-42
    const double A = nd0 * v0;
    const double B = nd0 * (v1 - v0) / dx;
    const double fu = nd1 / nd0;
    const double D = log(fu) / dx;
    return (A * D * (fu - 1.) + B * fu * (D * dx - 1.) + B) / (D * D);
}
__device__ __forceinline__ double curgod_seg3(double nd0, double nd1, double v0, double v1,
                                              double f0, double f1, double dx) {       // :58-70
    const double A = nd0 * v0 * f0;
    const double cc = (v1 - v0) / dx;
    const double bb = (f1 - f0) / dx;
    const double B = nd0 * (v0 * bb + f0 * cc);
    const double Cc = nd0 * bb * cc;
    const double fu = nd1 / nd0;
    const double D = log(fu) / dx;
    return (fu * (D * (A * D + B * (D * dx - 1.)) + Cc * (D * dx * (D * dx - 2.) + 2.)) +
            D * (B - A * D) - 2 * Cc) / (D * D * D);
}
__device__ __forceinline__ double curgod_seg4(double nd0, double nd1, double v0, double v1,
                                              double f0, double f1, double dx) {       // :86-94
    const double A = nd0 * v0 * f0;
    const double cc = (v1 - v0) / dx;
    const double B = nd0 * f0 * cc;
    const double fu = nd1 * f1 / (nd0 * f0);
    const double D = log(fu) / dx;
    return (A * D * (fu - 1.) + B * fu * (D * dx - 1.) + B) / (D * D);
}

}  // namespace srdev
