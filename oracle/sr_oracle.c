/*
 * sr_oracle.c -- CPU restatement of the SpectRobot line-by-line hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (spectrobot_b200/) may import,
 * link or call this file; it is used by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py, as the checker and as the timed CPU baseline.
 *
 * The reference cannot be built in this environment (no Fortran compiler, no meson for f2py,
 * Python-2 sources, spect_base_module missing), so every routine here follows the
 * reference source line by line and cites it.  PARITY PINNING: the
 * reference ships no golden vectors, fixtures or tests.  Everything
 * the reference does in PYTHON around these routines - window placement, widths, G coefficients,
 * level selection, clipping and the staging matrix, the LUT it builds and pickles, the Lagrange
 * rule of the partition sums, LutSet.calculate, make_abscoeff_LUTS_fast, the convolution - is
 * pinned by fixtures produced by EXECUTING the reference's own spect_classes.py /
 * spect_main_module.py (tests/golden/ref_exec.py, tests/test_ref_golden*.py), with this file
 * standing in for the f2py modules.  The Fortran itself cannot be compiled (no Fortran compiler
 * exists here); humliv_bb, humli_bb, sum_all_lines and curgod_fort_* below are BIT-IDENTICAL to
 * the reference's own Fortran source executed statement by statement by the mechanical FORTRAN 77
 * executor tests/golden/f77_exec.py (fixtures tests/golden/f77_golden.npz, all three humliv_bb
 * branches; tests/test_f77_golden.py), under gfortran / x86-64 semantics (no FMA contraction).
 * The TIPS tables (generated from fparts_mod.f by tools/gen_tables.py) equal what bd_tips_2003
 * -> QT_* return when executed the same way, for all 108 (mol, iso).  The identities the
 * reference states run on top (tests/test_oracle.py).  The LOS
 * integral (orc_los_*) restates OUR OWN published spec because the
 * reference's sbm.LineOfSight.radtran_fast is not in the tree: for that part "parity unpinned".
 *
 * Build: see oracle/Makefile (plain gcc, -ffp-contract=off so no FMA contraction sneaks in).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IMXSIG 13010 /* parameters.inc:65 */

/* ------------------------------------------------------------------------------------------
 * complex*16 helpers with gfortran semantics (-fcx-fortran-rules, the gfortran default):
 * plain multiplication; division by Smith's range-reduced algorithm.
 * ---------------------------------------------------------------------------------------- */
typedef struct { double re, im; } cplx;

static inline cplx c_mul(cplx a, cplx b) {
    cplx r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re };
    return r;
}
static inline cplx c_add_r(double r, cplx a) { cplx o = { r + a.re, a.im }; return o; }
static inline cplx c_r_sub(double r, cplx a) { cplx o = { r - a.re, -a.im }; return o; }
static inline cplx c_scale(cplx a, double r) { cplx o = { a.re * r, a.im * r }; return o; }
static inline cplx c_div(cplx a, cplx b) {
    cplx o;
    if (fabs(b.re) < fabs(b.im)) {
        double ratio = b.re / b.im;
        double div = b.re * ratio + b.im;
        o.re = (a.re * ratio + a.im) / div;
        o.im = (a.im * ratio - a.re) / div;
    } else {
        double ratio = b.im / b.re;
        double div = b.im * ratio + b.re;
        o.re = (a.im * ratio + a.re) / div;
        o.im = (a.im - a.re * ratio) / div;
    }
    return o;
}
/* Fortran CMPLX(a,b) without KIND: default (single precision) complex, then stored into a
 * complex*16 -> both parts rounded to float32 (lineshape.f:164,278,364,529). */
static inline cplx cmplx_default(double a, double b) {
    cplx o = { (double)(float)a, (double)(float)b };
    return o;
}
static inline long f_nint(double x) { return lround(x); } /* NINT: half away from zero */
static inline long lmax(long a, long b) { return a > b ? a : b; }
static inline long lmin(long a, long b) { return a < b ? a : b; }

/* Regions 3/4 of humliv_bb: coefficients are un-suffixed real*4 literals (lineshape.f:281-287,
 * 299-303, 367-373, 391-395, 539-545, 554-558) widened to double. */
#define F(x) ((double)x##f)
static inline double humliv_core(double rx, double ry) {
    double r2 = 0.195 * rx - 0.176;         /* lineshape.f:528 (D0 literals) */
    cplx c2 = cmplx_default(ry, -rx);       /* lineshape.f:529 */
    if (ry < r2) {                          /* region 4, lineshape.f:530-546 */
        cplx c1 = c_mul(c2, c2);
        cplx num = c_scale(c1, F(.56419));
        num = c_r_sub(F(1.320522), num);
        num = c_r_sub(F(35.76683), c_mul(c1, num));
        num = c_r_sub(F(219.0313), c_mul(c1, num));
        num = c_r_sub(F(1540.787), c_mul(c1, num));
        num = c_r_sub(F(3321.9905), c_mul(c1, num));
        num = c_r_sub(F(36183.31), c_mul(c1, num));
        num = c_mul(c2, num);
        cplx den = c_r_sub(F(1.841439), c1);
        den = c_r_sub(F(61.57037), c_mul(c1, den));
        den = c_r_sub(F(364.2191), c_mul(c1, den));
        den = c_r_sub(F(2186.181), c_mul(c1, den));
        den = c_r_sub(F(9022.228), c_mul(c1, den));
        den = c_r_sub(F(24322.84), c_mul(c1, den));
        den = c_r_sub(F(32066.6), c_mul(c1, den));
        cplx c3 = c_div(num, den);
        return exp(c1.re) * cos(c1.im) - c3.re;
    } else {                                /* region 3, lineshape.f:548-560 */
        cplx num = c_scale(c2, F(.5642236));
        num = c_add_r(F(3.778987), num);
        num = c_add_r(F(11.96482), c_mul(c2, num));
        num = c_add_r(F(20.20933), c_mul(c2, num));
        num = c_add_r(F(16.4955), c_mul(c2, num));
        cplx den = c_add_r(F(6.699398), c2);
        den = c_add_r(F(21.69274), c_mul(c2, den));
        den = c_add_r(F(39.27121), c_mul(c2, den));
        den = c_add_r(F(38.82363), c_mul(c2, den));
        den = c_add_r(F(16.4955), c_mul(c2, den));
        cplx c3 = c_div(num, den);
        return c3.re;
    }
}
#undef F

/* region-2 / region-1 real coefficient sets (lineshape.f:492-502 and 456-460) */
typedef struct { double a, b, c, d, e, f, g, h; } reg2_t;
static inline reg2_t reg2_coef(double ry, double ry2) {
    reg2_t k;
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
static inline double reg2_eval(const reg2_t* k, double x2) {
    return (k->a + x2 * (k->b + x2 * (k->c + k->d * x2))) /
           (k->e + x2 * (k->f + x2 * (k->g + x2 * (k->h + x2))));
}
typedef struct { double a, b, c, d, e; } reg1_t;
static inline reg1_t reg1_coef(double ry, double ry2) {
    reg1_t k;
    k.a = ry * (1.1283792 + 2.2567584 * ry2);
    k.b = 2.2567584 * ry;
    k.c = (1. + 2. * ry2) * (1. + 2. * ry2);
    k.d = -4. + 8. * ry2;
    k.e = 4.;
    return k;
}
static inline double reg1_eval(const reg1_t* k, double x2) {
    return (k->a + x2 * k->b) / (k->c + x2 * (k->d + k->e * x2));
}

/* ------------------------------------------------------------------------------------------
 * humliv_bb -- lineshape.f:226-569.  x has n >= i2 elements, i1/i2 are 1-based inclusive.
 * Returns 0, or 1 / 2 where the Fortran executes STOP (lineshape.f:255, 263).
 * y is written only on [i1,i2] (and only where the Fortran writes).
 * ---------------------------------------------------------------------------------------- */
int orc_humliv_bb(const double* x_, int n, int i1, int i2, double x0, double lw, double dw,
                  double* y_) {
    (void)n;
    const double* x = x_ - 1; /* 1-based views */
    double* y = y_ - 1;
    long j, k, l, ir, ir2, il, il2;
    double rx, ry, drun, dstep, tst, xrun, xstep, x2, ry2;

    if (i1 > i2) return 1;                       /* :253-256 */
    if (dw > 0.0) ry = lw / dw; else return 2;   /* :260-264 */
    dstep = (x[i1 + 1] - x[i1]) / dw;            /* :265 */
    xstep = dstep;
    ry2 = ry * ry;

    if (x0 <= x[i1]) {                           /* forward loop, :272-357 */
        tst = 5.5;
        j = i1;
        rx = (x[j] - x0) / dw;
        while ((rx + ry < tst) && (j <= i2)) {
            y[j] = humliv_core(rx, ry);          /* same formulas as :277-312 */
            j = j + 1;
            rx = rx + xstep;
        }
        if (j <= i2) {
            tst = 15.0;
            l = lmax(f_nint((tst - ry - rx) / xstep), 0) + j;
            l = lmin(l, i2);
            if (l > j) {
                reg2_t c = reg2_coef(ry, ry2);
                drun = (x[j] - x0) / dw;
                xrun = drun;
                for (k = j; k <= l; k++) {
                    x2 = xrun * xrun;
                    y[k] = reg2_eval(&c, x2);
                    xrun = xrun + xstep;
                }
                l = l + 1;
            }
            if (l < j) l = j;
            if (l < i2) {
                reg1_t c = reg1_coef(ry, ry2);
                drun = (x[l] - x0) / dw;
                xrun = drun;
                for (k = l; k <= i2; k++) {
                    x2 = xrun * xrun;
                    y[k] = reg1_eval(&c, x2);
                    xrun = xrun + xstep;
                }
            }
        }
    } else if (x0 >= x[i2]) {                    /* backward loop, :358-442 */
        tst = 5.5;
        j = i2;
        rx = (x0 - x[j]) / dw;
        while ((rx + ry < tst) && (j >= i1)) {
            y[j] = humliv_core(rx, ry);
            j = j - 1;
            rx = rx + xstep;
        }
        if (j >= i1) {
            tst = 15.0;
            l = j - lmax(f_nint((tst - ry - rx) / dw / xstep), 0); /* sic: extra /dw, :404 */
            l = lmax(l, i1);
            if (l == i2) l = i2 + 1;
            if (l < j) {
                reg2_t c = reg2_coef(ry, ry2);
                drun = (x0 - x[l]) / dw;
                xrun = drun;
                for (k = l; k <= j; k++) {
                    x2 = xrun * xrun;
                    y[k] = reg2_eval(&c, x2);
                    xrun = xrun - xstep;
                }
            }
            if (l >= i1) {
                reg1_t c = reg1_coef(ry, ry2);
                drun = (x0 - x[i1]) / dw;
                xrun = drun;
                for (k = i1; k <= l - 1; k++) {
                    x2 = xrun * xrun;
                    y[k] = reg1_eval(&c, x2);
                    xrun = xrun - xstep;
                }
            }
        }
    } else {                                     /* x(i1) < x0 < x(i2), :443-562 */
        rx = (x0 - x[i1]) / dw;
        tst = 15.0;
        il = i1;
        if (rx + ry >= tst) il = lmax(f_nint((rx - ry - tst) / xstep), 0) + i1;
        rx = (x[i2] - x0) / dw;
        ir = i2;
        if (rx + ry >= tst) ir = i2 - lmax(f_nint((rx - ry - tst) / xstep), 0);
        if (il > i1 || ir < i2) {
            reg1_t c = reg1_coef(ry, ry2);
            if (il > i1) {
                drun = (x0 - x[i1]) / dw;
                xrun = drun;
                for (k = i1; k <= il; k++) {
                    x2 = xrun * xrun;
                    y[k] = reg1_eval(&c, x2);
                    xrun = xrun - xstep;
                }
            }
            if (ir < i2) {
                drun = (x[ir] - x0) / dw;
                xrun = drun;
                for (k = ir; k <= i2; k++) {
                    x2 = xrun * xrun;
                    y[k] = reg1_eval(&c, x2);
                    xrun = xrun + xstep;
                }
            }
        }
        rx = (x0 - x[il]) / dw;
        tst = 5.5;
        il2 = il;
        if (rx + ry >= tst) il2 = il + lmax(f_nint((rx - ry - tst) / xstep), 0);
        ir2 = ir;
        rx = (x[ir] - x0) / dw;
        if (rx + ry >= tst) ir2 = ir - lmax(f_nint((rx - ry - tst) / xstep), 0);
        if (il2 > il || ir2 < ir) {
            reg2_t c = reg2_coef(ry, ry2);
            if (il < il2) {
                drun = (x0 - x[il]) / dw;
                xrun = drun;
                for (j = il; j <= il2; j++) {
                    x2 = xrun * xrun;
                    y[j] = reg2_eval(&c, x2);
                    xrun = xrun - xstep;
                }
            }
            if (ir2 < ir) {
                drun = (x[ir2] - x0) / dw;
                xrun = drun;
                for (j = ir2; j <= ir; j++) {
                    x2 = xrun * xrun;
                    y[j] = reg2_eval(&c, x2);
                    xrun = xrun + xstep;
                }
            }
        }
        if (il2 == il) il2 = il - 1;
        if (ir2 == ir) ir2 = ir + 1;
        for (j = il2 + 1; j <= ir2 - 1; j++) {
            rx = fabs(x[j] - x0) / dw;
            y[j] = humliv_core(rx, ry);
        }
    }
    return 0;
}

/* region boundaries of the "x0 inside" branch, for tests (1-based il, ir, il2, ir2 as left by
 * lineshape.f:443-490 BEFORE the :524-525 adjustment). */
int orc_humliv_regions(const double* x_, int i1, int i2, double x0, double lw, double dw,
                       int* out4) {
    const double* x = x_ - 1;
    if (i1 > i2 || !(dw > 0.0)) return 1;
    if (x0 <= x[i1] || x0 >= x[i2]) return 2;
    double ry = lw / dw, xstep = (x[i1 + 1] - x[i1]) / dw, rx;
    long il = i1, ir = i2, il2, ir2;
    rx = (x0 - x[i1]) / dw;
    if (rx + ry >= 15.0) il = lmax(f_nint((rx - ry - 15.0) / xstep), 0) + i1;
    rx = (x[i2] - x0) / dw;
    if (rx + ry >= 15.0) ir = i2 - lmax(f_nint((rx - ry - 15.0) / xstep), 0);
    rx = (x0 - x[il]) / dw;
    il2 = il;
    if (rx + ry >= 5.5) il2 = il + lmax(f_nint((rx - ry - 5.5) / xstep), 0);
    rx = (x[ir] - x0) / dw;
    ir2 = ir;
    if (rx + ry >= 5.5) ir2 = ir - lmax(f_nint((rx - ry - 5.5) / xstep), 0);
    out4[0] = (int)il; out4[1] = (int)ir; out4[2] = (int)il2; out4[3] = (int)ir2;
    return 0;
}

/* humli_bb -- lineshape.f:150-205 (scalar routine, D0 coefficients; cross-check only). */
double orc_humli_bb(double rx, double ry) {
    double r1 = fabs(rx) + ry;
    double r2 = 0.195 * fabs(rx) - 0.176;
    cplx c2 = cmplx_default(ry, -rx), c1, c3;
    if (r1 >= 15.0) {
        cplx den = c_add_r(0.5, c_mul(c2, c2));
        c3 = c_div(c_scale(c2, 0.5641896), den);
        return c3.re;
    } else if (r1 >= 5.5) {
        c1 = c_mul(c2, c2);
        cplx num = c_mul(c2, c_add_r(1.410474, c_scale(c1, .5641896)));
        cplx den = c_add_r(.75, c_mul(c1, c_add_r(3.0, c1)));
        c3 = c_div(num, den);
        return c3.re;
    } else if (ry >= r2) {
        cplx num = c_scale(c2, .5642236);
        num = c_add_r(3.778987, num);
        num = c_add_r(11.96482, c_mul(c2, num));
        num = c_add_r(20.20933, c_mul(c2, num));
        num = c_add_r(16.4955, c_mul(c2, num));
        cplx den = c_add_r(6.699398, c2);
        den = c_add_r(21.69274, c_mul(c2, den));
        den = c_add_r(39.27121, c_mul(c2, den));
        den = c_add_r(38.82363, c_mul(c2, den));
        den = c_add_r(16.4955, c_mul(c2, den));
        c3 = c_div(num, den);
        return c3.re;
    } else {
        c1 = c_mul(c2, c2);
        cplx num = c_scale(c1, .56419);
        num = c_r_sub(1.320522, num);
        num = c_r_sub(35.76683, c_mul(c1, num));
        num = c_r_sub(219.0313, c_mul(c1, num));
        num = c_r_sub(1540.787, c_mul(c1, num));
        num = c_r_sub(3321.9905, c_mul(c1, num));
        num = c_r_sub(36183.31, c_mul(c1, num));
        num = c_mul(c2, num);
        cplx den = c_r_sub(1.841439, c1);
        den = c_r_sub(61.57037, c_mul(c1, den));
        den = c_r_sub(364.2191, c_mul(c1, den));
        den = c_r_sub(2186.181, c_mul(c1, den));
        den = c_r_sub(9022.228, c_mul(c1, den));
        den = c_r_sub(24322.84, c_mul(c1, den));
        den = c_r_sub(32066.6, c_mul(c1, den));
        c3 = c_div(num, den);
        return exp(c1.re) * cos(c1.im) - c3.re;
    }
}

/* sum_all_lines -- lineshape.f:2-25.  matrix is column-major (Fortran) with leading dimension
 * ld (= imxlines in the reference): element (ilin, i) at matrix[(i-1)*ld + (ilin-1)].
 * init/fin are 1-based inclusive.  n_spe = length of spe_ini/spe_fin. */
void orc_sum_all_lines(const double* spe_ini, const double* matrix, const int* init,
                       const int* fin, int n_lines, int ld, int n_spe, double* spe_fin) {
    memcpy(spe_fin, spe_ini, (size_t)n_spe * sizeof(double));
    for (int ilin = 0; ilin < n_lines; ilin++) {
        size_t i = 0;
        for (int j = init[ilin]; j <= fin[ilin]; j++) {
            spe_fin[j - 1] += matrix[i * (size_t)ld + (size_t)ilin];
            i++;
        }
    }
}

/* curgod_fort_1..4 -- curgods.f:2-98 (no guard for nd(i+1)==nd(i), like the reference). */
double orc_curgod_1(const double* nd, const double* x, int n_p) {
    double res = 0.0;
    for (int i = 0; i < n_p - 1; i++) {
        double dx = x[i + 1] - x[i];
        double fu = nd[i + 1] / nd[i];
        double D = log(fu) / dx;
        res = res + (nd[i + 1] - nd[i]) / D;
    }
    return res;
}
double orc_curgod_2(const double* nd, const double* vmr, const double* x, int n_p) {
    double res = 0.0;
    for (int i = 0; i < n_p - 1; i++) {
        double dx = x[i + 1] - x[i];
        double A = nd[i] * vmr[i];
        double B = nd[i] * (vmr[i + 1] - vmr[i]) / dx;
        double fu = nd[i + 1] / nd[i];
        double D = log(fu) / dx;
        res = res + (A * D * (fu - 1.) + B * fu * (D * dx - 1.) + B) / (D * D);
    }
    return res;
}
double orc_curgod_3(const double* nd, const double* vmr, const double* f, const double* x,
                    int n_p) {
    double res = 0.0;
    for (int i = 0; i < n_p - 1; i++) {
        double dx = x[i + 1] - x[i];
        double A = nd[i] * vmr[i] * f[i];
        double cc = (vmr[i + 1] - vmr[i]) / dx;
        double bb = (f[i + 1] - f[i]) / dx;
        double B = nd[i] * (vmr[i] * bb + f[i] * cc);
        double C = nd[i] * bb * cc;
        double fu = nd[i + 1] / nd[i];
        double D = log(fu) / dx;
        res = res + (fu * (D * (A * D + B * (D * dx - 1.)) + C * (D * dx * (D * dx - 2.) + 2.)) +
                     D * (B - A * D) - 2 * C) / (D * D * D);
    }
    return res;
}
double orc_curgod_4(const double* nd, const double* vmr, const double* f, const double* x,
                    int n_p) {
    double res = 0.0;
    for (int i = 0; i < n_p - 1; i++) {
        double dx = x[i + 1] - x[i];
        double A = nd[i] * vmr[i] * f[i];
        double cc = (vmr[i + 1] - vmr[i]) / dx;
        double B = nd[i] * f[i] * cc;
        double fu = nd[i + 1] * f[i + 1] / (nd[i] * f[i]);
        double D = log(fu) / dx;
        res = res + (A * D * (fu - 1.) + B * fu * (D * dx - 1.) + B) / (D * D);
    }
    return res;
}

/* bd_tips_2003 -- fparts_mod.f:33-295: (gi, T grid 60..3010 step 25, Q[119]).  Unknown
 * molecule / isotopologue -> returns 1 (the Fortran leaves the outputs uninitialised). */
#include "../spectrobot_b200/csrc/tips2003_tables.inc"
int orc_bd_tips_2003(int mol, int iso, double* gi, double* t119, double* q119) {
    for (int i = 0; i < SR_TIPS_NT; i++) t119[i] = 60.0 + 25.0 * i; /* fparts_mod.f:58-76 */
    for (int m = 0; m < SR_TIPS_NMOL; m++) {
        if (sr_tips_index[m][0] != mol) continue;
        if (iso < 1 || iso > sr_tips_index[m][2]) return 1;
        int row = sr_tips_index[m][1] + iso - 1;
        *gi = sr_tips_gj[row];
        for (int i = 0; i < SR_TIPS_NT; i++) {
            union { unsigned int u; float f; } cv;
            cv.u = sr_tips_qbits[row][i];
            q119[i] = (double)cv.f;
        }
        return 0;
    }
    return 1;
}

/* CalcPartitionSum -- spect_classes.py:1692-1710: Lagrange polynomial through the (up to) two
 * nodes <= T and the (up to) two nodes > T.  scipy.interpolate.lagrange builds the monomial
 * coefficients and evaluates them with Horner; this is the same polynomial evaluated in the
 * Lagrange basis (difference ~1e-13 relative; the NumPy oracle in oracle/ref_py.py calls scipy
 * itself and is the tighter checker for this function). */
double orc_partition_sum(int mol, int iso, double temp) {
    double gi, t[SR_TIPS_NT], q[SR_TIPS_NT];
    if (orc_bd_tips_2003(mol, iso, &gi, t, q)) return NAN;
    int nle = 0;
    while (nle < SR_TIPS_NT && t[nle] <= temp) nle++;
    int lo = nle - 2 < 0 ? 0 : nle - 2;
    int hi = nle + 2 > SR_TIPS_NT ? SR_TIPS_NT : nle + 2;
    double acc = 0.0;
    for (int a = lo; a < hi; a++) {
        double w = q[a];
        for (int b = lo; b < hi; b++)
            if (b != a) w *= (temp - t[b]) / (t[a] - t[b]);
        acc += w;
    }
    return acc;
}

/* ------------------------------------------------------------------------------------------
 * Per-line physics (spect_classes.py): widths and G coefficients.
 * consts = {h_cgs, c_cgs, k_cgs, N_Avogadro} as the host reads them from scipy
 * (spect_classes.py:44-47, 1984).
 * ---------------------------------------------------------------------------------------- */
#define T_REF 296.0                      /* spect_classes.py:39 */
#define HPA_TO_ATM 0.00098692326671601   /* spect_classes.py:40 */

void orc_widths(double freq, double air_broad, double t_dep, double temp, double pres_hpa,
                double mm, const double* consts, double* lw, double* dw) {
    double pres_atm = pres_hpa * HPA_TO_ATM;                 /* :2034 */
    /* Lorenz_width :1972 with Self_broad = Self_pres_atm = 0 */
    *lw = pow(T_REF / temp, t_dep) * (air_broad * (pres_atm - 0.0) + 0.0 * 0.0);
    /* Doppler_width :1984 */
    *dw = freq / consts[1] * sqrt(2 * consts[3] * consts[2] * temp * log(2.0) / mm);
}

/* Calc_Gcoeffs :312-343 with Einstein_A_to_Gcoeff_{spem,indem,abs} :1806-1853.
 * g[0]=sp_emission, g[1]=ind_emission, g[2]=absorption. */
void orc_gcoeffs(double freq, double a_coeff, double e_lower, double g_up, double g_lo,
                 double e_vib_up, double e_vib_lo, double temp, const double* consts,
                 double* g) {
    if (!(a_coeff != 0.0 && g_lo != 0.0 && g_up != 0.0)) { g[0] = g[1] = g[2] = 0.0; return; }
    double h = consts[0], c = consts[1], k = consts[2];
    double c2 = h * c / k;                                      /* :47 */
    double fact_2 = 2 * h * pow(c, 2.0) * pow(freq, 3.0);       /* :1743 */
    double b21 = a_coeff / fact_2;                              /* :1750 */
    double rot_up = g_up * exp(-c2 * (e_lower + freq - e_vib_up) / temp); /* :1850, :1877 */
    g[0] = h * c * freq * rot_up * a_coeff / (4 * M_PI);        /* :1851 */
    g[1] = h * c * freq * rot_up * b21 / (4 * M_PI);            /* :1840 */
    double b12 = b21 * g_up / g_lo;                             /* :1783, :1814 */
    double rot_lo = g_lo * exp(-c2 * (e_lower - e_vib_lo) / temp);        /* :1815 */
    g[2] = h * c * freq * rot_lo * b12 / (4 * M_PI);            /* :1817 */
}

/* closest_grid :1937-1943: argmin |grid - wn| (ties -> lowest index) on an ascending grid. */
long orc_closest_grid(const double* grid, long n, double wn) {
    long lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        long mid = (lo + hi) / 2;
        if (grid[mid] <= wn) lo = mid; else hi = mid;
    }
    long best = lo;
    double bd = fabs(grid[lo] - wn);
    for (long c = lo - 1; c <= hi + 1; c++) {
        if (c < 0 || c >= n) continue;
        double d = fabs(grid[c] - wn);
        if (d < bd || (d == bd && c < best)) { bd = d; best = c; }
    }
    return best;
}

/* One line's normalised shape on its own window: PrepareCalcShapes :1453-1457 ->
 * MakeShapeLine :174-206 -> MakeShape :1990-2008.  lin_grid = window offsets L[13010].
 * shape[j] = humliv_bb(L + grid[ind], 1, 13010, freq, lw, dw/sqrt(ln2))[j] / (dw*sqrt(pi/ln2)) */
int orc_line_shape(double freq, double lw, double dw, double centre, const double* lin_grid,
                   double* xbuf, double* shape) {
    for (int j = 0; j < IMXSIG; j++) xbuf[j] = lin_grid[j] + centre;       /* :1455 */
    double fac = dw * sqrt(M_PI / log(2.0));                               /* :1997 */
    int rc = orc_humliv_bb(xbuf, IMXSIG, 1, IMXSIG, freq, lw, dw / sqrt(log(2.0)), shape);
    if (rc) return rc;
    for (int j = 0; j < IMXSIG; j++) shape[j] = 1.0 * shape[j] / fac;      /* :2003 */
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * One (P,T) LUT cell: LookUpTable.make loop body (spect_main_module.py:753-774) =
 * calc_shapes_lines (spect_classes.py:1378) + LutSet.add_PT per set (smm:1122) ->
 * BuildCoeff(preCalc_shapes=True) (spcl:1277) -> add_lines_to_spectrum (spcl:1016) ->
 * sum_all_lines (lineshape.f:2).
 *
 * Lines are SoA arrays.  up_set/lo_set give the output set (level) index a line feeds as upper
 * (sp_/ind_emission) and lower (absorption) level; a negative index on EITHER drops the line,
 * which is LinkToMolec's filter (spcl:1384-1388).  For an LTE isotopologue (single set 'all',
 * smm:742-748) pass n_sets=1, up_set=lo_set=0 and e_vib_*=0 (spcl:318-321, 1316-1319).
 * out[n_sets][3][n_grid], ctype order sp_emission, ind_emission, absorption; out is ZEROED here.
 * Summation order: lines in input order at every grid point, as lineshape.f:17-23, for any
 * n_threads (threads own disjoint spectrum slabs, see below), so the result does not depend
 * on the thread count.
 * Returns 0, else the humliv_bb status of the first failing line.
 * ---------------------------------------------------------------------------------------- */
int orc_gcoeff_cell(int n_lines, const double* freq, const double* a_coeff,
                    const double* air_broad, const double* t_dep, const double* e_lower,
                    const double* g_up, const double* g_lo, const double* e_vib_up,
                    const double* e_vib_lo, const int* up_set, const int* lo_set,
                    const double* grid, long n_grid, const double* lin_grid, double temp,
                    double pres_hpa, double mm, const double* consts, int n_sets, double* out,
                    int n_threads) {
    memset(out, 0, (size_t)n_sets * 3 * (size_t)n_grid * sizeof(double));
    int status = 0;
    if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
    omp_set_num_threads(n_threads);
#endif
    /* The spectrum is cut into n_threads contiguous slabs; every thread evaluates the lines whose
     * window touches its slab and adds only the part inside the slab.  Lines near a slab border
     * are evaluated by both neighbours (that costs a little extra work but keeps the addition
     * order = input line order at every point, i.e. bit-identical to the sequential loop). */
#pragma omp parallel
    {
        double* xbuf = (double*)malloc(IMXSIG * sizeof(double));
        double* shape = (double*)malloc(IMXSIG * sizeof(double));
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int tid = 0, nt = 1;
#endif
        long s0 = n_grid * tid / nt, s1 = n_grid * (tid + 1) / nt;
        for (int l = 0; l < n_lines; l++) {
            if (up_set[l] < 0 || lo_set[l] < 0) continue;
            long ind = orc_closest_grid(grid, n_grid, freq[l]);          /* :1454 */
            long w0 = ind - IMXSIG / 2, w1 = w0 + IMXSIG;                /* window [w0,w1) */
            if (w1 <= s0 || w0 >= s1) continue;
            double lw, dw, g[3];
            orc_widths(freq[l], air_broad[l], t_dep[l], temp, pres_hpa, mm, consts, &lw, &dw);
            int rc = orc_line_shape(freq[l], lw, dw, grid[ind], lin_grid, xbuf, shape);
            if (rc) {
#pragma omp critical
                if (!status) status = rc;
                continue;
            }
            orc_gcoeffs(freq[l], a_coeff[l], e_lower[l], g_up[l], g_lo[l], e_vib_up[l],
                        e_vib_lo[l], temp, consts, g);
            long a = w0 > s0 ? w0 : s0, b = w1 < s1 ? w1 : s1;
            for (int ct = 0; ct < 3; ct++) {
                int set = ct < 2 ? up_set[l] : lo_set[l];
                double* dst = out + ((size_t)set * 3 + ct) * (size_t)n_grid;
                double gg = g[ct];
                for (long s = a; s < b; s++) dst[s] += gg * shape[s - w0]; /* spcl:688, lineshape.f:20 */
            }
        }
        free(xbuf);
        free(shape);
    }
    return status;
}

/* ==========================================================================================
 * LOS radiative transfer (subsystem 3).  PARITY UNPINNED against the original author's code:
 * sbm.LineOfSight.radtran_fast is not in the reference tree.  What IS restated from
 * the reference: LutSet.calculate (spect_main_module.py:997-1066), SpectralGcoeff.interpolate
 * (spect_classes.py:1349-1375), make_abscoeff_LUTS_fast (spect_main_module.py:2134-2299),
 * the float32 LUT (spect_main_module.py:1676, spect_classes.py:732).  The layer recursion and
 * sbm.weight follow DESIGN.md section 6.
 * ======================================================================================== */

/* sbm.weight(v, v1, v2, 'lin') -- DESIGN.md 6.2: w1 = (v2-v)/(v2-v1), w2 = (v-v1)/(v2-v1) */
static void orc_weight(double v, double v1, double v2, double* w1, double* w2) {
    *w1 = (v2 - v) / (v2 - v1);
    *w2 = (v - v1) / (v2 - v1);
}

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}
static int unique_sorted(const double* v, int n, int stride, double* out) {
    for (int i = 0; i < n; i++) out[i] = v[i * stride];
    qsort(out, n, sizeof(double), cmp_double);
    int m = 0;
    for (int i = 0; i < n; i++)
        if (m == 0 || out[i] != out[m - 1]) out[m++] = out[i];
    return m;
}
/* index of the nearest and second nearest node: np.argmin / np.argsort(...)[1] with a stable
 * order (ties -> lower index), spect_main_module.py:1010-1011, 1027-1028, 1034-1035 */
static void nearest_two(const double* nodes, int n, double v, int* i1, int* i2) {
    int a = 0;
    for (int i = 1; i < n; i++)
        if (fabs(nodes[i] - v) < fabs(nodes[a] - v)) a = i;
    int b = -1;
    for (int i = 0; i < n; i++) {
        if (i == a) continue;
        if (b < 0 || fabs(nodes[i] - v) < fabs(nodes[b] - v)) b = i;
    }
    *i1 = a;
    *i2 = b;
}
static int find_cell(const double* pt, int n_cells, double p, double t) {   /* LutSet.find :985 */
    for (int i = 0; i < n_cells; i++)
        if (pt[2 * i] == p && pt[2 * i + 1] == t) return i;
    return -1;
}

/* LutSet.calculate :997-1066 as (cell, weight) pairs: result = sum_i w[i] * set[cell[i]].
 * Order: (P1,T1), (P1,T2), (P2,T1), (P2,T2); in the P <= min(Ps) branch only the first two
 * (cells 2,3 = -1, weights 0).  Returns 0, 6 = "couple not found", 7 = "Extrapolating in P". */
int orc_lut_weights(const double* pt, int n_cells, double pres, double temp, int* cell,
                    double* w) {
    double* Ps = (double*)malloc(sizeof(double) * n_cells);
    double* Ts = (double*)malloc(sizeof(double) * n_cells);
    int nP = unique_sorted(pt, n_cells, 2, Ps), nT = unique_sorted(pt + 1, n_cells, 2, Ts);
    int rc = 0;
    cell[0] = cell[1] = cell[2] = cell[3] = -1;
    w[0] = w[1] = w[2] = w[3] = 0.0;
    if (nT < 2) { rc = 6; goto done; }
    if (pres <= Ps[0]) {                                         /* :1007-1025 */
        int ia, ib;
        nearest_two(Ts, nT, temp, &ia, &ib);
        cell[0] = find_cell(pt, n_cells, Ps[0], Ts[ia]);
        cell[1] = find_cell(pt, n_cells, Ps[0], Ts[ib]);
        if (cell[0] < 0 || cell[1] < 0) { rc = 6; goto done; }
        orc_weight(temp, Ts[ia], Ts[ib], &w[0], &w[1]);
    } else if (pres <= Ps[nP - 1]) {                             /* :1026-1056 */
        int p1, p2, t1, t2;
        if (nP < 2) { rc = 6; goto done; }
        nearest_two(Ps, nP, pres, &p1, &p2);
        nearest_two(Ts, nT, temp, &t1, &t2);
        cell[0] = find_cell(pt, n_cells, Ps[p1], Ts[t1]);
        cell[1] = find_cell(pt, n_cells, Ps[p1], Ts[t2]);
        cell[2] = find_cell(pt, n_cells, Ps[p2], Ts[t1]);
        cell[3] = find_cell(pt, n_cells, Ps[p2], Ts[t2]);
        if (cell[0] < 0 || cell[1] < 0 || cell[2] < 0 || cell[3] < 0) { rc = 6; goto done; }
        double wp1, wp2, wt1, wt2;
        orc_weight(pres, Ps[p1], Ps[p2], &wp1, &wp2);            /* P first, :1054-1055 */
        orc_weight(temp, Ts[t1], Ts[t2], &wt1, &wt2);            /* then T, :1056 */
        w[0] = wt1 * wp1;   /* (P1,T1) */
        w[1] = wt2 * wp1;   /* (P1,T2) */
        w[2] = wt1 * wp2;   /* (P2,T1) */
        w[3] = wt2 * wp2;   /* (P2,T2) */
    } else {
        rc = 7;                                                  /* :1058 */
    }
done:
    free(Ps);
    free(Ts);
    return rc;
}

/* One isotopologue's compressed LUT as the oracle sees it */
typedef struct {
    const float* g32;        /* [n_cells][n_sets][3][n_grid] */
    const double* pt;        /* [n_cells][2] */
    const double* level_energy; /* [n_sets] */
    int n_cells, n_sets, mol, iso, lte_unidentified;
    double iso_ratio;
} orc_lut;

/* abs / emission coefficient of one gas at one step for grid points [pt0, pt0+n_pts):
 * make_abscoeff_LUTS_fast :2200-2250 with LutSet.calculate + interpolate done literally
 * (w1*spec1 + w2*spec2 in the reference's order). tvib: [n_sets] or NULL (LTE, :2231-2232). */
static int orc_abs_emi(const orc_lut* L, long n_grid, double pres, double temp,
                       const double* tvib, const double* consts, long pt0, long n_pts,
                       double* abs_c, double* emi_c) {
    int cell[4];
    double w[4];
    int rc = orc_lut_weights(L->pt, L->n_cells, pres, temp, cell, w);
    if (rc) return rc;
    double c2 = consts[0] * consts[1] / consts[2];
    double q_part = orc_partition_sum(L->mol, L->iso, temp);             /* :2212 */
    for (long i = 0; i < n_pts; i++) abs_c[i] = emi_c[i] = 0.0;
    int two = (cell[2] < 0);
    /* recover the separate P and T weights for the literal two-stage interpolation */
    for (int s = 0; s < L->n_sets; s++) {
        double pop;
        if (L->lte_unidentified) pop = 1 / q_part;                       /* :2218 */
        else {
            double vibt = tvib ? tvib[s] : temp;                         /* :2231-2234 */
            pop = exp(-c2 * L->level_energy[s] / vibt) / q_part;         /* :2241 */
        }
        for (int ct = 0; ct < 3; ct++) {
            const float* base[4];
            for (int c = 0; c < 4; c++)
                base[c] = cell[c] < 0 ? NULL
                    : L->g32 + (((size_t)cell[c] * L->n_sets + s) * 3 + ct) * (size_t)n_grid + pt0;
            for (long i = 0; i < n_pts; i++) {
                double v;
                if (two) v = w[0] * (double)base[0][i] + w[1] * (double)base[1][i];
                else v = w[0] * (double)base[0][i] + w[2] * (double)base[2][i] +
                         (w[1] * (double)base[1][i] + w[3] * (double)base[3][i]);
                if (ct == 2) abs_c[i] += v * pop;                        /* :2245 */
                else if (ct == 1) abs_c[i] -= v * pop;                   /* :2247 */
                else emi_c[i] += v * pop;                                /* :2249 */
            }
        }
    }
    return 0;
}

/* DESIGN.md 6.4 layer update: I <- I exp(-tau) + J phi(tau), phi = -expm1(-tau)/tau (1 at 0) */
static inline double layer_update(double I, double tau, double J, int solo_absorption) {
    double em = expm1(-tau);
    double t = exp(-tau);
    if (solo_absorption) return I * t;
    double phi = (tau == 0.0) ? 1.0 : -em / tau;
    return I * t + J * phi;
}

/* Radiances of a LOS batch on points [pt0, pt0+n_pts).  Step tables as in sr_los_steps
 * (spectrobot.h).  luts: n_gas entries.  rad [n_los][n_pts]; tau_out/src_out (optional, may be
 * NULL): materialised [n_los][n_steps_max][n_pts] layer optical depths and source functions
 * S = J/tau (0 where tau == 0). */
int orc_los_rt(const orc_lut* luts, int n_gas, long n_grid, int n_los, int n_steps_max,
               int n_sets_max, const int* n_steps, const double* temp, const double* pres,
               const double* column, const double* tvib, const double* consts, long pt0,
               long n_pts, const double* i0, int solo_absorption, double* rad, double* tau_out,
               double* src_out, int n_threads) {
    int status = 0;
#ifdef _OPENMP
    omp_set_num_threads(n_threads < 1 ? 1 : n_threads);
#endif
#pragma omp parallel
    {
        double* tau = (double*)malloc(sizeof(double) * n_pts);
        double* J = (double*)malloc(sizeof(double) * n_pts);
        double* a = (double*)malloc(sizeof(double) * n_pts);
        double* e = (double*)malloc(sizeof(double) * n_pts);
        double* tv = (double*)malloc(sizeof(double) * (n_sets_max > 0 ? n_sets_max : 1));
#pragma omp for schedule(dynamic)
        for (int l = 0; l < n_los; l++) {
            double* I = rad + (size_t)l * n_pts;
            for (long i = 0; i < n_pts; i++) I[i] = i0 ? i0[(size_t)l * n_pts + i] : 0.0;
            for (int k = 0; k < n_steps[l]; k++) {
                size_t sk = (size_t)l * n_steps_max + k;
                for (long i = 0; i < n_pts; i++) tau[i] = J[i] = 0.0;
                for (int m = 0; m < n_gas; m++) {
                    const double* tvp = NULL;
                    if (tvib && !luts[m].lte_unidentified) {
                        for (int s = 0; s < luts[m].n_sets; s++)
                            tv[s] = tvib[(((size_t)m * n_sets_max + s) * n_los + l) * n_steps_max + k];
                        tvp = tv;
                    }
                    int rc = orc_abs_emi(&luts[m], n_grid, pres[sk], temp[sk], tvp, consts, pt0,
                                         n_pts, a, e);
                    if (rc) {
#pragma omp critical
                        if (!status) status = rc;
                        continue;
                    }
                    double col = luts[m].iso_ratio * column[((size_t)m * n_los + l) * n_steps_max + k];
                    for (long i = 0; i < n_pts; i++) { tau[i] += a[i] * col; J[i] += e[i] * col; }
                }
                for (long i = 0; i < n_pts; i++) {
                    I[i] = layer_update(I[i], tau[i], J[i], solo_absorption);
                    if (tau_out) {
                        size_t o = ((size_t)l * n_steps_max + k) * n_pts + i;
                        tau_out[o] = tau[i];
                        src_out[o] = (tau[i] == 0.0) ? 0.0 : J[i] / tau[i];
                    }
                }
            }
        }
        free(tau); free(J); free(a); free(e); free(tv);
    }
    return status;
}

/* K3 on materialised layers: I <- I e^-tau + S (1 - e^-tau)  (north_star's statement of the
 * integral). */
void orc_los_layers(const double* tau, const double* src, const int* n_steps, int n_los,
                    int n_steps_max, long n_pts, const double* i0, int solo_absorption,
                    double* rad) {
    for (int l = 0; l < n_los; l++)
        for (long i = 0; i < n_pts; i++) {
            double I = i0 ? i0[(size_t)l * n_pts + i] : 0.0;
            for (int k = 0; k < n_steps[l]; k++) {
                size_t o = ((size_t)l * n_steps_max + k) * n_pts + i;
                double em = expm1(-tau[o]), t = exp(-tau[o]);
                I = solo_absorption ? I * t : I * t + src[o] * (-em);
            }
            rad[(size_t)l * n_pts + i] = I;
        }
}

/* Analytic Jacobians of the layer recursion, written in the closed "sum over
 * layers" form rather than as the recursion the CUDA kernel runs, so that the two are independent:
 *   I_N = I_0 T(1..N) + sum_k J_k phi_k T(k+1..N),           T(a..b) = exp(-sum_{a..b} tau)
 *   dI_N/dp = sum_k f_kp [ (J_g,k phi_k + J_k phi'_k tau_g,k) T(k+1..N)
 *                          - tau_g,k ( I_0 T(1..N) + sum_{j<k} J_j phi_j T(j+1..N) ) ]
 * tau/emi: whole mixture, tau_g/emi_g: the retrieved gas alone (NULL = same arrays);
 * dfrac [n_los][n_steps_max][n_par]; jac [n_los][n_par][n_pts]. */
static double orc_phi(double t) { return (t == 0.0) ? 1.0 : -expm1(-t) / t; }
static double orc_dphi(double t) {
    if (fabs(t) < 0.02) {   /* series of d/dt (1-e^-t)/t */
        double s = 0.0, term = 1.0;   /* sum_{n>=1} n (-1)^n t^(n-1) / (n+1)! */
        double fact = 1.0;            /* (n+1)! built incrementally */
        for (int n = 1; n <= 12; n++) {
            fact *= (double)(n + 1);
            s += (n % 2 ? -1.0 : 1.0) * (double)n * term / fact;
            term *= t;
        }
        return s;
    }
    return (exp(-t) - orc_phi(t)) / t;
}

void orc_los_layers_jac(const double* tau, const double* emi, const double* tau_g,
                        const double* emi_g, const double* dfrac, int n_par, const int* n_steps,
                        int n_los, int n_steps_max, long n_pts, const double* i0,
                        int solo_absorption, double* rad, double* jac) {
    if (!tau_g) { tau_g = tau; emi_g = emi; }
    double* below = (double*)malloc(sizeof(double) * (n_steps_max + 1));
    for (int l = 0; l < n_los; l++)
        for (long i = 0; i < n_pts; i++) {
            const int ns = n_steps[l];
            const size_t base = (size_t)l * n_steps_max * n_pts + i;
            const double I0 = i0 ? i0[(size_t)l * n_pts + i] : 0.0;
            /* suffix optical depths: od[k] = sum_{j>k} tau_j */
            double total = 0.0;
            for (int k = 0; k < ns; k++) total += tau[base + (size_t)k * n_pts];
            /* below[k] = I_0 T(1..N) + sum_{j<k} J_j phi_j T(j+1..N) */
            double acc = I0 * exp(-total), run = 0.0;
            for (int k = 0; k < ns; k++) {
                const double t = tau[base + (size_t)k * n_pts];
                below[k] = acc;
                run += t;
                if (!solo_absorption)
                    acc += emi[base + (size_t)k * n_pts] * orc_phi(t) * exp(-(total - run));
            }
            rad[(size_t)l * n_pts + i] = acc;
            for (int p = 0; p < n_par; p++) {
                double d = 0.0;
                run = 0.0;
                for (int k = 0; k < ns; k++) {
                    const double t = tau[base + (size_t)k * n_pts];
                    const double tg = tau_g[base + (size_t)k * n_pts];
                    run += t;
                    const double f = dfrac[((size_t)l * n_steps_max + k) * n_par + p];
                    if (f == 0.0) continue;
                    double own = 0.0;
                    if (!solo_absorption)
                        own = (emi_g[base + (size_t)k * n_pts] * orc_phi(t) +
                               emi[base + (size_t)k * n_pts] * orc_dphi(t) * tg) *
                              exp(-(total - run));
                    d += f * (own - tg * below[k]);
                }
                jac[((size_t)l * n_par + p) * n_pts + i] = d;
            }
        }
    free(below);
}
