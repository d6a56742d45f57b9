// TEST INFRASTRUCTURE: the product's device functions (spectrobot_b200/csrc/sr_device.cuh - the
// arithmetic every kernel is built from) compiled for the host through shim/cuda_runtime.h and
// exported with a C ABI for tests/test_device_math.py.  Built into a temporary directory by the
// test; never shipped, never loaded by the product.
#include "cuda_runtime.h"
#include "sr_device.cuh"

extern "C" {

// humliv_bb, line inside the window (lineshape.f:443-562), assembled the way k_humliv_bb does it:
// closed-form abscissae, regions by the Fortran's index rules.  mode 0: the literal forms
// (humliv_core / humliv_reg2 / humliv_reg1), mode 1: the hot-loop forms of the tile kernel
// (humliv_core_fast / humliv_reg2_eval / humliv_reg1_fast).  reg[4] = il, ir, il2, ir2 (1-based).
int dm_humliv_inside(const double* x, int n, double x0, double lw, double dw, int mode, double* y,
                     int* reg) {
    using namespace srdev;
    const double ry = lw / dw, xstep = (x[1] - x[0]) / dw;
    const int i1 = 1, i2 = n;
    double rx = (x0 - x[i1 - 1]) / dw;
    int il = i1, ir = i2;
    if (rx + ry >= 15.0) il = (int)(f_nint((rx - ry - 15.0) / xstep) > 0 ? f_nint((rx - ry - 15.0) / xstep) : 0) + i1;
    rx = (x[i2 - 1] - x0) / dw;
    if (rx + ry >= 15.0) ir = i2 - (int)(f_nint((rx - ry - 15.0) / xstep) > 0 ? f_nint((rx - ry - 15.0) / xstep) : 0);
    int il2 = il, ir2 = ir;
    rx = (x0 - x[il - 1]) / dw;
    if (rx + ry >= 5.5) il2 = il + (int)(f_nint((rx - ry - 5.5) / xstep) > 0 ? f_nint((rx - ry - 5.5) / xstep) : 0);
    rx = (x[ir - 1] - x0) / dw;
    if (rx + ry >= 5.5) ir2 = ir - (int)(f_nint((rx - ry - 5.5) / xstep) > 0 ? f_nint((rx - ry - 5.5) / xstep) : 0);
    reg[0] = il; reg[1] = ir; reg[2] = il2; reg[3] = ir2;
    const bool r1 = il > i1 || ir < i2, r2 = il2 > il || ir2 < ir;
    const int c_lo = (il2 == il ? il - 1 : il2) + 1, c_hi = (ir2 == ir ? ir + 1 : ir2) - 1;
    const reg2_coef k2 = humliv_reg2_coefs(ry);
    const double c2 = 2.0 * ry * ry, b4 = 0.25 * 2.2567584 * ry;
    for (int j = 1; j <= n; j++) {
        const double xr = fabs(x[j - 1] - x0) / dw;
        const bool in1 = r1 && ((il > i1 && j <= il) || (ir < i2 && j >= ir));
        const bool in2 = r2 && ((il < il2 && j >= il && j <= il2) || (ir2 < ir && j >= ir2 && j <= ir));
        const bool core = j >= c_lo && j <= c_hi;
        double v;
        if (core) v = mode ? humliv_core_fast(xr, ry) : humliv_core(xr, ry);
        else if (in2) v = mode ? humliv_reg2_eval(k2, xr * xr) : humliv_reg2(xr * xr, ry);
        else if (in1) v = mode ? b4 * humliv_reg1_fast(xr * xr + ry * ry - 0.5, c2) : humliv_reg1(xr * xr, ry);
        else return 1;
        y[j - 1] = v;
    }
    return 0;
}

void dm_reg1_forms(const double* u, int n, double c2, double* fast, double* uw, double* uu) {
    for (int i = 0; i < n; i++) {
        fast[i] = srdev::humliv_reg1_fast(u[i], c2);
        uw[i] = srdev::humliv_reg1_uw(u[i], u[i] + 1.0, c2);
        uu[i] = srdev::humliv_reg1_u(u[i], c2);
    }
}

double dm_curgod(int k, const double* nd, const double* vmr, const double* f, const double* x, int n_p) {
    double res = 0.0;
    for (int i = 0; i + 1 < n_p; i++) {
        const double dx = x[i + 1] - x[i];
        if (k == 1) res += srdev::curgod_seg1(nd[i], nd[i + 1], dx);
        else if (k == 2) res += srdev::curgod_seg2(nd[i], nd[i + 1], vmr[i], vmr[i + 1], dx);
        else if (k == 3) res += srdev::curgod_seg3(nd[i], nd[i + 1], vmr[i], vmr[i + 1], f[i], f[i + 1], dx);
        else res += srdev::curgod_seg4(nd[i], nd[i + 1], vmr[i], vmr[i + 1], f[i], f[i + 1], dx);
    }
    return res;
}

void dm_exp_pair(const double* x, int n, double* ex, double* em) {
    for (int i = 0; i < n; i++) srdev::exp_pair(x[i], ex[i], em[i]);
}
void dm_exp_phi(const double* t, int n, double* ex, double* phi) {
    for (int i = 0; i < n; i++) srdev::exp_phi(t[i], ex[i], phi[i]);
}
// one layer of the LOS recursion, I <- I e^-t + J (1 - e^-t)/t  (DESIGN.md 6.4), every short form
void dm_layer_update(const double* I, const double* t, const double* J, int n, int form, int solo,
                     double* out) {
    for (int i = 0; i < n; i++) {
        if (form == 0) out[i] = srdev::layer_update_j(I[i], t[i], J[i], solo != 0);
        else if (form == 1) out[i] = srdev::layer_update_j_small(I[i], t[i], J[i], solo != 0);
        else if (form == 2) out[i] = srdev::layer_update_j_medium(I[i], t[i], J[i], solo != 0);
        else out[i] = srdev::layer_update_j_f32in(I[i], t[i], J[i], solo != 0);
    }
}

}  // extern "C"
