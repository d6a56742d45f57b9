// sr_tier1.cu -- Tier-1 drop-ins for the reference's f2py modules (host pointers, synchronous):
//   lineshape.humliv_bb      lineshape.f:226-569
//   lineshape.sum_all_lines  lineshape.f:2-25
//   fparts_mod.bd_tips_2003  fparts_mod.f:33-295   (+ CalcPartitionSum, spect_classes.py:1692)
//   curgods.curgod_fort_1..4 curgods.f:2-98
// These exist so that spect_classes.py-style code that still calls the f2py API one line at a
// time keeps working; the fused Tier-2 path (sr_voigt.cu, sr_los.cu) is the product.
#include <cmath>
#include <vector>
#include "sr_common.h"
#include "sr_device.cuh"

#include "tips2003_tables.inc"

namespace {

// ---------------------------------------------------------------------------------------------
// humliv_bb: thread 0 walks the Fortran's control flow and emits up to 5 segments; the block
// then fills the segments in the Fortran's write order (later segments overwrite earlier ones,
// exactly like lineshape.f:455-562).  Inside a segment the running abscissa xrun = xrun +- xstep
// (lineshape.f:467,476,510,520,338,354) is evaluated in closed form xrun0 + m*step.
// ---------------------------------------------------------------------------------------------
enum { SEG_REG1 = 1, SEG_REG2 = 2, SEG_CORE_RUN = 3, SEG_CORE_X = 4 };
struct Seg { int type, kb, ke; double x0run, step; };

__device__ void plan_humliv(const double* __restrict__ xv, int n, int i1, int i2, double x0,
                            double ry, double dw, double xstep, Seg* seg, int* n_seg) {
    auto X = [&](long k) { k = k < 1 ? 1 : (k > n ? n : k); return xv[k - 1]; };  // 1-based
    int ns = 0;
    auto push = [&](int type, long kb, long ke, double x0run, double step) {
        if (kb < i1) { x0run += step * (double)(i1 - kb); kb = i1; }
        if (ke > i2) ke = i2;
        if (kb > ke) return;
        seg[ns].type = type; seg[ns].kb = (int)kb; seg[ns].ke = (int)ke;
        seg[ns].x0run = x0run; seg[ns].step = step;
        ns++;
    };
    using srdev::f_nint;
    if (x0 <= X(i1)) {                                  // forward, lineshape.f:272-357
        long j = i1;
        double rx0 = (X(j) - x0) / dw, rx = rx0;
        while ((rx + ry < 5.5) && (j <= i2)) { j++; rx += xstep; }
        push(SEG_CORE_RUN, i1, j - 1, rx0, xstep);
        if (j <= i2) {
            long l = max(f_nint((15.0 - ry - rx) / xstep), 0LL) + j;
            l = min(l, (long)i2);
            if (l > j) { push(SEG_REG2, j, l, (X(j) - x0) / dw, xstep); l = l + 1; }
            if (l < j) l = j;
            if (l < i2) push(SEG_REG1, l, i2, (X(l) - x0) / dw, xstep);
        }
    } else if (x0 >= X(i2)) {                           // backward, lineshape.f:358-442
        long j = i2;
        double rx0 = (x0 - X(j)) / dw, rx = rx0;
        while ((rx + ry < 5.5) && (j >= i1)) { j--; rx += xstep; }
        // points j+1..i2 hold rx0 + (i2-k)*xstep
        push(SEG_CORE_RUN, j + 1, i2, rx0 + (double)(i2 - (j + 1)) * xstep, -xstep);
        if (j >= i1) {
            long l = j - max(f_nint((15.0 - ry - rx) / dw / xstep), 0LL);   // sic, :404
            l = max(l, (long)i1);
            if (l == i2) l = i2 + 1;
            if (l < j) push(SEG_REG2, l, j, (x0 - X(l)) / dw, -xstep);
            if (l >= i1) push(SEG_REG1, i1, l - 1, (x0 - X(i1)) / dw, -xstep);
        }
    } else {                                            // inside, lineshape.f:443-562
        double rx = (x0 - X(i1)) / dw;
        long il = i1, ir = i2;
        if (rx + ry >= 15.0) il = max(f_nint((rx - ry - 15.0) / xstep), 0LL) + i1;
        rx = (X(i2) - x0) / dw;
        if (rx + ry >= 15.0) ir = i2 - max(f_nint((rx - ry - 15.0) / xstep), 0LL);
        if (il > i1) push(SEG_REG1, i1, il, (x0 - X(i1)) / dw, -xstep);
        if (ir < i2) push(SEG_REG1, ir, i2, (X(ir) - x0) / dw, xstep);
        rx = (x0 - X(il)) / dw;
        long il2 = il, ir2 = ir;
        if (rx + ry >= 5.5) il2 = il + max(f_nint((rx - ry - 5.5) / xstep), 0LL);
        rx = (X(ir) - x0) / dw;
        if (rx + ry >= 5.5) ir2 = ir - max(f_nint((rx - ry - 5.5) / xstep), 0LL);
        if (il < il2) push(SEG_REG2, il, il2, (x0 - X(il)) / dw, -xstep);
        if (ir2 < ir) push(SEG_REG2, ir2, ir, (X(ir2) - x0) / dw, xstep);
        if (il2 == il) il2 = il - 1;
        if (ir2 == ir) ir2 = ir + 1;
        push(SEG_CORE_X, il2 + 1, ir2 - 1, 0.0, 0.0);
    }
    *n_seg = ns;
}

__global__ void k_humliv_bb(const double* __restrict__ x, int n, int i1, int i2, double x0,
                            double lw, double dw, double* __restrict__ y) {
    __shared__ Seg seg[6];
    __shared__ int n_seg;
    const double ry = lw / dw;                               // :261
    const double xstep = (x[i1] - x[i1 - 1]) / dw;           // :265 (x(i1+1)-x(i1))
    if (threadIdx.x == 0) plan_humliv(x, n, i1, i2, x0, ry, dw, xstep, seg, &n_seg);
    __syncthreads();
    for (int s = 0; s < n_seg; s++) {
        const Seg sg = seg[s];
        for (int k = sg.kb + threadIdx.x; k <= sg.ke; k += blockDim.x) {
            double v;
            if (sg.type == SEG_CORE_X) {
                v = srdev::humliv_core(fabs(x[k - 1] - x0) / dw, ry);
            } else {
                const double xr = fma((double)(k - sg.kb), sg.step, sg.x0run);
                if (sg.type == SEG_CORE_RUN) v = srdev::humliv_core(xr, ry);
                else if (sg.type == SEG_REG2) v = srdev::humliv_reg2(xr * xr, ry);
                else v = srdev::humliv_reg1(xr * xr, ry);
            }
            y[k - 1] = v;
        }
        __syncthreads();
    }
}

// sum_all_lines: threads run over lines (coalesced reads of the column-major matrix), blocks in
// y over window columns; FP64 atomics into the spectrum (order of addition is not the Fortran's
// line order -> differences at the 1e-16 level).
__global__ void k_sum_all_lines(const double* __restrict__ matrix, const int* __restrict__ init,
                                const int* __restrict__ fin, int n_lines, int ld, int n_win,
                                int n_spe, double* __restrict__ spe, int* __restrict__ err) {
    const int ilin = blockIdx.x * blockDim.x + threadIdx.x;
    if (ilin >= n_lines) return;
    const int i0 = blockIdx.y * 64;
    const int first = init[ilin], last = fin[ilin];
    for (int i = i0; i < min(i0 + 64, n_win); i++) {
        const int j = first + i;  // 1-based spectrum index
        if (j > last) break;
        if (j < 1 || j > n_spe) { atomicOr(err, 1); break; }
        const double v = matrix[(size_t)i * ld + ilin];
        if (v != 0.0) atomicAdd(spe + (j - 1), v);
    }
}

__global__ void k_curgod(int kind, const double* __restrict__ nd, const double* __restrict__ vmr,
                         const double* __restrict__ f, const double* __restrict__ x, int n_p,
                         int n_batch, double* __restrict__ res_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batch) return;
    const size_t o = (size_t)b * n_p;
    double res = 0.0;
    for (int i = 0; i < n_p - 1; i++) {
        const double dx = x[o + i + 1] - x[o + i];
        if (kind == 1)
            res = res + srdev::curgod_seg1(nd[o + i], nd[o + i + 1], dx);
        else if (kind == 2)
            res = res + srdev::curgod_seg2(nd[o + i], nd[o + i + 1], vmr[o + i], vmr[o + i + 1], dx);
        else if (kind == 3)
            res = res + srdev::curgod_seg3(nd[o + i], nd[o + i + 1], vmr[o + i], vmr[o + i + 1],
                                           f[o + i], f[o + i + 1], dx);
        else
            res = res + srdev::curgod_seg4(nd[o + i], nd[o + i + 1], vmr[o + i], vmr[o + i + 1],
                                           f[o + i], f[o + i + 1], dx);
    }
    res_out[b] = res;
}

}  // namespace

extern "C" {

int sr_humliv_bb(const double* x, int n, int i1, int i2, double x0, double lw, double dw,
                 double* y) {
    if (!x || !y || n < 2) return sr::fail(SR_ERR_ARG, "sr_humliv_bb: bad argument");
    if (i1 > i2) return sr::fail(SR_ERR_ARG, "Error in humliv: called with i1 > i2");   // :255
    if (!(dw > 0.0)) return sr::fail(SR_ERR_DW, "Error in humliv: called with dw <=0"); // :263
    if (i1 < 1 || i2 > n || i1 + 1 > n)
        return sr::fail(SR_ERR_ARG, "sr_humliv_bb: i1=%d i2=%d outside 1..%d", i1, i2, n);
    sr::DevBuf<double> dx, dy;
    SR_CUDA(dx.upload(x, n));
    SR_CUDA(dy.upload(y, n));   // entries outside the written ranges keep the caller's values
    SR_LAUNCH(k_humliv_bb, 1, 256, 0, 0, dx.p, n, i1, i2, x0, lw, dw, dy.p);
    SR_CUDA(cudaMemcpy(y, dy.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return SR_OK;
}

int sr_sum_all_lines(const double* spe_ini, const double* matrix_colmajor, const int* init,
                     const int* fin, int n_lines, int ld_lines, int n_win, int n_spe,
                     double* spe_fin) {
    if (!spe_ini || !matrix_colmajor || !init || !fin || !spe_fin || n_lines < 0 ||
        ld_lines < n_lines || n_win < 1 || n_spe < 1)
        return sr::fail(SR_ERR_ARG, "sr_sum_all_lines: bad argument");
    sr::DevBuf<double> dspe, dmat;
    sr::DevBuf<int> dinit, dfin, derr;
    SR_CUDA(dspe.upload(spe_ini, n_spe));
    if (n_lines > 0) {
        // only the first n_lines rows of every column are touched (lineshape.f:17)
        SR_CUDA(dmat.alloc((size_t)n_lines * n_win));
        SR_CUDA(cudaMemcpy2D(dmat.p, sizeof(double) * n_lines, matrix_colmajor,
                             sizeof(double) * ld_lines, sizeof(double) * n_lines, n_win,
                             cudaMemcpyHostToDevice));
        SR_CUDA(dinit.upload(init, n_lines));
        SR_CUDA(dfin.upload(fin, n_lines));
        SR_CUDA(derr.alloc(1));
        SR_CUDA(cudaMemset(derr.p, 0, sizeof(int)));
        dim3 grid((n_lines + 127) / 128, (n_win + 63) / 64);
        SR_LAUNCH(k_sum_all_lines, grid, 128, 0, 0, dmat.p, dinit.p, dfin.p, n_lines, n_lines,
                  n_win, n_spe, dspe.p, derr.p);
        int err = 0;
        SR_CUDA(cudaMemcpy(&err, derr.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (err) return sr::fail(SR_ERR_ARG, "sum_all_lines: init/fin outside 1..n_spe");
    }
    SR_CUDA(cudaMemcpy(spe_fin, dspe.p, sizeof(double) * n_spe, cudaMemcpyDeviceToHost));
    return SR_OK;
}

int sr_bd_tips_2003(int mol, int iso, double* gi, double* t119, double* q119) {
    if (!gi || !t119 || !q119) return sr::fail(SR_ERR_ARG, "sr_bd_tips_2003: bad argument");
    for (int i = 0; i < SR_TIPS_NT; i++) t119[i] = 60.0 + 25.0 * i;   // fparts_mod.f:58-76
    for (int m = 0; m < SR_TIPS_NMOL; m++) {
        if (sr_tips_index[m][0] != mol) continue;
        if (iso < 1 || iso > sr_tips_index[m][2]) break;
        const int row = sr_tips_index[m][1] + iso - 1;
        *gi = sr_tips_gj[row];
        for (int i = 0; i < SR_TIPS_NT; i++) {
            float fv;
            memcpy(&fv, &sr_tips_qbits[row][i], sizeof(float));
            q119[i] = (double)fv;   // real*4 literal stored in a DOUBLE PRECISION array (F3)
        }
        return SR_OK;
    }
    return sr::fail(SR_ERR_TABLE, "bd_tips_2003: no partition-sum table for mol %d iso %d", mol,
                    iso);
}

int sr_partition_sum(int mol, int iso, double temp, double* q) {
    if (!q) return sr::fail(SR_ERR_ARG, "sr_partition_sum: bad argument");
    double gi, t[SR_TIPS_NT], qt[SR_TIPS_NT];
    int rc = sr_bd_tips_2003(mol, iso, &gi, t, qt);
    if (rc) return rc;
    // nodes: the last two with T_grid <= temp and the first two with T_grid > temp
    // (spect_classes.py:1698-1704); scipy.interpolate.lagrange through them, evaluated at temp.
    int nle = 0;
    while (nle < SR_TIPS_NT && t[nle] <= temp) nle++;
    const int lo = nle - 2 < 0 ? 0 : nle - 2;
    const int hi = nle + 2 > SR_TIPS_NT ? SR_TIPS_NT : nle + 2;
    double acc = 0.0;
    for (int a = lo; a < hi; a++) {
        double w = qt[a];
        for (int b = lo; b < hi; b++)
            if (b != a) w *= (temp - t[b]) / (t[a] - t[b]);
        acc += w;
    }
    *q = acc;
    return SR_OK;
}

int sr_curgod(int k, const double* nd, const double* vmr, const double* f, const double* x,
              int n_p, int n_batch, double* res) {
    if (k < 1 || k > 4 || !nd || !x || !res || n_p < 1 || n_batch < 1 || (k >= 2 && !vmr) ||
        (k >= 3 && !f))
        return sr::fail(SR_ERR_ARG, "sr_curgod: bad argument");
    const size_t n = (size_t)n_p * n_batch;
    sr::DevBuf<double> dnd, dvmr, df, dx, dres;
    SR_CUDA(dnd.upload(nd, n));
    SR_CUDA(dx.upload(x, n));
    if (k >= 2) SR_CUDA(dvmr.upload(vmr, n));
    if (k >= 3) SR_CUDA(df.upload(f, n));
    SR_CUDA(dres.alloc(n_batch));
    SR_LAUNCH(k_curgod, (n_batch + 127) / 128, 128, 0, 0, k, dnd.p, dvmr.p, df.p, dx.p, n_p,
              n_batch, dres.p);
    SR_CUDA(cudaMemcpy(res, dres.p, sizeof(double) * n_batch, cudaMemcpyDeviceToHost));
    return SR_OK;
}

// The four f2py entry points under their own names (curgods.f:2, 24, 48, 76): one integral each.
int sr_curgod_1(const double* nd, const double* x, int n_p, double* res) {
    return sr_curgod(1, nd, nullptr, nullptr, x, n_p, 1, res);
}
int sr_curgod_2(const double* nd, const double* vmr, const double* x, int n_p, double* res) {
    return sr_curgod(2, nd, vmr, nullptr, x, n_p, 1, res);
}
int sr_curgod_3(const double* nd, const double* vmr, const double* f, const double* x, int n_p,
                double* res) {
    return sr_curgod(3, nd, vmr, f, x, n_p, 1, res);
}
int sr_curgod_4(const double* nd, const double* vmr, const double* f, const double* x, int n_p,
                double* res) {
    return sr_curgod(4, nd, vmr, f, x, n_p, 1, res);
}

}  // extern "C"
