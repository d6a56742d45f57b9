// sr_voigt.cu -- K1/K2: fused Voigt + G-coefficient cross-sections for whole (P,T) cells.
//
// Replaces, for one isotopologue and a batch of LUT cells, the reference chain
//   calc_shapes_lines / PrepareCalcShapes   spect_classes.py:1378-1462
//   MakeShapeLine -> MakeShape -> humliv_bb  spect_classes.py:174-206, 1990-2008; lineshape.f:226-569
//   Calc_Gcoeffs                             spect_classes.py:312-343, 1806-1853
//   BuildCoeff -> add_lines_to_spectrum      spect_classes.py:1277-1337, 1016-1147
//   sum_all_lines                            lineshape.f:2-25
//   LookUpTable.make / LutSet.add_PT loops   spect_main_module.py:753-774, 1122-1168
//
// Data layout in HBM
//   line table      SoA doubles/ints, sorted by (group=(upper set, lower set), centre index)
//   LineCell rec    [cell][line] 128-byte records: per-(line,cell) widths, region boundaries and
//                   G coefficients, written by k_line_cell_params, read by the tile kernel
//   out             [cell][set][ctype][n_grid] doubles, each element written exactly once
//
// Kernel k_voigt_tile: one CTA owns TP = 256*PPT consecutive grid points of one cell and keeps
// the n_sets*3 output rows of that tile in shared memory.  It walks the lines whose 13010-point
// window touches the tile, group by group; every thread accumulates the three G-weighted sums of
// its PPT points in FP64 registers and flushes them to the shared tile when the group changes.
// ~98% of the (line, point) pairs are far-wing (region 1) and take the branch-free 10-FP64-op
// path; pairs near the line centre take the general path (regions 2/3/4).
#include <algorithm>
#include <numeric>
#include <vector>
#include "sr_common.h"
#include "sr_device.cuh"

namespace {

constexpr int N_WIN = SR_IMXSIG;       // 13010
constexpr int HALF = SR_IMXSIG / 2;    // 6505: window index of the centre point (0-based)
constexpr int THREADS = 256;
constexpr int CHUNK = 128;             // lines staged per step
constexpr int MAX_GROUPS = 1024;
constexpr double HPA_TO_ATM = 0.00098692326671601;  // spect_classes.py:40
constexpr double T_REF = 296.0;                      // spect_classes.py:39

enum : int { FLAG_OUTSIDE = 1, FLAG_GEOMETRY = 2, FLAG_NONFINITE = 4 };

struct __align__(16) LineCell {
    double xs;      // xstep = (x(2)-x(1))/dw'                      lineshape.f:265
    double c1;      // ry^2 - 0.5   (region-1 rewrite, sr_device.cuh)
    double c2;      // 2 ry^2
    double b4;      // b/4 = 0.5641896 ry
    double baseL1;  // region-1 left : x(j0) = baseL1 + j0*xs   (j0 = 0-based window index)
    double baseR1;  // region-1 right
    double baseL2;  // region-2 left
    double baseR2;  // region-2 right
    double ry;      // lw/dw'
    double dwp;     // dw' = dw/sqrt(ln2)
    double gs[3];   // G_ctype / fac,  fac = dw*sqrt(pi/ln2)       spect_classes.py:1997,2003
    int il, ir;     // 1-based, as left by lineshape.f:446-454
    int il2, ir2;   // 1-based, as left by lineshape.f:482-490 (before :524-525)
    int flags, pad;
};
static_assert(sizeof(LineCell) == 128, "LineCell must be 128 bytes");

struct __align__(16) Staged {
    double A, B, C, c2;  // u(p) = A + B p + C p^2 for the fast region-1 path
    double g1[3];        // gs * b4
    int type;            // 0 skip, 1 fast region 1, 2 general
    int grp;
    int line;            // sorted line index
    int j0_first;        // 0-based window index of tile point 0
    int pad[2];
};
static_assert(sizeof(Staged) == 80, "Staged must be 80 bytes");

struct LineArrays {
    const double *freq, *a_coeff, *air, *tdep, *e_lower, *g_up, *g_lo, *evu, *evl;
    const double* gc;   // grid[ind]
    const int* ind;     // closest grid index
};

// ---------------------------------------------------------------------------------------------
// closest_grid (spect_classes.py:1937-1943): argmin |grid - nu0|, ties -> lowest index
// ---------------------------------------------------------------------------------------------
__global__ void k_closest_grid(const double* __restrict__ grid, long n_grid,
                               const double* __restrict__ freq, int n_lines,
                               int* __restrict__ ind, double* __restrict__ gc) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    double nu0 = freq[l];
    long lo = 0, hi = n_grid - 1;
    while (hi - lo > 1) {
        long mid = (lo + hi) >> 1;
        if (grid[mid] <= nu0) lo = mid; else hi = mid;
    }
    long best = lo;
    double bd = fabs(grid[lo] - nu0);
    for (long c = lo - 1; c <= hi + 1; c++) {
        if (c < 0 || c >= n_grid) continue;
        double d = fabs(grid[c] - nu0);
        if (d < bd || (d == bd && c < best)) { bd = d; best = c; }
    }
    ind[l] = (int)best;
    gc[l] = grid[best];
}

// ---------------------------------------------------------------------------------------------
// per-(cell,line) prologue: widths, region boundaries, G coefficients
// ---------------------------------------------------------------------------------------------
struct ParamsArgs {
    LineArrays L;
    const double* lin;   // window offsets [N_WIN]
    const double* pt;    // [n_cells][2]
    LineCell* rec;       // [n_cells][n_lines]
    int* flags;          // [1] OR of all record flags
    int n_lines, n_cells;
    double mm;
    sr_consts c;
};

__global__ void k_line_cell_params(ParamsArgs a) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    int cell = blockIdx.y;
    if (l >= a.n_lines) return;
    const double pres = a.pt[2 * cell], temp = a.pt[2 * cell + 1];
    const double nu0 = a.L.freq[l], gc = a.L.gc[l];
    const sr_consts& k = a.c;
    LineCell r;
    int flags = 0;

    // --- widths: convert_to_atm :2034, Lorenz_width :1972 (air only), Doppler :1984
    double pres_atm = pres * HPA_TO_ATM;
    double lw = pow(T_REF / temp, a.L.tdep[l]) * (a.L.air[l] * (pres_atm - 0.0) + 0.0 * 0.0);
    double dw = nu0 / k.c_cgs * sqrt(2 * k.avogadro * k.k_cgs * temp * k.ln2 / a.mm);
    double fac = dw * k.sqrt_pi_ln2;   // MakeShape :1997
    double dwp = dw / k.sqrt_ln2;      // MakeShape :1999
    if (!(dwp > 0.0)) flags |= FLAG_NONFINITE;

    // --- G coefficients: Calc_Gcoeffs :312-343, Einstein_A_to_Gcoeff_* :1806-1853
    double g[3] = {0.0, 0.0, 0.0};
    {
        double A = a.L.a_coeff[l], gu = a.L.g_up[l], gl = a.L.g_lo[l], el = a.L.e_lower[l];
        if (A != 0.0 && gl != 0.0 && gu != 0.0) {
            double c2k = k.h_cgs * k.c_cgs / k.k_cgs;                         // :47
            double fact_2 = 2 * k.h_cgs * (k.c_cgs * k.c_cgs) * (nu0 * nu0 * nu0);   // :1743
            double b21 = A / fact_2;                                          // :1750
            double rot_up = gu * exp(-c2k * (el + nu0 - a.L.evu[l]) / temp);  // :1850
            g[0] = k.h_cgs * k.c_cgs * nu0 * rot_up * A / (4 * M_PI);         // :1851
            g[1] = k.h_cgs * k.c_cgs * nu0 * rot_up * b21 / (4 * M_PI);       // :1840
            double b12 = b21 * gu / gl;                                       // :1783
            double rot_lo = gl * exp(-c2k * (el - a.L.evl[l]) / temp);        // :1815
            g[2] = k.h_cgs * k.c_cgs * nu0 * rot_lo * b12 / (4 * M_PI);       // :1817
        }
    }
    r.gs[0] = g[0] / fac;
    r.gs[1] = g[1] / fac;
    r.gs[2] = g[2] / fac;

    // --- humliv_bb set-up, branch x(i1) < x0 < x(i2), lineshape.f:260-267, 443-490
    const int i1 = 1, i2 = N_WIN;
    double ry = lw / dwp;
    double x1 = a.lin[0] + gc, x2 = a.lin[1] + gc, xN = a.lin[N_WIN - 1] + gc;  // spcl:1455
    double xs = (x2 - x1) / dwp;
    if (!(nu0 > x1 && nu0 < xN)) flags |= FLAG_OUTSIDE;
    double rx = (nu0 - x1) / dwp;
    int il = i1;
    if (rx + ry >= 15.0) il = (int)max(srdev::f_nint((rx - ry - 15.0) / xs), 0LL) + i1;
    double rxL = rx;
    rx = (xN - nu0) / dwp;
    int ir = i2;
    if (rx + ry >= 15.0) ir = i2 - (int)max(srdev::f_nint((rx - ry - 15.0) / xs), 0LL);
    il = min(max(il, 1), N_WIN);  // keep the table reads below in range; geometry is re-checked
    ir = min(max(ir, 1), N_WIN);
    double x_il = a.lin[il - 1] + gc, x_ir = a.lin[ir - 1] + gc;
    double dL = (nu0 - x_il) / dwp;
    int il2 = il;
    if (dL + ry >= 5.5) il2 = il + (int)max(srdev::f_nint((dL - ry - 5.5) / xs), 0LL);
    double dR = (x_ir - nu0) / dwp;
    int ir2 = ir;
    if (dR + ry >= 5.5) ir2 = ir - (int)max(srdev::f_nint((dR - ry - 5.5) / xs), 0LL);
    // geometry the tile kernel relies on (always true for a window centred on the line)
    if (!(il2 <= ir && ir2 >= il && il2 >= il && ir2 <= ir && il2 < N_WIN && ir2 > 1))
        flags |= FLAG_GEOMETRY;
    il2 = min(max(il2, 1), N_WIN);
    ir2 = min(max(ir2, 1), N_WIN);
    double x_ir2 = a.lin[ir2 - 1] + gc;

    r.xs = xs;
    r.ry = ry;
    r.dwp = dwp;
    r.c1 = ry * ry - 0.5;
    r.c2 = 2.0 * ry * ry;
    r.b4 = 0.5641896 * ry;                       // (2.2567584 ry)/4, lineshape.f:457
    r.baseL1 = -rxL;                             // :462-467  xrun = rxL - j0*xs  (sign dropped)
    r.baseR1 = dR - (double)(ir - 1) * xs;       // :471-476
    r.baseL2 = (double)(1 - il) * xs - dL;       // :504-510
    r.baseR2 = (x_ir2 - nu0) / dwp - (double)(ir2 - 1) * xs;   // :514-520
    r.il = il; r.ir = ir; r.il2 = il2; r.ir2 = ir2;
    if (!isfinite(xs) || !isfinite(ry) || !isfinite(r.gs[0]) || !isfinite(r.gs[1]) ||
        !isfinite(r.gs[2]))
        flags |= FLAG_NONFINITE;
    r.flags = flags;
    r.pad = 0;
    a.rec[(size_t)cell * a.n_lines + l] = r;
    if (flags) atomicOr(a.flags, flags);
}

// ---------------------------------------------------------------------------------------------
// general evaluation of one window point (any region); j1 = 1-based window index.
// Last-writer-wins order of lineshape.f:455-562: core > region-2 right > region-2 left >
// region-1 right > region-1 left.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ double eval_window_point(const LineCell* __restrict__ rc, int j1,
                                                 double nu0, double gc,
                                                 const double* __restrict__ lin) {
    const int il = rc->il, ir = rc->ir, il2 = rc->il2, ir2 = rc->ir2;
    const int core_lo = (il2 > il) ? il2 + 1 : il;      // :524, :526
    const int core_hi = (ir2 < ir) ? ir2 - 1 : ir;      // :525, :526
    const double j0 = (double)(j1 - 1);
    if (j1 >= core_lo && j1 <= core_hi) {
        double x = lin[j1 - 1] + gc;                    // spect_classes.py:1455
        double rx = fabs(x - nu0) / rc->dwp;            // lineshape.f:527
        return srdev::humliv_core(rx, rc->ry);
    }
    if (ir2 < ir && j1 >= ir2 && j1 <= ir) {
        double x = fma(j0, rc->xs, rc->baseR2);
        return srdev::humliv_reg2(x * x, rc->ry);
    }
    if (il2 > il && j1 >= il && j1 <= il2) {
        double x = fma(j0, rc->xs, rc->baseL2);
        return srdev::humliv_reg2(x * x, rc->ry);
    }
    if (ir < N_WIN && j1 >= ir) {
        double x = fma(j0, rc->xs, rc->baseR1);
        return rc->b4 * srdev::humliv_reg1_fast(fma(x, x, rc->c1), rc->c2);
    }
    if (il > 1 && j1 <= il) {
        double x = fma(j0, rc->xs, rc->baseL1);
        return rc->b4 * srdev::humliv_reg1_fast(fma(x, x, rc->c1), rc->c2);
    }
    return 0.0;  // never written by the Fortran (cannot happen for a centred window)
}

// ---------------------------------------------------------------------------------------------
// K1/K2 tile kernel
// ---------------------------------------------------------------------------------------------
struct TileArgs {
    const LineCell* rec;     // [n_cells][n_lines]
    const double* nu0;       // sorted line arrays
    const double* gc;
    const int* ind;
    const int* grp_begin;    // [n_groups+1] offsets into the sorted line arrays
    const int* grp_up;       // [n_groups]
    const int* grp_lo;
    const double* lin;       // [N_WIN]
    double* out;             // [n_cells][n_sets][3][n_grid]
    long n_grid;
    int n_lines, n_sets, n_groups;
};

template <int PPT>
__global__ void __launch_bounds__(THREADS, 1) k_voigt_tile(TileArgs a) {
    constexpr int TP = THREADS * PPT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* acc_s = reinterpret_cast<double*>(smem_raw);                      // [n_sets*3][TP]
    Staged* stage = reinterpret_cast<Staged*>(acc_s + (size_t)a.n_sets * 3 * TP);   // [2][CHUNK]
    int* g_first = reinterpret_cast<int*>(stage + 2 * CHUNK);                 // [n_groups]
    int* g_cum = g_first + a.n_groups;                                        // [n_groups+1]

    const int tid = threadIdx.x;
    const int cell = blockIdx.y;
    const long tile0 = (long)blockIdx.x * TP;
    const LineCell* __restrict__ rec = a.rec + (size_t)cell * a.n_lines;

    for (int i = tid; i < a.n_sets * 3 * TP; i += THREADS) acc_s[i] = 0.0;

    // lines of each group whose window [ind-HALF, ind+HALF-1] touches [tile0, tile0+TP-1]
    const long ind_lo = tile0 - (HALF - 1), ind_hi = tile0 + TP - 1 + HALF;
    for (int g = tid; g < a.n_groups; g += THREADS) {
        int b = a.grp_begin[g], e = a.grp_begin[g + 1];
        int lo = b, hi = e;
        while (lo < hi) { int m = (lo + hi) >> 1; if (a.ind[m] < ind_lo) lo = m + 1; else hi = m; }
        int first = lo;
        hi = e;
        while (lo < hi) { int m = (lo + hi) >> 1; if (a.ind[m] <= ind_hi) lo = m + 1; else hi = m; }
        g_first[g] = first;
        g_cum[g + 1] = lo - first;
    }
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        g_cum[0] = 0;
        for (int g = 0; g < a.n_groups; g++) { run += g_cum[g + 1]; g_cum[g + 1] = run; }
    }
    __syncthreads();
    const int n_tot = g_cum[a.n_groups];
    const int n_chunks = (n_tot + CHUNK - 1) / CHUNK;

    double pd[PPT], acc0[PPT], acc1[PPT], acc2[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        pd[k] = (double)(tid + k * THREADS);
        acc0[k] = acc1[k] = acc2[k] = 0.0;
    }
    int cur_grp = -1;

    auto flush = [&](int grp) {
        if (grp < 0) return;
        const int up = a.grp_up[grp], lo = a.grp_lo[grp];
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int p = tid + k * THREADS;
            acc_s[(up * 3 + 0) * TP + p] += acc0[k];
            acc_s[(up * 3 + 1) * TP + p] += acc1[k];
            acc_s[(lo * 3 + 2) * TP + p] += acc2[k];
            acc0[k] = acc1[k] = acc2[k] = 0.0;
        }
    };

    // staging: slot n of the tile's line list -> Staged record (threads 0..CHUNK-1)
    auto stage_slot = [&](int n, Staged& s) {
        s.type = 0;
        if (n >= n_tot) return;
        int lo = 0, hi = a.n_groups;  // find g with g_cum[g] <= n < g_cum[g+1]
        while (hi - lo > 1) { int m = (lo + hi) >> 1; if (g_cum[m] <= n) lo = m; else hi = m; }
        const int g = lo;
        const int line = g_first[g] + (n - g_cum[g]);
        const LineCell* rc = rec + line;
        const int j0_first = (int)(tile0 - ((long)a.ind[line] - HALF));
        s.grp = g;
        s.line = line;
        s.j0_first = j0_first;
        const double4 q0 = *reinterpret_cast<const double4*>(&rc->xs);       // xs c1 c2 b4
        const int2 ilr = *reinterpret_cast<const int2*>(&rc->il);
        const int j1_first = j0_first + 1, j1_last = j0_first + TP;
        double base;
        bool fast = false;
        if (ilr.x > 1 && j1_first >= 1 && j1_last <= ilr.x - 1) { base = rc->baseL1; fast = true; }
        else if (j1_first >= ilr.y + 1 && j1_last <= N_WIN) { base = rc->baseR1; fast = true; }
        if (fast) {
            const double xoff = fma((double)j0_first, q0.x, base);
            s.A = fma(xoff, xoff, q0.y);
            s.B = 2.0 * xoff * q0.x;
            s.C = q0.x * q0.x;
            s.c2 = q0.z;
            s.g1[0] = rc->gs[0] * q0.w;
            s.g1[1] = rc->gs[1] * q0.w;
            s.g1[2] = rc->gs[2] * q0.w;
            s.type = 1;
        } else {
            s.type = 2;
        }
    };

    Staged pre;
    if (tid < CHUNK) stage_slot(tid, pre);
    for (int c = 0; c < n_chunks; c++) {
        Staged* buf = stage + (c & 1) * CHUNK;
        if (tid < CHUNK) buf[tid] = pre;
        __syncthreads();
        if (tid < CHUNK && c + 1 < n_chunks) stage_slot((c + 1) * CHUNK + tid, pre);
        const int n_here = min(CHUNK, n_tot - c * CHUNK);
        for (int i = 0; i < n_here; i++) {
            const Staged& s = buf[i];
            const int grp = s.grp;
            if (grp != cur_grp) { flush(cur_grp); cur_grp = grp; }
            if (s.type == 1) {
                const double A = s.A, B = s.B, C = s.C, c2 = s.c2;
                const double g0 = s.g1[0], g1 = s.g1[1], g2 = s.g1[2];
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    const double u = fma(fma(C, pd[k], B), pd[k], A);
                    const double kp = srdev::humliv_reg1_fast(u, c2);
                    acc0[k] = fma(g0, kp, acc0[k]);
                    acc1[k] = fma(g1, kp, acc1[k]);
                    acc2[k] = fma(g2, kp, acc2[k]);
                }
            } else {
                const int line = s.line;
                const LineCell* rc = rec + line;
                const double nu0 = a.nu0[line], gc = a.gc[line];
                const double g0 = rc->gs[0], g1 = rc->gs[1], g2 = rc->gs[2];
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    const int j1 = s.j0_first + tid + k * THREADS + 1;
                    if (j1 >= 1 && j1 <= N_WIN) {
                        const double v = eval_window_point(rc, j1, nu0, gc, a.lin);
                        acc0[k] = fma(g0, v, acc0[k]);
                        acc1[k] = fma(g1, v, acc1[k]);
                        acc2[k] = fma(g2, v, acc2[k]);
                    }
                }
            }
        }
    }
    flush(cur_grp);
    __syncthreads();

    // write the tile: every output element exactly once, coalesced
    const int n_rows = a.n_sets * 3;
    double* __restrict__ out = a.out + (size_t)cell * n_rows * a.n_grid;
    for (int row = 0; row < n_rows; row++) {
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int p = tid + k * THREADS;
            const long s = tile0 + p;
            if (s < a.n_grid) __stcs(out + (size_t)row * a.n_grid + s, acc_s[row * TP + p]);
        }
    }
}

// per-line shapes (MakeShapeLine keep_memory): one CTA per line, general evaluator
__global__ void k_line_shapes(const LineCell* __restrict__ rec, const double* __restrict__ nu0,
                              const double* __restrict__ gc, const double* __restrict__ lin,
                              double* __restrict__ shapes, double* __restrict__ gout,
                              const double* __restrict__ facs) {
    const int line = blockIdx.x;
    const LineCell* rc = rec + line;
    // shape = K/fac (spect_classes.py:2003); gs = G/fac  =>  G = gs*fac
    const double fac = facs[line];
    for (int j1 = threadIdx.x + 1; j1 <= N_WIN; j1 += blockDim.x) {
        double v = eval_window_point(rc, j1, nu0[line], gc[line], lin);
        shapes[(size_t)line * N_WIN + (j1 - 1)] = v / fac;
    }
    if (threadIdx.x < 3) gout[line * 3 + threadIdx.x] = rc->gs[threadIdx.x] * fac;
}

__global__ void k_line_fac(const double* __restrict__ nu0, int n_lines, double temp, double mm,
                           sr_consts k, double* __restrict__ facs) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    double dw = nu0[l] / k.c_cgs * sqrt(2 * k.avogadro * k.k_cgs * temp * k.ln2 / mm);
    facs[l] = dw * k.sqrt_pi_ln2;
}

__global__ void k_f64_to_f32(const double* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (float)in[i];   // numpy astype(float32): round-to-nearest
}

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
struct sr_lineset {
    int device = 0;
    long n_grid = 0;
    int n_sets = 0, n_in = 0, n_act = 0, n_groups = 0;
    double mm = 0.0;
    sr_consts c{};
    sr::DevBuf<double> grid, lin, freq, a_coeff, air, tdep, e_lower, g_up, g_lo, evu, evl, gc;
    sr::DevBuf<int> ind, grp_begin, grp_up, grp_lo, flags;
    sr::DevBuf<LineCell> rec;
    sr::DevBuf<double> pt, facs;
    std::vector<int> order;    // sorted position -> input line
    std::vector<int> ind_in;   // input line -> centre index (-1 dropped)
    int max_cells_per_batch = 1;
};

namespace {

template <int PPT>
size_t tile_smem(int n_sets, int n_groups) {
    return (size_t)n_sets * 3 * THREADS * PPT * sizeof(double) + 2 * CHUNK * sizeof(Staged) +
           (size_t)(2 * n_groups + 1) * sizeof(int) + 16;
}

int pick_ppt(int n_sets, int n_groups, size_t smem_max) {
    if (tile_smem<4>(n_sets, n_groups) <= smem_max && n_sets <= 3) return 4;
    if (tile_smem<2>(n_sets, n_groups) <= smem_max) return 2;
    if (tile_smem<1>(n_sets, n_groups) <= smem_max) return 1;
    return 0;
}

template <int PPT>
int launch_tile(const TileArgs& ta, int n_cells, size_t smem, cudaStream_t st) {
    SR_CUDA(cudaFuncSetAttribute(k_voigt_tile<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    const int TP = THREADS * PPT;
    dim3 grid((unsigned)((ta.n_grid + TP - 1) / TP), (unsigned)n_cells);
    SR_LAUNCH(k_voigt_tile<PPT>, grid, THREADS, smem, st, ta);
    return SR_OK;
}

int flags_to_status(int f) {
    if (f & FLAG_NONFINITE)
        return sr::fail(SR_ERR_DW, "humliv_bb: dw <= 0 or non-finite line parameters "
                                   "(lineshape.f:260-264)");
    if (f & FLAG_OUTSIDE)
        return sr::fail(SR_ERR_GEOMETRY, "line centre outside its own Voigt window");
    if (f & FLAG_GEOMETRY)
        return sr::fail(SR_ERR_GEOMETRY, "Voigt window with overlapping region boundaries "
                                         "(grid step larger than ~30 Doppler widths?)");
    return SR_OK;
}

}  // namespace

extern "C" {

void sr_default_consts(sr_consts* c) {
    c->h_cgs = 6.62607015e-34 * 1.e7;
    c->c_cgs = 299792458.0 * 1.e2;
    c->k_cgs = 1.380649e-23 * 1.e7;
    c->avogadro = 6.02214076e23;
    c->ln2 = log(2.0);
    c->sqrt_ln2 = sqrt(log(2.0));
    c->sqrt_pi_ln2 = sqrt(M_PI / log(2.0));
}

int sr_lineset_create(const sr_lines* lines, const double* grid, long n_grid,
                      const double* lin_grid, int n_sets, double mm, const sr_consts* consts,
                      sr_lineset** out) {
    if (!lines || !grid || !lin_grid || !out || n_grid < 2 || n_sets < 1 || !(mm > 0.0))
        return sr::fail(SR_ERR_ARG, "sr_lineset_create: bad argument");
    if (n_grid > SR_IMXSIG_LONG)
        return sr::fail(SR_ERR_LIMIT, "grid longer than imxsig_long = %d (spect_classes.py:362)",
                        SR_IMXSIG_LONG);
    sr_lineset* ls = new sr_lineset();
    SR_CUDA(cudaGetDevice(&ls->device));
    ls->n_grid = n_grid;
    ls->n_sets = n_sets;
    ls->mm = mm;
    if (consts) ls->c = *consts; else sr_default_consts(&ls->c);
    const int n_in = lines->n_lines;
    ls->n_in = n_in;
    ls->ind_in.assign(n_in, -1);

    // level-link filter (spect_classes.py:1384-1388)
    std::vector<int> act;
    act.reserve(n_in);
    for (int i = 0; i < n_in; i++) {
        int u = lines->up_set[i], l = lines->lo_set[i];
        if (u < 0 || l < 0) continue;
        if (u >= n_sets || l >= n_sets) {
            delete ls;
            return sr::fail(SR_ERR_ARG, "line %d: set index out of range", i);
        }
        act.push_back(i);
    }
    const int n_act = (int)act.size();
    ls->n_act = n_act;
    cudaStream_t st = 0;
    int rc = SR_OK;
    auto guard = [&](int code) { if (code != SR_OK && rc == SR_OK) rc = code; return code; };
    auto body = [&]() -> int {
        SR_CUDA(ls->grid.upload(grid, (size_t)n_grid, st));
        SR_CUDA(ls->lin.upload(lin_grid, N_WIN, st));
        SR_CUDA(ls->flags.alloc(1));
        SR_CUDA(cudaMemsetAsync(ls->flags.p, 0, sizeof(int), st));
        if (n_act == 0) { ls->n_groups = 0; return SR_OK; }
        // closest_grid on the device for the active lines (input order)
        std::vector<double> tmp(n_act);
        for (int i = 0; i < n_act; i++) tmp[i] = lines->freq[act[i]];
        SR_CUDA(ls->freq.upload(tmp.data(), n_act, st));
        SR_CUDA(ls->ind.alloc(n_act));
        SR_CUDA(ls->gc.alloc(n_act));
        SR_LAUNCH(k_closest_grid, (n_act + 255) / 256, 256, 0, st, ls->grid.p, n_grid,
                  ls->freq.p, n_act, ls->ind.p, ls->gc.p);
        std::vector<int> ind(n_act);
        SR_CUDA(cudaMemcpyAsync(ind.data(), ls->ind.p, n_act * sizeof(int),
                                cudaMemcpyDeviceToHost, st));
        SR_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < n_act; i++) ls->ind_in[act[i]] = ind[i];
        // sort by (group, centre index, input position)
        std::vector<int> perm(n_act);
        std::iota(perm.begin(), perm.end(), 0);
        auto key = [&](int i) { return (long long)lines->up_set[act[i]] * n_sets + lines->lo_set[act[i]]; };
        std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) {
            long long kx = key(x), ky = key(y);
            if (kx != ky) return kx < ky;
            return ind[x] < ind[y];
        });
        ls->order.resize(n_act);
        std::vector<int> gb, gu, gl, sind(n_act);
        long long prev = -1;
        for (int s = 0; s < n_act; s++) {
            int i = perm[s];
            ls->order[s] = act[i];
            sind[s] = ind[i];
            long long k = key(i);
            if (k != prev) {
                gb.push_back(s);
                gu.push_back(lines->up_set[act[i]]);
                gl.push_back(lines->lo_set[act[i]]);
                prev = k;
            }
        }
        gb.push_back(n_act);
        ls->n_groups = (int)gu.size();
        if (ls->n_groups > MAX_GROUPS)
            return sr::fail(SR_ERR_LIMIT, "%d (upper,lower) level pairs > %d", ls->n_groups,
                            MAX_GROUPS);
        auto up_sorted = [&](const double* src, sr::DevBuf<double>& dst) -> int {
            for (int s = 0; s < n_act; s++) tmp[s] = src[ls->order[s]];
            SR_CUDA(dst.upload(tmp.data(), n_act, st));
            SR_CUDA(cudaStreamSynchronize(st));  // tmp is reused
            return SR_OK;
        };
        if (guard(up_sorted(lines->freq, ls->freq))) return rc;
        if (guard(up_sorted(lines->a_coeff, ls->a_coeff))) return rc;
        if (guard(up_sorted(lines->air_broad, ls->air))) return rc;
        if (guard(up_sorted(lines->t_dep, ls->tdep))) return rc;
        if (guard(up_sorted(lines->e_lower, ls->e_lower))) return rc;
        if (guard(up_sorted(lines->g_up, ls->g_up))) return rc;
        if (guard(up_sorted(lines->g_lo, ls->g_lo))) return rc;
        if (guard(up_sorted(lines->e_vib_up, ls->evu))) return rc;
        if (guard(up_sorted(lines->e_vib_lo, ls->evl))) return rc;
        SR_CUDA(ls->ind.upload(sind.data(), n_act, st));
        SR_CUDA(ls->grp_begin.upload(gb.data(), gb.size(), st));
        SR_CUDA(ls->grp_up.upload(gu.data(), gu.size(), st));
        SR_CUDA(ls->grp_lo.upload(gl.data(), gl.size(), st));
        // gc in sorted order
        SR_LAUNCH(k_closest_grid, (n_act + 255) / 256, 256, 0, st, ls->grid.p, n_grid,
                  ls->freq.p, n_act, ls->ind.p, ls->gc.p);
        SR_CUDA(cudaStreamSynchronize(st));
        return SR_OK;
    };
    int code = body();
    if (code != SR_OK) { delete ls; return code; }
    // cells per batch: keep the LineCell table under ~2 GiB
    size_t per_cell = (size_t)std::max(n_act, 1) * sizeof(LineCell);
    ls->max_cells_per_batch = (int)std::max<size_t>(1, std::min<size_t>(4096, ((size_t)2 << 30) / per_cell));
    *out = ls;
    return SR_OK;
}

int sr_lineset_destroy(sr_lineset* ls) {
    delete ls;
    return SR_OK;
}

long sr_lineset_n_active(const sr_lineset* ls) { return ls ? ls->n_act : 0; }

int sr_lineset_centres(const sr_lineset* ls, int* ind_host) {
    if (!ls || !ind_host) return sr::fail(SR_ERR_ARG, "sr_lineset_centres: bad argument");
    std::copy(ls->ind_in.begin(), ls->ind_in.end(), ind_host);
    return SR_OK;
}

int sr_lineset_order(const sr_lineset* ls, int* order_host) {
    if (!ls || !order_host) return sr::fail(SR_ERR_ARG, "sr_lineset_order: bad argument");
    std::copy(ls->order.begin(), ls->order.end(), order_host);
    return SR_OK;
}

static int run_params(sr_lineset* ls, const double* pt_host, int n_cells, cudaStream_t st) {
    SR_CUDA(ls->pt.ensure((size_t)2 * n_cells));
    SR_CUDA(cudaMemcpyAsync(ls->pt.p, pt_host, sizeof(double) * 2 * n_cells,
                            cudaMemcpyHostToDevice, st));
    SR_CUDA(ls->rec.ensure((size_t)n_cells * ls->n_act));
    ParamsArgs pa;
    pa.L = {ls->freq.p, ls->a_coeff.p, ls->air.p, ls->tdep.p, ls->e_lower.p, ls->g_up.p,
            ls->g_lo.p, ls->evu.p, ls->evl.p, ls->gc.p, ls->ind.p};
    pa.lin = ls->lin.p;
    pa.pt = ls->pt.p;
    pa.rec = ls->rec.p;
    pa.flags = ls->flags.p;
    pa.n_lines = ls->n_act;
    pa.n_cells = n_cells;
    pa.mm = ls->mm;
    pa.c = ls->c;
    dim3 grid((ls->n_act + 127) / 128, n_cells);
    SR_LAUNCH(k_line_cell_params, grid, 128, 0, st, pa);
    return SR_OK;
}

int sr_gcoeff_cells_dev(sr_lineset* ls, const double* pt_host, int n_cells, double* out_dev,
                        void* stream) {
    if (!ls || !pt_host || n_cells < 0 || !out_dev)
        return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_dev: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n_cells; i++)
        if (!(pt_host[2 * i] >= 0.0) || !(pt_host[2 * i + 1] > 0.0))
            return sr::fail(SR_ERR_ARG, "cell %d: P=%g hPa T=%g K", i, pt_host[2 * i],
                            pt_host[2 * i + 1]);
    const size_t cell_elems = (size_t)ls->n_sets * 3 * ls->n_grid;
    if (ls->n_act == 0) {
        SR_CUDA(cudaMemsetAsync(out_dev, 0, cell_elems * n_cells * sizeof(double), st));
        return SR_OK;
    }
    int dev = 0, smem_max = 0;
    SR_CUDA(cudaGetDevice(&dev));
    SR_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int ppt = pick_ppt(ls->n_sets, ls->n_groups, (size_t)smem_max);
    if (ppt == 0)
        return sr::fail(SR_ERR_LIMIT, "n_sets = %d needs more than %d bytes of shared memory",
                        ls->n_sets, smem_max);
    for (int c0 = 0; c0 < n_cells; c0 += ls->max_cells_per_batch) {
        const int nb = std::min(ls->max_cells_per_batch, n_cells - c0);
        if (c0 > 0) SR_CUDA(cudaStreamSynchronize(st));  // rec/pt buffers are reused per batch
        int code = run_params(ls, pt_host + 2 * c0, nb, st);
        if (code) return code;
        TileArgs ta;
        ta.rec = ls->rec.p;
        ta.nu0 = ls->freq.p;
        ta.gc = ls->gc.p;
        ta.ind = ls->ind.p;
        ta.grp_begin = ls->grp_begin.p;
        ta.grp_up = ls->grp_up.p;
        ta.grp_lo = ls->grp_lo.p;
        ta.lin = ls->lin.p;
        ta.out = out_dev + (size_t)c0 * cell_elems;
        ta.n_grid = ls->n_grid;
        ta.n_lines = ls->n_act;
        ta.n_sets = ls->n_sets;
        ta.n_groups = ls->n_groups;
        if (ppt == 4) code = launch_tile<4>(ta, nb, tile_smem<4>(ls->n_sets, ls->n_groups), st);
        else if (ppt == 2) code = launch_tile<2>(ta, nb, tile_smem<2>(ls->n_sets, ls->n_groups), st);
        else code = launch_tile<1>(ta, nb, tile_smem<1>(ls->n_sets, ls->n_groups), st);
        if (code) return code;
    }
    return SR_OK;
}

// sync + translate the device flag word into a status (humliv_bb's STOP conditions etc.)
static int check_flags(sr_lineset* ls, cudaStream_t st) {
    int f = 0;
    SR_CUDA(cudaMemcpyAsync(&f, ls->flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SR_CUDA(cudaStreamSynchronize(st));
    if (f) SR_CUDA(cudaMemsetAsync(ls->flags.p, 0, sizeof(int), st));
    return flags_to_status(f);
}

int sr_lineset_check(sr_lineset* ls, void* stream) {
    if (!ls) return sr::fail(SR_ERR_ARG, "sr_lineset_check: bad argument");
    return check_flags(ls, (cudaStream_t)stream);
}

int sr_gcoeff_cells_host(sr_lineset* ls, const double* pt_host, int n_cells, double* out_host) {
    if (!ls || !out_host) return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_host: bad argument");
    const size_t cell_elems = (size_t)ls->n_sets * 3 * ls->n_grid;
    // stream the result out in slabs of cells so the device buffer stays bounded (~4 GiB)
    const int slab = (int)std::max<size_t>(1, std::min<size_t>(n_cells, ((size_t)4 << 30) / (cell_elems * sizeof(double))));
    sr::DevBuf<double> buf;
    SR_CUDA(buf.alloc(cell_elems * slab));
    for (int c0 = 0; c0 < n_cells; c0 += slab) {
        const int nb = std::min(slab, n_cells - c0);
        int code = sr_gcoeff_cells_dev(ls, pt_host + 2 * c0, nb, buf.p, nullptr);
        if (code) return code;
        code = check_flags(ls, 0);
        if (code) return code;
        SR_CUDA(cudaMemcpy(out_host + (size_t)c0 * cell_elems, buf.p,
                           cell_elems * nb * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return SR_OK;
}

int sr_gcoeff_cells_dev_f32(sr_lineset* ls, const double* pt_host, int n_cells, float* out32_dev,
                            double* scratch_dev, void* stream) {
    if (!ls || !out32_dev || !scratch_dev)
        return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_dev_f32: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cell_elems = (size_t)ls->n_sets * 3 * ls->n_grid;
    // scratch_dev holds ONE cell in FP64; cells are converted one by one
    for (int c = 0; c < n_cells; c++) {
        int code = sr_gcoeff_cells_dev(ls, pt_host + 2 * c, 1, scratch_dev, stream);
        if (code) return code;
        SR_LAUNCH(k_f64_to_f32, 148 * 8, 256, 0, st, scratch_dev,
                  out32_dev + (size_t)c * cell_elems, cell_elems);
    }
    return SR_OK;
}

int sr_line_shapes_dev(sr_lineset* ls, double pres_hpa, double temp, double* shapes_dev,
                       double* g_dev, void* stream) {
    if (!ls || !shapes_dev || !g_dev) return sr::fail(SR_ERR_ARG, "sr_line_shapes_dev: bad argument");
    if (ls->n_act == 0) return SR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    double pt[2] = {pres_hpa, temp};
    SR_CUDA(cudaStreamSynchronize(st));
    int code = run_params(ls, pt, 1, st);
    if (code) return code;
    SR_CUDA(ls->facs.ensure(ls->n_act));
    SR_LAUNCH(k_line_fac, (ls->n_act + 255) / 256, 256, 0, st, ls->freq.p, ls->n_act, temp,
              ls->mm, ls->c, ls->facs.p);
    SR_LAUNCH(k_line_shapes, ls->n_act, 256, 0, st, ls->rec.p, ls->freq.p, ls->gc.p, ls->lin.p,
              shapes_dev, g_dev, ls->facs.p);
    SR_CUDA(cudaStreamSynchronize(st));  // pt is a stack variable
    return check_flags(ls, st);
}

}  // extern "C"
