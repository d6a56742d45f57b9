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
//   line table   SoA doubles/ints, sorted by (group = (upper set, lower set), centre index)
//   LineCell     [cell][line] 128 B: widths, region boundaries, G coefficients of one (line, cell)
//   LineRec      [cell][line] 112 B: the far-wing (region 1) form of the same line in ABSOLUTE grid
//                coordinates; read once per (tile, candidate), compacted into shared memory
//   core         [cell][line][CORE_STRIDE] K(x,y) of the window points il..ir (regions 2/3/4)
//   out          [cell][set][ctype][n_grid] doubles, each element written exactly once
//
// Kernels
//   k_line_cell_params  one thread per (cell, line): everything that is per line, not per point
//   k_core_eval         one warp per (cell, line): the ~250 centre points (divergent complex
//                       rational code, exp, cos) into the core buffer
//   k_voigt_tile        one CTA per (tile of TP grid points, cell): far-wing evaluation of every
//                       line whose 13010-point window touches the tile (10 FP64 ops + 1 MUFU per
//                       line*point), FP64 register accumulation per group, shared-memory tile of
//                       all n_sets*3 output rows, plus the gather of the buffered centre values
#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <type_traits>
#include <vector>
#include "sr_common.h"
#include "sr_device.cuh"

namespace {

constexpr int N_WIN = SR_IMXSIG;       // 13010
constexpr int HALF = SR_IMXSIG / 2;    // 6505: window index of the centre point (0-based)
constexpr int MAX_GROUPS = 1024;
constexpr int CORE_STRIDE = 512;       // buffered non-region-1 points per (line, cell)
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

// Far-wing record in absolute grid coordinates: for grid point P (0-based index into the
// spectral grid)  x = b{L,R} + P*xs,  u = x^2 + c1,  contribution g{0,1,2} * reg1_fast(u, c2).
//   left wing  : Pwin_lo <= P <= PL_end        right wing : PR_beg <= P <= Pwin_hi
//   centre (regions 2/3/4, from the core buffer): PL_end < P < PR_beg, core index P - PL_end - 1
struct __align__(16) LineRec {
    double xs, bL, bR, c1, c2, g0, g1, g2;   // g = gs * b4 (far wing)
    double gs0, gs1, gs2;                     // G/fac (centre values from the core buffer)
    int Pwin_lo, PL_end, PR_beg, Pwin_hi;
    int pad[2];
};
static_assert(sizeof(LineRec) == 112, "LineRec must be 112 bytes");

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
    LineRec* lrec;       // [n_cells][n_lines]
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
    if (!(il2 <= ir && ir2 >= il && il2 >= il && ir2 <= ir && il2 < N_WIN && ir2 > 1 && il <= ir))
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
    {
        // absolute-index form: window index j0 = P - (ind - HALF)
        const int w0 = a.L.ind[l] - HALF;
        LineRec f;
        f.xs = xs;
        f.bL = fma(-(double)w0, xs, r.baseL1);
        f.bR = fma(-(double)w0, xs, r.baseR1);
        f.c1 = r.c1;
        f.c2 = r.c2;
        f.g0 = r.gs[0] * r.b4;
        f.g1 = r.gs[1] * r.b4;
        f.g2 = r.gs[2] * r.b4;
        f.gs0 = r.gs[0];
        f.gs1 = r.gs[1];
        f.gs2 = r.gs[2];
        f.pad[0] = f.pad[1] = 0;
        f.Pwin_lo = w0;
        f.Pwin_hi = w0 + N_WIN - 1;
        f.PL_end = w0 + il - 2;   // last region-1-left point (== Pwin_lo - 1 when il == 1)
        f.PR_beg = w0 + ir;       // first region-1-right point (== Pwin_hi + 1 when ir == N)
        a.lrec[(size_t)cell * a.n_lines + l] = f;
    }
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

// Regions 2/3/4 of every (cell, line): one warp per line evaluates the window points il..ir
// (everything that is not pure far wing, ~250 points at Titan pressures) into the core buffer,
// so that the tile kernel never runs the divergent complex-rational code itself.
__global__ void __launch_bounds__(128) k_core_eval(const LineCell* __restrict__ rec,
                                                   const double* __restrict__ nu0,
                                                   const double* __restrict__ gc,
                                                   const double* __restrict__ lin, int n_lines,
                                                   double* __restrict__ core) {
    const int line = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int cell = blockIdx.y;
    if (line >= n_lines) return;
    const LineCell* rc = rec + (size_t)cell * n_lines + line;
    const int il = rc->il, ir = rc->ir;
    if (ir - il + 1 > CORE_STRIDE) return;   // wide centre: evaluated inline by the tile kernel
    const double f = nu0[line], g = gc[line];
    double* __restrict__ dst = core + ((size_t)cell * n_lines + line) * CORE_STRIDE;
    // same precedence as eval_window_point (last writer of lineshape.f:455-562 wins)
    const int il2 = rc->il2, ir2 = rc->ir2;
    const int core_lo = (il2 > il) ? il2 + 1 : il, core_hi = (ir2 < ir) ? ir2 - 1 : ir;
    const double ry = rc->ry, dwp = rc->dwp, xs = rc->xs;
    const double bL2 = rc->baseL2, bR2 = rc->baseR2, bL1 = rc->baseL1, bR1 = rc->baseR1;
    const double c1 = rc->c1, c2 = rc->c2, b4 = rc->b4;
    const srdev::reg2_coef k2 = srdev::humliv_reg2_coefs(ry);
    // Three index ranges with one code path each (the per-point precedence of eval_window_point
    // reduces to them: region 2 left [il, il2] when il2 > il, centre [core_lo, core_hi], region 2
    // right [ir2, ir] when ir2 < ir; region 1 only ever appears at il / ir when that side has no
    // region 2, and then the centre range covers the point).  Separate loops keep a warp on one
    // formula except where regions 3 and 4 meet inside the centre.
    const int lane = threadIdx.x & 31;
    if (il2 > il)
        for (int j1 = il + lane; j1 <= il2; j1 += 32) {
            const double x = fma((double)(j1 - 1), xs, bL2);
            dst[j1 - il] = srdev::humliv_reg2_eval(k2, x * x);
        }
    __syncwarp();   // (ranges can only overlap in degenerate geometry; keep the old precedence then)
    for (int j1 = core_lo + lane; j1 <= core_hi; j1 += 32) {
        const double x = lin[j1 - 1] + g;               // spect_classes.py:1455
        dst[j1 - il] = srdev::humliv_core_fast(fabs(x - f) / dwp, ry);   // lineshape.f:527
    }
    __syncwarp();
    if (ir2 < ir)
        for (int j1 = ir2 + lane; j1 <= ir; j1 += 32) {
            const double x = fma((double)(j1 - 1), xs, bR2);
            dst[j1 - il] = srdev::humliv_reg2_eval(k2, x * x);
        }
    (void)bL1; (void)bR1; (void)c1; (void)c2; (void)b4;
}

// ---------------------------------------------------------------------------------------------
// Far field of the far wings.  A far wing that covers a whole tile and whose line centre lies more
// than two tile lengths (+ ry + 1 Doppler widths) away is, over the TP points of the tile, a
// smooth function with its poles far outside: FAR_NN = 12 Chebyshev nodes reproduce it to 4e-11 of
// its own value (tools/far_field_check.py: all ry in [1e-4, 80], xs in [0.02, 0.5], both wings).
// k_far_nodes evaluates such records at the 12 nodes instead of the TP points, sums them per output
// row in a fixed order and leaves 12 monomial coefficients per (cell, tile, row); k_voigt_tile
// skips the same records (same predicate on the same numbers) and adds the polynomial when it
// stores the row.  84 % of the full far wings of the CH4 workload go this way.
// ---------------------------------------------------------------------------------------------
constexpr int FAR_NN = 12;
constexpr int FAR_MAX_GROUPS = 128;
__constant__ double c_far_t[FAR_NN];             // Chebyshev nodes on [-1, 1]
__constant__ double c_far_M[FAR_NN * FAR_NN];    // node values -> monomial coefficients, [power][node]

__device__ __forceinline__ bool far_full_half(double xs, double b, double c2, int tile0, int tile_last) {
    const double x0 = fma((double)tile0, xs, b), x1 = fma((double)tile_last, xs, b);
    const double d = fmin(fabs(x0), fabs(x1));
    return x0 * x1 > 0.0 && d >= 2.0 * fabs(x1 - x0) + sqrt(0.5 * c2) + 1.0;
}

struct FarArgs {
    const LineRec* lrec;     // [n_cells][n_lines]
    const int* tile_rng;     // [n_tiles][n_groups][2]
    const int* grp_up;       // [n_groups] upper set of the group
    const int* grp_lo;       // [n_groups] lower set
    const int* row_ptr;      // [n_sets*3 + 1] CSR: the groups that feed each output row,
    const int* row_grp;      //                 ascending (3 entries per group in total)
    double* coef;            // [n_cells][n_tiles_window][n_sets*3][FAR_NN]
    int n_lines, n_groups, n_sets, tile_base, tp;
};

constexpr int FAR_NT = 256;
constexpr int FAR_CAP = 512;   // candidates per pass

// One CTA per (tile, cell).  Pass over the tile's candidate list in slices of FAR_CAP: every thread
// loads one or two candidates (all loads of a slice are in flight at once: one L2 latency), flags
// the wings that qualify and leaves the record in shared memory; then thread (node n, lane q) walks
// the groups q, q+16, ... and, inside a group, the records in order - a single writer and a single
// order per (group, node), so the sums do not depend on scheduling.
// BIG (long line lists: thousands of candidates per tile, a slice then holds two or three groups):
// the run of a group inside a slice that is >= 64 candidates long is shared by all 16 lanes
// (candidate ca + q, ca + q + 16, ...), their partial sums are merged in lane order.
constexpr int FAR_BIGN = 8;    // at most 512 / 64 long runs per slice

template <bool BIG>
__global__ void __launch_bounds__(FAR_NT) k_far_nodes(FarArgs a) {
    extern __shared__ __align__(16) unsigned char fraw[];
    double* rec = reinterpret_cast<double*>(fraw);                       // [8][FAR_CAP] xs bL bR c1 c2 g0 g1 g2
    double* gsum = rec + 8 * FAR_CAP;                                    // [n_groups][3][FAR_NN]
    double* rowsum = gsum + (size_t)a.n_groups * 3 * FAR_NN;             // [n_sets*3][FAR_NN]
    double* Msm = rowsum + (size_t)a.n_sets * 3 * FAR_NN;                // [FAR_NN][FAR_NN]
    int* fl = reinterpret_cast<int*>(Msm + FAR_NN * FAR_NN);             // [FAR_CAP]
    int* cum = fl + FAR_CAP;                                             // [n_groups + 1]
    int* glo = cum + a.n_groups + 1;                                     // [n_groups]
    int* big = glo + a.n_groups;                                         // BIG: [1 + 3 * FAR_BIGN] n, g, ca, cb
    double* gpart = reinterpret_cast<double*>(
        fraw + ((reinterpret_cast<unsigned char*>(big + 1 + 3 * FAR_BIGN) - fraw + 15) / 16) * 16);
                                                                         // BIG: [FAR_BIGN][16][3][FAR_NN]
    const int cell = blockIdx.y, tile_idx = a.tile_base + blockIdx.x;
    const int tile0 = tile_idx * a.tp, tile_last = tile0 + a.tp - 1;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_rows = a.n_sets * 3;
    const LineRec* __restrict__ lrec = a.lrec + (size_t)cell * a.n_lines;
    for (int g = tid; g < a.n_groups; g += FAR_NT) {
        const int2 rg = __ldg(reinterpret_cast<const int2*>(a.tile_rng) + (size_t)tile_idx * a.n_groups + g);
        glo[g] = rg.x;
        cum[g + 1] = rg.y - rg.x;
    }
    for (int i = tid; i < a.n_groups * 3 * FAR_NN; i += FAR_NT) gsum[i] = 0.0;
    for (int i = tid; i < FAR_NN * FAR_NN; i += FAR_NT) Msm[i] = c_far_M[i];
    __syncthreads();
    if (wid == 0) {   // inclusive prefix sum of the group sizes (n_groups <= FAR_MAX_GROUPS = 128)
        int carry = 0;
        for (int g0 = 0; g0 < a.n_groups; g0 += 32) {
            const int g = g0 + lane;
            int v = g < a.n_groups ? cum[g + 1] : 0;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += t;
            }
            if (g < a.n_groups) cum[g + 1] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) cum[0] = 0;
    }
    __syncthreads();
    const int n_tot = cum[a.n_groups];
    const int node = tid & 15, q = tid >> 4;
    const double Pn = (double)tile0 + 0.5 * (double)(a.tp - 1) * (1.0 + c_far_t[node < FAR_NN ? node : 0]);
    for (int f0 = 0; f0 < n_tot; f0 += FAR_CAP) {
        const int f1 = min(n_tot, f0 + FAR_CAP);
        for (int f = f0 + tid; f < f1; f += FAR_NT) {
            int lo = 0, hi = a.n_groups;   // largest g with cum[g] <= f
            while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (cum[m] <= f) lo = m; else hi = m; }
            LineRec R;
            const int4* src = reinterpret_cast<const int4*>(lrec + glo[lo] + (f - cum[lo]));
            int4* dst = reinterpret_cast<int4*>(&R);
#pragma unroll
            for (int w = 0; w < (int)(sizeof(LineRec) / 16); w++) dst[w] = __ldg(src + w);
            const bool fL = R.PL_end >= R.Pwin_lo && R.Pwin_lo <= tile0 && R.PL_end >= tile_last;
            const bool fR = R.Pwin_hi >= R.PR_beg && R.PR_beg <= tile0 && R.Pwin_hi >= tile_last;
            int flag = 0;
            if (fL && far_full_half(R.xs, R.bL, R.c2, tile0, tile_last)) flag |= 1;
            if (fR && far_full_half(R.xs, R.bR, R.c2, tile0, tile_last)) flag |= 2;
            const int i = f - f0;
            fl[i] = flag;
            if (flag) {
                rec[0 * FAR_CAP + i] = R.xs; rec[1 * FAR_CAP + i] = R.bL; rec[2 * FAR_CAP + i] = R.bR;
                rec[3 * FAR_CAP + i] = R.c1; rec[4 * FAR_CAP + i] = R.c2; rec[5 * FAR_CAP + i] = R.g0;
                rec[6 * FAR_CAP + i] = R.g1; rec[7 * FAR_CAP + i] = R.g2;
            }
        }
        __syncthreads();
        // one candidate of the slice at this thread's node
        auto eval = [&](int i, double& s0, double& s1, double& s2) {
            const int flag = fl[i];
            if (!flag) return;
            const double xs = rec[i], c1 = rec[3 * FAR_CAP + i], c2 = rec[4 * FAR_CAP + i];
            double kp = 0.0;
            if (flag & 1) {
                const double x = fma(Pn, xs, rec[1 * FAR_CAP + i]);
                kp = srdev::humliv_reg1_u(fma(x, x, c1), c2);
            }
            if (flag & 2) {
                const double x = fma(Pn, xs, rec[2 * FAR_CAP + i]);
                kp += srdev::humliv_reg1_u(fma(x, x, c1), c2);
            }
            s0 = fma(rec[5 * FAR_CAP + i], kp, s0);
            s1 = fma(rec[6 * FAR_CAP + i], kp, s1);
            s2 = fma(rec[7 * FAR_CAP + i], kp, s2);
        };
        if (BIG) {   // the long runs of this slice
            if (tid == 0) {
                int nb = 0;
                for (int g = 0; g < a.n_groups && nb < FAR_BIGN; g++) {
                    const int ca = max(cum[g], f0), cb = min(cum[g + 1], f1);
                    if (cb - ca >= 64) { big[1 + nb] = g; big[1 + FAR_BIGN + nb] = ca; big[1 + 2 * FAR_BIGN + nb] = cb; nb++; }
                }
                big[0] = nb;
            }
            __syncthreads();
        }
        const int n_big = BIG ? big[0] : 0;
        if (node < FAR_NN)
            for (int g = q; g < a.n_groups; g += FAR_NT / 16) {
                const int ca = max(cum[g], f0), cb = min(cum[g + 1], f1);
                if (cb <= ca) continue;
                bool is_big = false;
                for (int b = 0; b < n_big; b++) is_big = is_big || big[1 + b] == g;
                if (is_big) continue;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                for (int c = ca; c < cb; c++) eval(c - f0, s0, s1, s2);
                gsum[((size_t)g * 3 + 0) * FAR_NN + node] += s0;
                gsum[((size_t)g * 3 + 1) * FAR_NN + node] += s1;
                gsum[((size_t)g * 3 + 2) * FAR_NN + node] += s2;
            }
        if (BIG && n_big > 0) {
            if (node < FAR_NN)
                for (int b = 0; b < n_big; b++) {
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                    for (int c = big[1 + FAR_BIGN + b] + q; c < big[1 + 2 * FAR_BIGN + b]; c += 16)
                        eval(c - f0, s0, s1, s2);
                    double* o = gpart + (((size_t)b * 16 + q) * 3) * FAR_NN + node;
                    o[0] = s0;
                    o[FAR_NN] = s1;
                    o[2 * FAR_NN] = s2;
                }
            __syncthreads();
            for (int item = tid; item < n_big * 3 * FAR_NN; item += FAR_NT) {   // lanes in order
                const int b = item / (3 * FAR_NN), r = item - b * 3 * FAR_NN;
                double sum = 0.0;
                for (int l = 0; l < 16; l++) sum += gpart[((size_t)b * 16 + l) * 3 * FAR_NN + r];
                gsum[(size_t)big[1 + b] * 3 * FAR_NN + r] += sum;
            }
        }
        __syncthreads();
    }
    // rows: sp / ind emission of a set = its groups as upper set, absorption = its groups as lower set
    for (int item = tid; item < n_rows * FAR_NN; item += FAR_NT) {
        const int row = item / FAR_NN, n = item - row * FAR_NN, ct = row % 3;
        double sum = 0.0;
        for (int e = __ldg(a.row_ptr + row); e < __ldg(a.row_ptr + row + 1); e++)
            sum += gsum[((size_t)__ldg(a.row_grp + e) * 3 + ct) * FAR_NN + n];
        rowsum[item] = sum;
    }
    __syncthreads();
    double* __restrict__ out = a.coef + ((size_t)cell * gridDim.x + blockIdx.x) * n_rows * FAR_NN;
    for (int item = tid; item < n_rows * FAR_NN; item += FAR_NT) {
        const int row = item / FAR_NN, j = item - row * FAR_NN;
        double c = 0.0;
#pragma unroll
        for (int n = 0; n < FAR_NN; n++) c = fma(Msm[j * FAR_NN + n], rowsum[row * FAR_NN + n], c);
        out[item] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// K1/K2 tile kernel (v5)
// ---------------------------------------------------------------------------------------------
// One far wing of one line restricted to nothing: for grid point P (absolute index)
//   x = b + P*xs,  u = x^2 + c1,  contribution g{0,1,2} * reg1_fast(u, c2)  iff  lo <= P <= lo+len
struct __align__(16) HalfRec {
    double xs, b, c1, c2, g0, g1, g2;
    int lo;
    unsigned len;
};
static_assert(sizeof(HalfRec) == 64, "HalfRec must be 64 bytes");
// Centre (regions 2/3/4) of one line: buffered K at core[line][P - le - 1] for le < P <= le + n
struct __align__(16) CentreRec {
    double gs0, gs1, gs2;
    int le, n, line, pad;
};
static_assert(sizeof(CentreRec) == 48, "CentreRec must be 48 bytes");

struct TileArgs {
    const LineCell* rec;     // [n_cells][n_lines]
    const LineRec* lrec;     // [n_cells][n_lines]
    const double* nu0;       // sorted line arrays
    const double* gc;
    const double* lin;       // [N_WIN]
    const double* core;      // [n_cells][n_lines][CORE_STRIDE] K of the points il..ir
    const int* tile_rng;     // [n_tiles][n_groups][2] lines whose window touches the tile
    const int* grp_upidx;    // [n_groups] index of the group's upper set in up_list
    const int* grp_loslot;   // [n_groups] index of the group's lower set in lo_list
    const int* up_list;      // [n_up] distinct upper sets, ascending (groups are sorted by them)
    const int* lo_list;      // [n_lo] distinct lower sets
    const int* zero_rows;    // [n_zero] rows (set*3+ctype) that no line feeds
    void* out;               // [n_cells][n_sets][3][n_grid] double (or float when F32)
    long n_grid;             // end of the output window (grid point index, exclusive)
    long row_stride;         // elements between consecutive output rows (>= window length)
    long pt_lo;              // first grid point of the output window (a multiple of the tile size)
    int tile_base;           // pt_lo / tile size: index of the window's first tile in tile_rng
    const double* far_coef;  // [n_cells][tiles of the window][n_sets*3][FAR_NN] or nullptr (k_far_nodes)
    int n_lines, n_sets, n_groups, n_up, n_lo, n_zero;
};

// Tile kernel.  One CTA owns TP = NT*PPT consecutive grid points of one cell; thread t owns the
// points tile0 + t + k*NT.  Lines are sorted by (upper set, lower set, centre index).
//   prologue   the candidate line range of every group for this tile (precomputed once per line
//              set: it depends only on the centre indices) -> flat candidate list;
//   rounds     NT candidates per round: thread i loads candidate i's LineRec, keeps the far wings
//              (left, right) and the centre that actually intersect the tile, and a block-wide
//              scan compacts them into HalfRec / CentreRec lists in shared memory;
//   compute    one flat loop over the half records of a group: 10 FP64 instructions + one
//              MUFU.RCP64H per line*point, ONE unsigned compare as window predicate, three FP64
//              register accumulators (sp_emission, ind_emission of the upper set; absorption of the
//              lower set); buffered centre values are prefetched before the wing loop;
//   rows       sp/ind accumulators live in registers until the upper set changes and are then
//              written straight to global memory; absorption accumulators are parked in a
//              thread-private shared-memory slot per lower set.  Every output element is written
//              exactly once, by exactly one thread: no atomics, no zero-fill pass.
template <int NT, int PPT, bool F32, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_voigt_tile(TileArgs a) {
    constexpr int TP = NT * PPT;
    constexpr int NW = NT / 32;
    const int cell = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile_idx = a.tile_base + blockIdx.x;
    const int tile0 = tile_idx * TP, tile_last = tile0 + TP - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* abs_s = reinterpret_cast<double*>(smem_raw);                  // [n_lo][PPT][NT]
    double* farc = abs_s + (size_t)a.n_lo * TP;                           // [n_sets*3][FAR_NN] (k_far_nodes)
    HalfRec* hbuf = reinterpret_cast<HalfRec*>(farc + (size_t)a.n_sets * 3 * FAR_NN);   // [2*NT]
    CentreRec* cbuf = reinterpret_cast<CentreRec*>(hbuf + 2 * NT);             // [NT]
    int* pre = reinterpret_cast<int*>(cbuf + NT);                              // [NT+1] h | c<<16
    int* wsum = pre + NT + 1;                                                  // [NW]
    int* cum = wsum + NW;                                                      // [n_groups+1]
    int* glo = cum + a.n_groups + 1;                                           // [n_groups]
    int* gup = glo + a.n_groups;                                               // [n_groups]
    int* gsl = gup + a.n_groups;                                               // [n_groups]

    const LineCell* __restrict__ rec = a.rec + (size_t)cell * a.n_lines;
    const LineRec* __restrict__ lrec = a.lrec + (size_t)cell * a.n_lines;
    const double* __restrict__ core = a.core + (size_t)cell * a.n_lines * CORE_STRIDE;

    // ---- prologue: candidate ranges -> prefix sums (warp 0) ------------------------------------
    {
        const int* __restrict__ rng = a.tile_rng + (size_t)tile_idx * a.n_groups * 2;
        if (a.far_coef) {   // the tile's far-field polynomials: fetched once, used when rows are stored
            const double* __restrict__ fc =
                a.far_coef + ((size_t)cell * gridDim.x + blockIdx.x) * (size_t)a.n_sets * 3 * FAR_NN;
            for (int i = tid; i < a.n_sets * 3 * FAR_NN; i += NT) farc[i] = __ldg(fc + i);
        }
        for (int g = tid; g < a.n_groups; g += NT) {
            const int2 r = __ldg(reinterpret_cast<const int2*>(rng) + g);
            glo[g] = r.x;
            cum[g + 1] = r.y - r.x;
            gup[g] = __ldg(a.grp_upidx + g);
            gsl[g] = __ldg(a.grp_loslot + g);
        }
        __syncthreads();
        if (wid == 0) {
            int carry = 0;
            for (int g0 = 0; g0 < a.n_groups; g0 += 32) {
                const int g = g0 + lane;
                int v = g < a.n_groups ? cum[g + 1] : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, v, d);
                    if (lane >= d) v += t;
                }
                if (g < a.n_groups) cum[g + 1] = carry + v;
                carry += __shfl_sync(0xffffffffu, v, 31);
            }
            if (lane == 0) cum[0] = 0;
        }
        for (int i = 0; i < a.n_lo * PPT; i++) abs_s[i * NT + tid] = 0.0;   // own slots only
        __syncthreads();
    }
    const int n_tot = cum[a.n_groups];

    double Pd[PPT], acc0[PPT], acc1[PPT], acc2[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        Pd[k] = (double)(tile0 + tid + k * NT);
        asm volatile("" : "+d"(Pd[k]));   // keep in registers (no I2F.F64 rematerialisation: XU pipe)
        acc0[k] = acc1[k] = acc2[k] = 0.0;
    }
    const int P0 = tile0 + tid;
    int cur_g = -1, cur_up = -1, stored_up = 0;
    const size_t rows_cell = (size_t)a.n_sets * 3;

    // far-field polynomial of this (cell, tile, row) at the thread's points (k_far_nodes)
    const double far_s = 2.0 / (double)(TP - 1), far_o = -fma((double)tile0, 2.0 / (double)(TP - 1), 1.0);
    auto store_row = [&](int row, const double (&v)[PPT], bool with_far = true) {
        const size_t o = ((size_t)cell * rows_cell + row) * (size_t)a.row_stride + (P0 - a.pt_lo);
        double w[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) w[k] = v[k];
        if (with_far && a.far_coef) {
            const double2* c2p = reinterpret_cast<const double2*>(farc + (size_t)row * FAR_NN);
            double c[FAR_NN];
#pragma unroll
            for (int j = 0; j < FAR_NN / 2; j++) {
                const double2 q = c2p[j];
                c[2 * j] = q.x;
                c[2 * j + 1] = q.y;
            }
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const double t = fma(Pd[k], far_s, far_o);
                double f = c[FAR_NN - 1];
#pragma unroll
                for (int j = FAR_NN - 2; j >= 0; j--) f = fma(f, t, c[j]);
                w[k] += f;
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            if ((long)P0 + k * NT < a.n_grid) {
                if (F32) __stcs(reinterpret_cast<float*>(a.out) + o + k * NT, (float)w[k]);
                else __stcs(reinterpret_cast<double*>(a.out) + o + k * NT, w[k]);
            }
        }
    };
    const double zero_v[PPT] = {};
    // write sp/ind of upper-set index ui (and zero rows of the upper sets skipped since the last)
    auto store_up = [&](int ui) {
        for (; stored_up < ui; stored_up++) {
            const int s = __ldg(a.up_list + stored_up);
            store_row(s * 3 + 0, zero_v, false);
            store_row(s * 3 + 1, zero_v, false);
        }
        const int s = __ldg(a.up_list + ui);
        store_row(s * 3 + 0, acc0);
        store_row(s * 3 + 1, acc1);
        stored_up = ui + 1;
#pragma unroll
        for (int k = 0; k < PPT; k++) acc0[k] = acc1[k] = 0.0;
    };
    auto park_abs = [&](int slot) {
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            abs_s[(slot * PPT + k) * NT + tid] += acc2[k];
            acc2[k] = 0.0;
        }
    };
    // FULL: the whole tile lies inside this wing (no predicate at all).  Partial records skip the
    // 32-point segments of this warp that lie completely outside the wing (warp-uniform branch).
    auto half_eval = [&](const HalfRec& h, auto full) {
        constexpr bool FULL = decltype(full)::value;
        const double xs = h.xs, b = h.b, c1 = h.c1, c2 = h.c2, g0 = h.g0, g1 = h.g1, g2 = h.g2;
        const int lo = h.lo;
        const unsigned len = h.len;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            bool in = true;
            if (!FULL) {
                in = (unsigned)(P0 + k * NT - lo) <= len;   // one compare
                if (!__any_sync(0xffffffffu, in)) continue;
            }
            const double x = fma(Pd[k], xs, b);
            double kp = srdev::humliv_reg1_u(fma(x, x, c1), c2);
            if (!FULL) kp = in ? kp : 0.0;
            acc0[k] = fma(g0, kp, acc0[k]);
            acc1[k] = fma(g1, kp, acc1[k]);
            acc2[k] = fma(g2, kp, acc2[k]);
        }
    };
    auto centre_load = [&](const CentreRec& c, double (&v)[PPT]) {
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int d = P0 + k * NT - c.le - 1;
            v[k] = 0.0;
            if ((unsigned)d < (unsigned)c.n) {
                if (c.n <= CORE_STRIDE) v[k] = __ldg(core + (size_t)c.line * CORE_STRIDE + d);
                else {
                    const int w0 = __ldg(&lrec[c.line].Pwin_lo);
                    v[k] = eval_window_point(rec + c.line, P0 + k * NT - w0 + 1, a.nu0[c.line],
                                             a.gc[c.line], a.lin);
                }
            }
        }
    };
    auto centre_add = [&](const CentreRec& c, const double (&v)[PPT]) {
        const double s0 = c.gs0, s1 = c.gs1, s2 = c.gs2;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            acc0[k] = fma(s0, v[k], acc0[k]);
            acc1[k] = fma(s1, v[k], acc1[k]);
            acc2[k] = fma(s2, v[k], acc2[k]);
        }
    };

    int g_run = 0;   // first group that may still have candidates (uniform)
    for (int r0 = 0; r0 < n_tot; r0 += NT) {
        const int r1 = min(n_tot, r0 + NT);
        // ---- loader: candidate r0+tid -> half / centre records ------------------------------
        {
            const int f = r0 + tid;
            int packed = 0, nL = 0, nR = 0, nC = 0, fL = 0, fR = 0, line = 0;
            LineRec R;
            if (f < r1) {
                int lo = g_run, hi = a.n_groups;   // largest g with cum[g] <= f
                while (hi - lo > 1) { const int m = (lo + hi) >> 1; if (cum[m] <= f) lo = m; else hi = m; }
                line = glo[lo] + (f - cum[lo]);
                const int4* src = reinterpret_cast<const int4*>(lrec + line);
                int4* dst = reinterpret_cast<int4*>(&R);
#pragma unroll
                for (int q = 0; q < (int)(sizeof(LineRec) / 16); q++) dst[q] = __ldg(src + q);
                nL = (R.PL_end >= R.Pwin_lo && R.Pwin_lo <= tile_last && R.PL_end >= tile0) ? 1 : 0;
                nR = (R.Pwin_hi >= R.PR_beg && R.PR_beg <= tile_last && R.Pwin_hi >= tile0) ? 1 : 0;
                nC = (R.PR_beg - R.PL_end > 1 && R.PL_end < tile_last && R.PR_beg > tile0) ? 1 : 0;
                fL = nL && R.Pwin_lo <= tile0 && R.PL_end >= tile_last;
                fR = nR && R.PR_beg <= tile0 && R.Pwin_hi >= tile_last;
                if (a.far_coef) {   // full wings of distant lines come in through k_far_nodes
                    if (fL && far_full_half(R.xs, R.bL, R.c2, tile0, tile_last)) nL = fL = 0;
                    if (fR && far_full_half(R.xs, R.bR, R.c2, tile0, tile_last)) nR = fR = 0;
                }
                // fields: full halves | partial halves << 11 | centres << 22
                packed = (fL + fR) | ((nL + nR - fL - fR) << 11) | (nC << 22);
            }
            int v = packed;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += t;
            }
            if (lane == 31) wsum[wid] = v;
            __syncthreads();   // also: every warp is done reading the previous round's records
            int base = 0;
#pragma unroll
            for (int w = 0; w < NW; w++) base += (w < wid) ? wsum[w] : 0;
            const int ex = base + v - packed;
            pre[tid] = ex;
            if (tid == NT - 1) pre[NT] = ex + packed;
            // full halves fill hbuf from the front, partial halves from the back
            int pf = ex & 0x7ff, pp = 2 * NT - 1 - ((ex >> 11) & 0x7ff);
            const int pc = ex >> 22;
            if (nL) {
                HalfRec h;
                h.xs = R.xs; h.b = R.bL; h.c1 = R.c1; h.c2 = R.c2; h.g0 = R.g0; h.g1 = R.g1; h.g2 = R.g2;
                h.lo = R.Pwin_lo; h.len = (unsigned)(R.PL_end - R.Pwin_lo);
                if (fL) hbuf[pf++] = h; else hbuf[pp--] = h;
            }
            if (nR) {
                HalfRec h;
                h.xs = R.xs; h.b = R.bR; h.c1 = R.c1; h.c2 = R.c2; h.g0 = R.g0; h.g1 = R.g1; h.g2 = R.g2;
                h.lo = R.PR_beg; h.len = (unsigned)(R.Pwin_hi - R.PR_beg);
                if (fR) hbuf[pf] = h; else hbuf[pp] = h;
            }
            if (nC) {
                CentreRec c;
                c.gs0 = R.gs0; c.gs1 = R.gs1; c.gs2 = R.gs2;
                c.le = R.PL_end; c.n = R.PR_beg - R.PL_end - 1; c.line = line; c.pad = 0;
                cbuf[pc] = c;
                // pull the slice of buffered centre values this tile will gather towards the SM
                if (c.n <= CORE_STRIDE) {
                    const int d0 = max(tile0 - c.le - 1, 0) & ~15, d1 = min(c.n, tile_last - c.le);
                    const double* src = core + (size_t)line * CORE_STRIDE;
                    for (int d = d0; d < d1; d += 16)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(src + d));
                }
            }
            __syncthreads();
        }
        // ---- compute: the groups that have candidates in this round ---------------------------
        while (cum[g_run + 1] <= r0) g_run++;
        for (int g = g_run; g < a.n_groups && cum[g] < r1; g++) {
            const int sa = max(cum[g], r0) - r0, sb = min(cum[g + 1], r1) - r0;
            if (sb <= sa) continue;
            if (g != cur_g) {
                if (cur_g >= 0) {
                    park_abs(gsl[cur_g]);
                    if (gup[g] != cur_up) store_up(cur_up);
                }
                cur_g = g;
                cur_up = gup[g];
            }
            const int pa = pre[sa], pb = pre[sb];
            const int fa = pa & 0x7ff, fb = pb & 0x7ff;
            const int qa = (pa >> 11) & 0x7ff, qb = (pb >> 11) & 0x7ff, ca = pa >> 22, cb = pb >> 22;
            // buffered centre values: issue the (L2/HBM) loads before the wing math
            double v[2][PPT];
            const int n_c = cb - ca;
            if (n_c > 0) centre_load(cbuf[ca], v[0]);
            if (n_c > 1) centre_load(cbuf[ca + 1], v[1]);
            int h = fa;
            for (; h + 2 <= fb; h += 2) {
                half_eval(hbuf[h], std::true_type{});
                half_eval(hbuf[h + 1], std::true_type{});
            }
            if (h < fb) half_eval(hbuf[h], std::true_type{});
            for (int q = qa; q < qb; q++) half_eval(hbuf[2 * NT - 1 - q], std::false_type{});
            if (n_c > 0) centre_add(cbuf[ca], v[0]);
            if (n_c > 1) centre_add(cbuf[ca + 1], v[1]);
            for (int c = ca + 2; c < cb; c++) {
                centre_load(cbuf[c], v[0]);
                centre_add(cbuf[c], v[0]);
            }
        }
    }
    // ---- epilogue: last group, remaining upper sets, absorption rows, empty rows -----------------
    if (cur_g >= 0) {
        park_abs(gsl[cur_g]);
        store_up(cur_up);
    }
    for (; stored_up < a.n_up; stored_up++) {
        const int s = __ldg(a.up_list + stored_up);
        store_row(s * 3 + 0, zero_v, false);
        store_row(s * 3 + 1, zero_v, false);
    }
    for (int j = 0; j < a.n_lo; j++) {
        double v[PPT];
#pragma unroll
        for (int k = 0; k < PPT; k++) v[k] = abs_s[(j * PPT + k) * NT + tid];
        store_row(__ldg(a.lo_list + j) * 3 + 2, v);
    }
    for (int j = 0; j < a.n_zero; j++) store_row(__ldg(a.zero_rows + j), zero_v, false);
}

// candidate line range of every (tile, group): lines whose 13010-point window touches the tile.
// Depends only on the centre indices, so it is built once per line set (and tile size).
__global__ void k_tile_ranges(const int* __restrict__ ind, const int* __restrict__ grp_begin,
                              int n_groups, int n_tiles, int tp, int* __restrict__ rng) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles * n_groups) return;
    const int t = i / n_groups, g = i % n_groups;
    const long tile0 = (long)t * tp;
    const long t_lo = tile0 - (HALF - 1);        // window reaches the tile: ind + HALF - 1 >= tile0
    const long t_hi = tile0 + tp + HALF;         // window starts beyond the tile: ind - HALF >= tile0+tp
    const int gb = grp_begin[g], ge = grp_begin[g + 1];
    int lo = gb, hi = ge;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (ind[m] < t_lo) lo = m + 1; else hi = m; }
    const int first = lo;
    hi = ge;
    while (lo < hi) { const int m = (lo + hi) >> 1; if (ind[m] < t_hi) lo = m + 1; else hi = m; }
    rng[2 * i] = first;
    rng[2 * i + 1] = lo;
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
    // per-batch tables, double-buffered: the per-line prologue (k_line_cell_params, k_core_eval)
    // of batch i+1 runs on `aux` while the tile kernel of batch i runs on the caller's stream
    sr::DevBuf<LineCell> rec_b[2];
    sr::DevBuf<LineRec> lrec_b[2];
    sr::DevBuf<double> pt_b[2], core_b[2], far_b[2], facs;
    bool far_ok = false;      // far-field path usable (group count, constant tables uploaded)
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_entry = nullptr, ev_core[2] = {nullptr, nullptr}, ev_tile[2] = {nullptr, nullptr};
    ~sr_lineset() {
        if (aux) cudaStreamDestroy(aux);
        if (ev_entry) cudaEventDestroy(ev_entry);
        for (int b = 0; b < 2; b++) {
            if (ev_core[b]) cudaEventDestroy(ev_core[b]);
            if (ev_tile[b]) cudaEventDestroy(ev_tile[b]);
        }
    }
    sr::DevBuf<int> grp_upidx, grp_loslot, up_list, lo_list, zero_rows, tile_rng, far_rowptr, far_rowgrp;
    long ind_span = 1;        // centre-index span of the active lines (line density for k_far_nodes)
    int n_up = 0, n_lo = 0, n_zero_rows = 0;
    int cfg = 0, tile_nt = 256, tile_ppt = 4, n_tiles = 0;   // tile geometry of the range table
    std::vector<int> order;    // sorted position -> input line
    std::vector<int> ind_in;   // input line -> centre index (-1 dropped)
    int max_cells_per_batch = 1;
};

namespace {

size_t tile_smem(int nt, int ppt, int n_lo, int n_groups, int n_sets) {
    return (size_t)n_lo * nt * ppt * sizeof(double) + (size_t)n_sets * 3 * FAR_NN * sizeof(double) +
           (size_t)2 * nt * sizeof(HalfRec) +
           (size_t)nt * sizeof(CentreRec) +
           (size_t)(nt + 1 + nt / 32 + 4 * n_groups + 1) * sizeof(int) + 16;
}

template <int NT, int PPT, bool F32, int MINB>
int launch_tile(const TileArgs& ta, int n_cells, cudaStream_t st) {
    const size_t smem = tile_smem(NT, PPT, ta.n_lo, ta.n_groups, ta.n_sets);
    SR_CUDA(cudaFuncSetAttribute(k_voigt_tile<NT, PPT, F32, MINB>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int TP = NT * PPT;
    if (ta.pt_lo % TP)
        return sr::fail(SR_ERR_ARG, "output window must start on a multiple of %d grid points", TP);
    TileArgs tw = ta;
    tw.tile_base = (int)(ta.pt_lo / TP);
    dim3 grid((unsigned)((ta.n_grid - ta.pt_lo + TP - 1) / TP), (unsigned)n_cells);
    SR_LAUNCH((k_voigt_tile<NT, PPT, F32, MINB>), grid, NT, smem, st, tw);
    return SR_OK;
}

// tile geometries (threads, points per thread): the first whose absorption rows fit in shared
// memory twice per SM is used; SR_K1_CFG=<index> forces one (tuning aid)
struct TileCfg { int nt, ppt; };
constexpr TileCfg kTileCfgs[] = {{128, 4}, {256, 4}, {256, 2}, {128, 2}, {256, 1}, {128, 4}, {64, 8}, {64, 4},
                                 {128, 8}, {128, 8}, {128, 6}, {128, 4}, {128, 3}, {128, 3}, {128, 2},
                                 {128, 4}};
constexpr int kNumCfgs = (int)(sizeof(kTileCfgs) / sizeof(kTileCfgs[0]));

int pick_cfg(int n_lo, int n_groups, int n_sets, size_t smem_max) {
    if (const char* e = getenv("SR_K1_CFG")) {
        const int i = atoi(e);
        if (i >= 0 && i < kNumCfgs &&
            tile_smem(kTileCfgs[i].nt, kTileCfgs[i].ppt, n_lo, n_groups, n_sets) <= smem_max)
            return i;
    }
    if (2 * (tile_smem(kTileCfgs[5].nt, kTileCfgs[5].ppt, n_lo, n_groups, n_sets) + 1024) <= (size_t)228 * 1024)
        return 5;                    // measured best on B200 (tools/tune.py k1): 128 thr x 4 pts, 4 CTAs/SM
    for (int i = 0; i < 5; i++)      // at least two CTAs per SM
        if (2 * (tile_smem(kTileCfgs[i].nt, kTileCfgs[i].ppt, n_lo, n_groups, n_sets) + 1024) <= (size_t)228 * 1024)
            return i;
    for (int i = 0; i < 5; i++)      // one CTA per SM
        if (tile_smem(kTileCfgs[i].nt, kTileCfgs[i].ppt, n_lo, n_groups, n_sets) <= smem_max) return i;
    return -1;
}

template <bool F32>
int launch_cfg(int cfg, const TileArgs& ta, int n_cells, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_tile<128, 4, F32, 3>(ta, n_cells, st);
        case 1: return launch_tile<256, 4, F32, 2>(ta, n_cells, st);
        case 2: return launch_tile<256, 2, F32, 2>(ta, n_cells, st);
        case 3: return launch_tile<128, 2, F32, 3>(ta, n_cells, st);
        case 4: return launch_tile<256, 1, F32, 2>(ta, n_cells, st);
        case 5: return launch_tile<128, 4, F32, 4>(ta, n_cells, st);
        case 6: return launch_tile<64, 8, F32, 4>(ta, n_cells, st);
        case 7: return launch_tile<64, 4, F32, 6>(ta, n_cells, st);
        case 8: return launch_tile<128, 8, F32, 4>(ta, n_cells, st);
        case 9: return launch_tile<128, 8, F32, 3>(ta, n_cells, st);
        case 10: return launch_tile<128, 6, F32, 4>(ta, n_cells, st);
        case 11: return launch_tile<128, 4, F32, 5>(ta, n_cells, st);
        case 12: return launch_tile<128, 3, F32, 5>(ta, n_cells, st);
        case 13: return launch_tile<128, 3, F32, 6>(ta, n_cells, st);
        case 14: return launch_tile<128, 2, F32, 8>(ta, n_cells, st);
        case 15: return launch_tile<128, 4, F32, 6>(ta, n_cells, st);
    }
    return sr::fail(SR_ERR_ARG, "bad tile configuration");
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
        ls->ind_span = (long)*std::max_element(ind.begin(), ind.end()) - *std::min_element(ind.begin(), ind.end()) + 1;
        // sort by (group, centre index, input position)
        std::vector<int> perm(n_act);
        std::iota(perm.begin(), perm.end(), 0);
        auto key = [&](int i) {
            return (long long)lines->up_set[act[i]] * n_sets + lines->lo_set[act[i]];
        };
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
        int rc;
        if ((rc = up_sorted(lines->freq, ls->freq))) return rc;
        if ((rc = up_sorted(lines->a_coeff, ls->a_coeff))) return rc;
        if ((rc = up_sorted(lines->air_broad, ls->air))) return rc;
        if ((rc = up_sorted(lines->t_dep, ls->tdep))) return rc;
        if ((rc = up_sorted(lines->e_lower, ls->e_lower))) return rc;
        if ((rc = up_sorted(lines->g_up, ls->g_up))) return rc;
        if ((rc = up_sorted(lines->g_lo, ls->g_lo))) return rc;
        if ((rc = up_sorted(lines->e_vib_up, ls->evu))) return rc;
        if ((rc = up_sorted(lines->e_vib_lo, ls->evl))) return rc;
        SR_CUDA(ls->ind.upload(sind.data(), n_act, st));
        SR_CUDA(ls->grp_begin.upload(gb.data(), gb.size(), st));
        SR_CUDA(ls->grp_up.upload(gu.data(), gu.size(), st));
        SR_CUDA(ls->grp_lo.upload(gl.data(), gl.size(), st));
        {
            // distinct upper / lower sets (groups are sorted by upper set, then lower set) and the
            // rows no line feeds
            std::vector<int> up_list, lo_list, upidx(ls->n_groups), loslot(ls->n_groups), zrows;
            std::vector<int> lo_of(n_sets, -1), up_seen(n_sets, 0);
            for (int g = 0; g < ls->n_groups; g++) {
                if (up_list.empty() || up_list.back() != gu[g]) up_list.push_back(gu[g]);
                upidx[g] = (int)up_list.size() - 1;
                up_seen[gu[g]] = 1;
                if (lo_of[gl[g]] < 0) { lo_of[gl[g]] = (int)lo_list.size(); lo_list.push_back(gl[g]); }
                loslot[g] = lo_of[gl[g]];
            }
            for (int s2 = 0; s2 < n_sets; s2++) {
                if (!up_seen[s2]) { zrows.push_back(s2 * 3 + 0); zrows.push_back(s2 * 3 + 1); }
                if (lo_of[s2] < 0) zrows.push_back(s2 * 3 + 2);
            }
            ls->n_up = (int)up_list.size();
            ls->n_lo = (int)lo_list.size();
            ls->n_zero_rows = (int)zrows.size();
            {   // groups per output row (far-field row sums), ascending group order
                const int n_rows = n_sets * 3;
                std::vector<int> rp(n_rows + 1, 0), rg;
                for (int g = 0; g < ls->n_groups; g++) { rp[gu[g] * 3 + 0 + 1]++; rp[gu[g] * 3 + 1 + 1]++; rp[gl[g] * 3 + 2 + 1]++; }
                for (int r = 0; r < n_rows; r++) rp[r + 1] += rp[r];
                rg.resize(rp[n_rows]);
                std::vector<int> fillp(rp.begin(), rp.end() - 1);
                for (int g = 0; g < ls->n_groups; g++) {
                    rg[fillp[gu[g] * 3 + 0]++] = g;
                    rg[fillp[gu[g] * 3 + 1]++] = g;
                    rg[fillp[gl[g] * 3 + 2]++] = g;
                }
                SR_CUDA(ls->far_rowptr.upload(rp.data(), rp.size(), st));
                if (!rg.empty()) SR_CUDA(ls->far_rowgrp.upload(rg.data(), rg.size(), st));
            }
            SR_CUDA(ls->grp_upidx.upload(upidx.data(), upidx.size(), st));
            SR_CUDA(ls->grp_loslot.upload(loslot.data(), loslot.size(), st));
            SR_CUDA(ls->up_list.upload(up_list.data(), up_list.size(), st));
            SR_CUDA(ls->lo_list.upload(lo_list.data(), lo_list.size(), st));
            if (ls->n_zero_rows) SR_CUDA(ls->zero_rows.upload(zrows.data(), zrows.size(), st));
            SR_CUDA(cudaStreamSynchronize(st));
        }
        // ind / gc in sorted order
        SR_LAUNCH(k_closest_grid, (n_act + 255) / 256, 256, 0, st, ls->grid.p, n_grid,
                  ls->freq.p, n_act, ls->ind.p, ls->gc.p);
        // tile geometry + candidate line ranges per (tile, group)
        int smem_max = 0;
        SR_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin,
                                       ls->device));
        ls->cfg = pick_cfg(ls->n_lo, ls->n_groups, ls->n_sets, (size_t)smem_max);
        if (ls->cfg < 0)
            return sr::fail(SR_ERR_LIMIT, "%d lower levels x %d level pairs need more than %d bytes "
                            "of shared memory per tile", ls->n_lo, ls->n_groups, smem_max);
        ls->tile_nt = kTileCfgs[ls->cfg].nt;
        ls->tile_ppt = kTileCfgs[ls->cfg].ppt;
        {   // far-field tables: Chebyshev nodes and the node-values -> monomial-coefficients matrix
            double t[FAR_NN], M[FAR_NN * FAR_NN], T[FAR_NN][FAR_NN], C2P[FAR_NN][FAR_NN];
            for (int n = 0; n < FAR_NN; n++) t[n] = cos((2 * n + 1) * M_PI / (2.0 * FAR_NN));
            for (int j = 0; j < FAR_NN; j++)        // Chebyshev coefficient j from the node values
                for (int n = 0; n < FAR_NN; n++)
                    T[j][n] = (j == 0 ? 1.0 : 2.0) / FAR_NN * cos(j * (2 * n + 1) * M_PI / (2.0 * FAR_NN));
            for (int j = 0; j < FAR_NN; j++)        // monomial coefficients of T_j: T_{j+1} = 2 x T_j - T_{j-1}
                for (int p = 0; p < FAR_NN; p++) C2P[p][j] = 0.0;
            C2P[0][0] = 1.0;
            if (FAR_NN > 1) C2P[1][1] = 1.0;
            for (int j = 2; j < FAR_NN; j++)
                for (int p = 0; p < FAR_NN; p++)
                    C2P[p][j] = (p > 0 ? 2.0 * C2P[p - 1][j - 1] : 0.0) - C2P[p][j - 2];
            for (int p = 0; p < FAR_NN; p++)
                for (int n = 0; n < FAR_NN; n++) {
                    double v = 0.0;
                    for (int j = 0; j < FAR_NN; j++) v += C2P[p][j] * T[j][n];
                    M[p * FAR_NN + n] = v;
                }
            SR_CUDA(cudaMemcpyToSymbolAsync(c_far_t, t, sizeof(t), 0, cudaMemcpyHostToDevice, st));
            SR_CUDA(cudaMemcpyToSymbolAsync(c_far_M, M, sizeof(M), 0, cudaMemcpyHostToDevice, st));
            SR_CUDA(cudaStreamSynchronize(st));      // t and M are stack variables
            ls->far_ok = ls->n_groups <= FAR_MAX_GROUPS;
        }
        const int tp = ls->tile_nt * ls->tile_ppt;
        ls->n_tiles = (int)((n_grid + tp - 1) / tp);
        SR_CUDA(ls->tile_rng.alloc((size_t)ls->n_tiles * ls->n_groups * 2));
        {
            const int n = ls->n_tiles * ls->n_groups;
            SR_LAUNCH(k_tile_ranges, (n + 255) / 256, 256, 0, st, ls->ind.p, ls->grp_begin.p,
                      ls->n_groups, ls->n_tiles, tp, ls->tile_rng.p);
        }
        SR_CUDA(cudaStreamSynchronize(st));
        return SR_OK;
    };
    int code = body();
    if (code != SR_OK) { delete ls; return code; }
    // cells per batch: keep the per-(line,cell) tables under ~2 GiB (8 cells per launch already
    // fill the machine: 2344 tiles x 8 cells on 148 SMs x 4 CTAs)
    size_t per_cell = (size_t)std::max(n_act, 1) *
                      (sizeof(LineCell) + sizeof(LineRec) + CORE_STRIDE * sizeof(double));
    ls->max_cells_per_batch =
        (int)std::max<size_t>(1, std::min<size_t>(4096, ((size_t)2 << 30) / per_cell));
    if (const char* e = getenv("SR_K2_BATCH"))   // test aid: cells per batch
        ls->max_cells_per_batch = std::max(1, atoi(e));
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

static int run_params(sr_lineset* ls, const double* pt_host, int n_cells, cudaStream_t st,
                      int buf = 0) {
    // the per-batch tables are allocated once for a full batch: later (larger) calls never
    // reallocate them, so consecutive batches need no host synchronisation (stream order suffices)
    const size_t cap = (size_t)std::max(n_cells, std::min(ls->max_cells_per_batch, 16));
    SR_CUDA(ls->pt_b[buf].ensure(2 * cap));
    SR_CUDA(cudaMemcpyAsync(ls->pt_b[buf].p, pt_host, sizeof(double) * 2 * n_cells,
                            cudaMemcpyHostToDevice, st));
    SR_CUDA(ls->rec_b[buf].ensure(cap * ls->n_act));
    SR_CUDA(ls->lrec_b[buf].ensure(cap * ls->n_act));
    ParamsArgs pa;
    pa.L = {ls->freq.p, ls->a_coeff.p, ls->air.p, ls->tdep.p, ls->e_lower.p, ls->g_up.p,
            ls->g_lo.p, ls->evu.p, ls->evl.p, ls->gc.p, ls->ind.p};
    pa.lin = ls->lin.p;
    pa.pt = ls->pt_b[buf].p;
    pa.rec = ls->rec_b[buf].p;
    pa.lrec = ls->lrec_b[buf].p;
    pa.flags = ls->flags.p;
    pa.n_lines = ls->n_act;
    pa.n_cells = n_cells;
    pa.mm = ls->mm;
    pa.c = ls->c;
    dim3 grid((ls->n_act + 127) / 128, n_cells);
    SR_LAUNCH(k_line_cell_params, grid, 128, 0, st, pa);
    return SR_OK;
}

static int gcoeff_cells_impl(sr_lineset* ls, const double* pt_host, int n_cells, void* out_dev,
                             bool f32, cudaStream_t st, long row_stride = 0, long win0 = 0,
                             long win_n = -1) {
    if (win_n < 0) win_n = ls->n_grid - win0;
    if (win0 < 0 || win_n < 1 || win0 + win_n > ls->n_grid)
        return sr::fail(SR_ERR_ARG, "output window [%ld,%ld) outside the grid of %ld points", win0,
                        win0 + win_n, ls->n_grid);
    if (row_stride == 0) row_stride = win_n;
    if (row_stride < win_n)
        return sr::fail(SR_ERR_ARG, "row stride %ld < window length %ld", row_stride, win_n);
    for (int i = 0; i < n_cells; i++)
        if (!(pt_host[2 * i] >= 0.0) || !(pt_host[2 * i + 1] > 0.0))
            return sr::fail(SR_ERR_ARG, "cell %d: P=%g hPa T=%g K", i, pt_host[2 * i],
                            pt_host[2 * i + 1]);
    const size_t cell_elems = (size_t)ls->n_sets * 3 * row_stride;
    const size_t esz = f32 ? sizeof(float) : sizeof(double);
    if (ls->n_act == 0) {
        SR_CUDA(cudaMemsetAsync(out_dev, 0, cell_elems * n_cells * esz, st));
        return SR_OK;
    }
    // Two-stream pipeline over sub-batches of cells: params + core evaluation of sub-batch i+1 on
    // the (high-priority) aux stream under the tile kernel of sub-batch i on the caller's stream.
    // k_core_eval CTAs have the footprint of one tile CTA (128 threads, <= 128 registers), so they
    // slot in wherever a tile CTA retires and fill the FP64 issue slots the tile kernel's
    // record-loading phases leave idle.
    static const int pipe_on = getenv("SR_K2_PIPE") ? atoi(getenv("SR_K2_PIPE")) : 0;
    static const int sub_env = getenv("SR_K2_SUB") ? atoi(getenv("SR_K2_SUB")) : 4;
    const bool pipe = pipe_on != 0 && n_cells > 1;
    const int sub = pipe ? std::max(1, std::min(ls->max_cells_per_batch, sub_env))
                         : ls->max_cells_per_batch;
    if (pipe && !ls->aux) {
        int lo_pri = 0, hi_pri = 0;
        SR_CUDA(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        SR_CUDA(cudaStreamCreateWithPriority(&ls->aux, cudaStreamNonBlocking, hi_pri));
        SR_CUDA(cudaEventCreateWithFlags(&ls->ev_entry, cudaEventDisableTiming));
        for (int b = 0; b < 2; b++) {
            SR_CUDA(cudaEventCreateWithFlags(&ls->ev_core[b], cudaEventDisableTiming));
            SR_CUDA(cudaEventCreateWithFlags(&ls->ev_tile[b], cudaEventDisableTiming));
        }
    }
    if (pipe) {   // the aux stream starts where the caller's stream is now
        SR_CUDA(cudaEventRecord(ls->ev_entry, st));
        SR_CUDA(cudaStreamWaitEvent(ls->aux, ls->ev_entry, 0));
    }
    // far field of the far wings (k_far_nodes): SR_K1_FAR=0 evaluates every wing point by point
    const int far_env = getenv("SR_K1_FAR") ? atoi(getenv("SR_K1_FAR")) : 1;
    const int tp = ls->tile_nt * ls->tile_ppt;
    const long n_tiles_w = (win_n + tp - 1) / tp;
    // (the extra kernel pays when a launch has at least two waves of CTAs; a single cell on a
    // narrow slab - half a wave - is faster point by point.  SR_K1_FAR=2 forces it.)
    // long line lists (thousands of candidates per tile): the far field pays even for half a wave,
    // and k_far_nodes shares the long runs of a group among its lanes
    const double cand_per_tile =
        (double)ls->n_act * (double)(N_WIN + tp) / (double)std::max<long>(ls->ind_span, N_WIN);
    const bool far_big = cand_per_tile > 1024.0;
    const bool use_far = far_env != 0 && ls->far_ok && win0 % tp == 0 &&
                         (far_env == 2 || far_big || n_tiles_w * std::min(sub, n_cells) >= 2L * 148 * 4);
    const int far_cap_cells = std::max(std::min(sub, n_cells), std::min(ls->max_cells_per_batch, pipe ? sub : 16));
    const size_t core_cap = (size_t)std::max(std::min(sub, n_cells),
                                             std::min(ls->max_cells_per_batch, pipe ? sub : 16)) *
                            ls->n_act * CORE_STRIDE;
    auto prologue = [&](int i, cudaStream_t ps) -> int {   // params + core of sub-batch i
        const int c0 = i * sub, nb = std::min(sub, n_cells - c0), buf = pipe ? (i & 1) : 0;
        if (pipe && i >= 2) SR_CUDA(cudaStreamWaitEvent(ps, ls->ev_tile[buf], 0));   // tile(i-2) done
        int code = run_params(ls, pt_host + 2 * c0, nb, ps, buf);
        if (code) return code;
        SR_CUDA(ls->core_b[buf].ensure(core_cap));   // (a growing cudaFree synchronises)
        dim3 cgrid((ls->n_act + 3) / 4, nb);
        {
            sr::ProfScope pr(SR_PROF_VOIGT_CORE, 0.0, ps);
            SR_LAUNCH(k_core_eval, cgrid, 128, 0, ps, ls->rec_b[buf].p, ls->freq.p, ls->gc.p,
                      ls->lin.p, ls->n_act, ls->core_b[buf].p);
        }
        if (use_far) {   // node sums of the distant full wings, per (cell, tile of the window, row)
            const int n_rows = ls->n_sets * 3;
            SR_CUDA(ls->far_b[buf].ensure((size_t)far_cap_cells * n_tiles_w * n_rows * FAR_NN));
            FarArgs fa;
            fa.lrec = ls->lrec_b[buf].p;
            fa.tile_rng = ls->tile_rng.p;
            fa.grp_up = ls->grp_up.p;
            fa.grp_lo = ls->grp_lo.p;
            fa.row_ptr = ls->far_rowptr.p;
            fa.row_grp = ls->far_rowgrp.p;
            fa.coef = ls->far_b[buf].p;
            fa.n_lines = ls->n_act;
            fa.n_groups = ls->n_groups;
            fa.n_sets = ls->n_sets;
            fa.tile_base = (int)(win0 / tp);
            fa.tp = tp;
            const size_t fsmem = (8 * (size_t)FAR_CAP + ((size_t)ls->n_groups * 3 + n_rows) * FAR_NN +
                                  FAR_NN * FAR_NN) * sizeof(double) +
                                 ((size_t)FAR_CAP + 2 * ls->n_groups + 1) * sizeof(int) + 16;
            if (far_big) {
                const size_t bsmem = fsmem + (1 + 3 * FAR_BIGN) * sizeof(int) + 16 +
                                     (size_t)FAR_BIGN * 16 * 3 * FAR_NN * sizeof(double);
                SR_CUDA(cudaFuncSetAttribute(k_far_nodes<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
                SR_LAUNCH(k_far_nodes<true>, dim3((unsigned)n_tiles_w, (unsigned)nb), FAR_NT, bsmem, ps, fa);
            } else {
                SR_CUDA(cudaFuncSetAttribute(k_far_nodes<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                SR_LAUNCH(k_far_nodes<false>, dim3((unsigned)n_tiles_w, (unsigned)nb), FAR_NT, fsmem, ps, fa);
            }
        }
        if (pipe) SR_CUDA(cudaEventRecord(ls->ev_core[buf], ps));
        return SR_OK;
    };
    const int n_sub = (n_cells + sub - 1) / sub;
    // line*gridpoint evaluations inside the output window, per cell (profiling only)
    double win_evals = 0.0;
    if (sr::g_prof_on.load())
        for (int ind : ls->ind_in) {
            if (ind < 0) continue;
            const long lo = std::max<long>(ind - HALF, win0), hi = std::min<long>(ind - HALF + N_WIN, win0 + win_n);
            if (hi > lo) win_evals += (double)(hi - lo);
        }
    if (pipe) { int code = prologue(0, ls->aux); if (code) return code; }
    for (int i = 0; i < n_sub; i++) {
        const int c0 = i * sub, nb = std::min(sub, n_cells - c0), buf = pipe ? (i & 1) : 0;
        if (!pipe) { int code = prologue(i, st); if (code) return code; }
        else SR_CUDA(cudaStreamWaitEvent(st, ls->ev_core[buf], 0));
        TileArgs ta;
        ta.rec = ls->rec_b[buf].p;
        ta.lrec = ls->lrec_b[buf].p;
        ta.nu0 = ls->freq.p;
        ta.gc = ls->gc.p;
        ta.lin = ls->lin.p;
        ta.core = ls->core_b[buf].p;
        ta.tile_rng = ls->tile_rng.p;
        ta.grp_upidx = ls->grp_upidx.p;
        ta.grp_loslot = ls->grp_loslot.p;
        ta.up_list = ls->up_list.p;
        ta.lo_list = ls->lo_list.p;
        ta.zero_rows = ls->zero_rows.p;
        ta.out = (char*)out_dev + (size_t)c0 * cell_elems * esz;
        ta.n_grid = win0 + win_n;
        ta.row_stride = row_stride;
        ta.pt_lo = win0;
        ta.tile_base = 0;
        ta.far_coef = use_far ? ls->far_b[buf].p : nullptr;
        ta.n_lines = ls->n_act;
        ta.n_sets = ls->n_sets;
        ta.n_groups = ls->n_groups;
        ta.n_up = ls->n_up;
        ta.n_lo = ls->n_lo;
        ta.n_zero = ls->n_zero_rows;
        // the next sub-batch's prologue is queued BEFORE this tile kernel so that its CTAs are
        // already pending when tile CTAs start to retire
        if (pipe && i + 1 < n_sub) { int code = prologue(i + 1, ls->aux); if (code) return code; }
        int code;
        {
            sr::ProfScope pr(SR_PROF_VOIGT_TILE, win_evals * (double)nb, st);
            code = f32 ? launch_cfg<true>(ls->cfg, ta, nb, st) : launch_cfg<false>(ls->cfg, ta, nb, st);
        }
        if (code) return code;
        if (pipe) SR_CUDA(cudaEventRecord(ls->ev_tile[buf], st));
    }
    return SR_OK;
}

int sr_gcoeff_cells_dev(sr_lineset* ls, const double* pt_host, int n_cells, double* out_dev,
                        void* stream) {
    if (!ls || !pt_host || n_cells < 0 || !out_dev)
        return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_dev: bad argument");
    return gcoeff_cells_impl(ls, pt_host, n_cells, out_dev, false, (cudaStream_t)stream);
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
    const int slab = (int)std::max<size_t>(
        1, std::min<size_t>(n_cells, ((size_t)4 << 30) / (cell_elems * sizeof(double))));
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
    if (!ls || !pt_host || n_cells < 0 || !out32_dev)
        return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_dev_f32: bad argument");
    (void)scratch_dev;   // the tile kernel rounds to float32 (numpy astype) in its store
    return gcoeff_cells_impl(ls, pt_host, n_cells, out32_dev, true, (cudaStream_t)stream);
}

int sr_gcoeff_cells_dev_f32_ld(sr_lineset* ls, const double* pt_host, int n_cells,
                               float* out32_dev, long row_stride, void* stream) {
    if (!ls || !pt_host || n_cells < 0 || !out32_dev)
        return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_dev_f32_ld: bad argument");
    return gcoeff_cells_impl(ls, pt_host, n_cells, out32_dev, true, (cudaStream_t)stream, row_stride);
}

int sr_gcoeff_cells_window_dev(sr_lineset* ls, const double* pt_host, int n_cells, void* out_dev,
                               int f32, long row_stride, long pt0, long n_pts, void* stream) {
    if (!ls || !pt_host || n_cells < 0 || !out_dev)
        return sr::fail(SR_ERR_ARG, "sr_gcoeff_cells_window_dev: bad argument");
    return gcoeff_cells_impl(ls, pt_host, n_cells, out_dev, f32 != 0, (cudaStream_t)stream, row_stride,
                             pt0, n_pts);
}

int sr_lineset_tile_points(const sr_lineset* ls) { return ls ? ls->tile_nt * ls->tile_ppt : 0; }

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
    SR_LAUNCH(k_line_shapes, ls->n_act, 256, 0, st, ls->rec_b[0].p, ls->freq.p, ls->gc.p, ls->lin.p,
              shapes_dev, g_dev, ls->facs.p);
    SR_CUDA(cudaStreamSynchronize(st));  // pt is a stack variable
    return check_flags(ls, st);
}

}  // extern "C"
