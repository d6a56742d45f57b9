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
//                coordinates; contiguous runs are TMA-bulk-copied into shared memory
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
#include <vector>
#include "sr_common.h"
#include "sr_device.cuh"

namespace {

constexpr int N_WIN = SR_IMXSIG;       // 13010
constexpr int HALF = SR_IMXSIG / 2;    // 6505: window index of the centre point (0-based)
constexpr int MAX_GROUPS = 1024;
constexpr int CORE_STRIDE = 512;       // buffered non-region-1 points per (line, cell)
constexpr int REC_CAP = 256;           // LineRec slots in shared memory (28 KB)
constexpr int GR = 12;                 // ints per group in the tile kernel's range table
constexpr int MAX_CHUNKS = 8;
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
    int* il_min;         // [n_cells] min il over the lines of the cell (memset 0x7f before)
    int* ir_max;         // [n_cells] max ir (memset 0 before)
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
    atomicMin(a.il_min + cell, il);
    atomicMax(a.ir_max + cell, ir);
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
__global__ void __launch_bounds__(256) k_core_eval(const LineCell* __restrict__ rec,
                                                   const double* __restrict__ nu0,
                                                   const double* __restrict__ gc,
                                                   const double* __restrict__ lin, int n_lines,
                                                   double* __restrict__ core) {
    const int line = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int cell = blockIdx.y;
    if (line >= n_lines) return;
    const LineCell* rc = rec + (size_t)cell * n_lines + line;
    const int il = rc->il, ir = rc->ir;
    if (ir - il + 1 > CORE_STRIDE) return;   // wide centre: evaluated inline by the tile kernel
    const double f = nu0[line], g = gc[line];
    double* __restrict__ dst = core + ((size_t)cell * n_lines + line) * CORE_STRIDE;
    for (int j1 = il + (threadIdx.x & 31); j1 <= ir; j1 += 32)
        dst[j1 - il] = eval_window_point(rc, j1, f, g, lin);
}

// ---------------------------------------------------------------------------------------------
// K1/K2 tile kernel
// ---------------------------------------------------------------------------------------------
struct TileArgs {
    const LineCell* rec;     // [n_cells][n_lines]
    const LineRec* lrec;     // [n_cells][n_lines]
    const double* nu0;       // sorted line arrays
    const double* gc;
    const int* ind;
    const int* grp_begin;    // [n_groups+1] offsets into the sorted line arrays
    const int* grp_up;       // [n_groups]
    const int* grp_lo;
    const double* lin;       // [N_WIN]
    const double* core;      // [n_cells][n_lines][CORE_STRIDE] K of the points il..ir
    const int* il_min;       // [n_cells]
    const int* ir_max;       // [n_cells]
    double* out;             // [n_cells][n_sets][3][n_grid]
    long n_grid;
    int n_lines, n_sets, n_groups;
    // group chunks: a CTA handles one (tile, chunk of groups) and keeps only the output rows its
    // groups feed in shared memory (compact row numbering per chunk)
    const int* chunk_gbeg;   // [n_chunks+1] group ranges
    const int* grp_slot;     // [n_groups][3] compact rows: sp, ind (upper set), abs (lower set)
    const int* chunk_rows;   // [n_chunks][max_rows] global row (set*3+ctype) or -1
    const int* chunk_shared; // [n_chunks][max_rows] 1 = row also fed by another chunk: atomic add
    int n_chunks, max_rows;  //                          onto the pre-zeroed output row
};

// mbarrier / TMA bulk-copy helpers (sm_90+ PTX; SASS: SYNCS / UBLKCP)
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Tile kernel.  One CTA owns TP = NT*PPT consecutive grid points of one cell.
//   phase 0  per group (upper set, lower set): the lines whose window touches the tile (binary
//            search on the sorted centre index) and the split
//               edgeR | wingR | centre | wingL | edgeL
//            wingR / wingL are GUARANTEED (from the cell-wide min il / max ir) to see the whole
//            tile in one far wing of the line; edge lines are cut by their window end, centre
//            lines by their own regions 2/3/4;
//   phase 1  thread g issues the TMA bulk copy (cp.async.bulk + mbarrier) of group g's LineRec
//            run into shared memory; the output tile is zeroed while the copies land;
//   phase 2  far-wing evaluation: wing runs with a branch-free fixed-trip-count loop (two lines x
//            PPT points = independent FP64 chains per thread), edge/centre runs with the same loop
//            plus integer window predicates; FP64 register accumulation per group, flushed into
//            the shared output tile when the group changes;
//   phase 3  centre gather: buffered K of regions 2/3/4 (k_core_eval) added into the tile;
//   phase 4  coalesced store of the n_sets*3 rows.
template <int NT, int PPT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_voigt_tile(TileArgs a) {
    constexpr int TP = NT * PPT;
    const int chunk = blockIdx.y % a.n_chunks, cell = blockIdx.y / a.n_chunks;
    const int gb0 = a.chunk_gbeg[chunk], n_grp = a.chunk_gbeg[chunk + 1] - gb0;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* acc_s = reinterpret_cast<double*>(smem_raw);                          // [max_rows][TP]
    LineRec* recbuf = reinterpret_cast<LineRec*>(acc_s + (size_t)a.max_rows * TP);      // [REC_CAP]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(recbuf + REC_CAP);
    int* g_rng = reinterpret_cast<int*>(mbar + 2);   // [n_grp][GR] lo a b c d hi cs ce | 3 row slots
    int* cum = g_rng + GR * n_grp;                   // [n_grp+1] slot of group start

    const int tid = threadIdx.x;
    const long tile0 = (long)blockIdx.x * TP;
    const LineCell* __restrict__ rec = a.rec + (size_t)cell * a.n_lines;
    const LineRec* __restrict__ lrec = a.lrec + (size_t)cell * a.n_lines;
    const double* __restrict__ core = a.core + (size_t)cell * a.n_lines * CORE_STRIDE;

    // ---- phase 0: ranges --------------------------------------------------------------------
    // Eight lower_bound searches per group over the sorted centre indices, run in lock step so
    // that their (L2-latency-bound) probes overlap.
    const int il_min = a.il_min[cell], ir_max = a.ir_max[cell];
    long thr[8];
    thr[0] = tile0 - (HALF - 1);                  // lo : window reaches the tile
    thr[1] = tile0 + HALF + TP - N_WIN;           // a  : window end beyond the tile (wingR start)
    thr[2] = tile0 + HALF - ir_max + 1;           // b  : wingR end  (tile start beyond every ir)
    thr[3] = tile0 + HALF + TP - il_min + 1;      // c  : wingL start (tile end before every il)
    thr[4] = tile0 + HALF + 1;                    // d  : wingL end  (window start before the tile)
    thr[5] = tile0 + TP + HALF;                   // hi : window starts beyond the tile
    thr[6] = thr[2];                              // centre values can reach the tile from here ...
    thr[7] = thr[3];                              // ... to here
    for (int g = tid; g < n_grp; g += NT) {
        const int gb = a.grp_begin[gb0 + g], ge = a.grp_begin[gb0 + g + 1];
        int lo8[8], hi8[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { lo8[q] = gb; hi8[q] = ge; }
        for (int span = ge - gb; span > 0; span >>= 1) {   // ceil(log2(n+1)) lock-step rounds
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (lo8[q] < hi8[q]) {
                    const int m = (lo8[q] + hi8[q]) >> 1;
                    if (__ldg(a.ind + m) < thr[q]) lo8[q] = m + 1; else hi8[q] = m;
                }
            }
        }
        const int lo = lo8[0], hi = max(lo8[5], lo);
        int ra = min(max(lo8[1], lo), hi), rb = min(max(lo8[2], ra), hi);
        int rc = min(max(lo8[3], rb), hi), rd = min(max(lo8[4], rc), hi);
        if (il_min <= 1) { rc = hi; rd = hi; }
        int* r = g_rng + GR * g;
        r[0] = lo; r[1] = ra; r[2] = rb; r[3] = rc; r[4] = rd; r[5] = hi;
        r[6] = min(max(lo8[6], lo), hi);
        r[7] = (il_min <= 1) ? hi : min(max(lo8[7], r[6]), hi);
        r[8] = a.grp_slot[3 * (gb0 + g)];
        r[9] = a.grp_slot[3 * (gb0 + g) + 1];
        r[10] = a.grp_slot[3 * (gb0 + g) + 2];
    }
    if (tid == 0) mbar_init(mbar, NT);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        cum[0] = 0;
        for (int g = 0; g < n_grp; g++) { run += g_rng[GR * g + 5] - g_rng[GR * g]; cum[g + 1] = run; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int n_tot = cum[n_grp];
    const int n_rounds = (n_tot + REC_CAP - 1) / REC_CAP;

    // ---- phase 1: TMA bulk copies of round r (every thread arrives exactly once per round) ----
    auto issue_round = [&](int r) {
        const int r0 = r * REC_CAP, r1 = min(n_tot, r0 + REC_CAP);
        unsigned bytes = 0;
        for (int g = tid; g < n_grp; g += NT) {
            const int s0 = max(cum[g], r0), s1 = min(cum[g + 1], r1);
            if (s1 > s0) bytes += (unsigned)(s1 - s0) * (unsigned)sizeof(LineRec);
        }
        if (bytes) mbar_arrive_expect_tx(mbar, bytes); else mbar_arrive(mbar);
        for (int g = tid; g < n_grp; g += NT) {
            const int s0 = max(cum[g], r0), s1 = min(cum[g + 1], r1);
            if (s1 > s0)
                bulk_g2s(recbuf + (s0 - r0), lrec + g_rng[GR * g] + (s0 - cum[g]),
                         (unsigned)(s1 - s0) * (unsigned)sizeof(LineRec), mbar);
        }
    };
    if (n_rounds > 0) issue_round(0);

    for (int i = tid; i < a.max_rows * TP; i += NT) acc_s[i] = 0.0;

    double Pd[PPT], acc0[PPT], acc1[PPT], acc2[PPT];
    int Pi[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        Pi[k] = (int)tile0 + tid + k * NT;
        Pd[k] = (double)Pi[k];
        acc0[k] = acc1[k] = acc2[k] = 0.0;
    }
    int cur_g = -1;
    __syncthreads();   // tile zeroed before the first flush

    auto flush = [&](int g) {
        if (g < 0) return;
        const int r0 = g_rng[GR * g + 8], r1 = g_rng[GR * g + 9], r2 = g_rng[GR * g + 10];
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int p = tid + k * NT;
            acc_s[r0 * TP + p] += acc0[k];
            acc_s[r1 * TP + p] += acc1[k];
            acc_s[r2 * TP + p] += acc2[k];
            acc0[k] = acc1[k] = acc2[k] = 0.0;
        }
    };

    // whole tile in one wing of every line of the run: no predicates, two lines per iteration
    auto wing_run = [&](const LineRec* __restrict__ fr, int n, bool right) {
        int i = 0;
        for (; i + 2 <= n; i += 2) {
            const LineRec& f0 = fr[i];
            const LineRec& f1 = fr[i + 1];
            const double xs0 = f0.xs, xs1 = f1.xs;
            const double bs0 = right ? f0.bR : f0.bL, bs1 = right ? f1.bR : f1.bL;
            const double c10 = f0.c1, c20 = f0.c2, c11 = f1.c1, c21 = f1.c2;
            const double a0 = f0.g0, a1 = f0.g1, a2 = f0.g2, b0 = f1.g0, b1 = f1.g1, b2 = f1.g2;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const double x0 = fma(Pd[k], xs0, bs0);
                const double x1 = fma(Pd[k], xs1, bs1);
                const double k0 = srdev::humliv_reg1_fast(fma(x0, x0, c10), c20);
                const double k1 = srdev::humliv_reg1_fast(fma(x1, x1, c11), c21);
                acc0[k] = fma(a0, k0, acc0[k]);
                acc1[k] = fma(a1, k0, acc1[k]);
                acc2[k] = fma(a2, k0, acc2[k]);
                acc0[k] = fma(b0, k1, acc0[k]);
                acc1[k] = fma(b1, k1, acc1[k]);
                acc2[k] = fma(b2, k1, acc2[k]);
            }
        }
        if (i < n) {
            const LineRec& f0 = fr[i];
            const double xs0 = f0.xs, bs0 = right ? f0.bR : f0.bL, c10 = f0.c1, c20 = f0.c2;
            const double a0 = f0.g0, a1 = f0.g1, a2 = f0.g2;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const double x0 = fma(Pd[k], xs0, bs0);
                const double k0 = srdev::humliv_reg1_fast(fma(x0, x0, c10), c20);
                acc0[k] = fma(a0, k0, acc0[k]);
                acc1[k] = fma(a1, k0, acc1[k]);
                acc2[k] = fma(a2, k0, acc2[k]);
            }
        }
    };
    // lines cut by their window end or by their own centre: same math, per-point predicates
    auto pred_run = [&](const LineRec* __restrict__ fr, int n) {
        for (int i = 0; i < n; i++) {
            const LineRec& f = fr[i];
            const double xs = f.xs, bL = f.bL, bR = f.bR, c1 = f.c1, c2 = f.c2;
            const double a0 = f.g0, a1 = f.g1, a2 = f.g2;
            const int wlo = f.Pwin_lo, le = f.PL_end, rb = f.PR_beg, whi = f.Pwin_hi;
#pragma unroll
            for (int k = 0; k < PPT; k++) {
                const int P = Pi[k];
                const bool inL = P >= wlo && P <= le, inR = P >= rb && P <= whi;
                if (inL || inR) {
                    const double x = fma(Pd[k], xs, inL ? bL : bR);
                    const double kp = srdev::humliv_reg1_fast(fma(x, x, c1), c2);
                    acc0[k] = fma(a0, kp, acc0[k]);
                    acc1[k] = fma(a1, kp, acc1[k]);
                    acc2[k] = fma(a2, kp, acc2[k]);
                }
            }
        }
    };
    // K of regions 2/3/4 of one line at this thread's points (0 outside PL_end < P < PR_beg)
    auto centre_load = [&](const LineRec& f, int line, double (&v)[PPT]) {
        const int le = f.PL_end, rb = f.PR_beg;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int P = Pi[k];
            v[k] = 0.0;
            if (P > le && P < rb) {
                if (rb - le - 1 <= CORE_STRIDE)
                    v[k] = __ldg(core + (size_t)line * CORE_STRIDE + (P - le - 1));
                else
                    v[k] = eval_window_point(rec + line, P - f.Pwin_lo + 1, a.nu0[line],
                                             a.gc[line], a.lin);
            }
        }
    };
    auto centre_add = [&](const LineRec& f, const double (&v)[PPT]) {
        const double s0 = f.gs0, s1 = f.gs1, s2 = f.gs2;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            acc0[k] = fma(s0, v[k], acc0[k]);
            acc1[k] = fma(s1, v[k], acc1[k]);
            acc2[k] = fma(s2, v[k], acc2[k]);
        }
    };

    // ---- phase 2: per group: centre prefetch, far-wing runs, centre accumulate -------------------
    constexpr int NPRE = (PPT >= 4) ? 2 : 4;
    for (int r = 0; r < n_rounds; r++) {
        mbar_wait(mbar, r & 1);
        const int r0 = r * REC_CAP, r1 = min(n_tot, r0 + REC_CAP);
        int g;
        {
            int lo = 0, hi = n_grp;   // first group with cum[g+1] > r0
            while (hi - lo > 1) { int m = (lo + hi) >> 1; if (cum[m] <= r0) lo = m; else hi = m; }
            g = lo;
        }
        for (; g < n_grp && cum[g] < r1; g++) {
            const int* rr = g_rng + GR * g;
            if (rr[5] == rr[0]) continue;
            if (g != cur_g) { flush(cur_g); cur_g = g; }
            const int base = cum[g] - rr[0];   // slot = base + line
            // centre values of this group's lines: issue the (L2) loads before the wing math
            const int q0 = max(base + rr[6], r0), q1 = min(base + rr[7], r1);
            const int n_c = max(q1 - q0, 0);
            double v[NPRE][PPT];
#pragma unroll
            for (int q = 0; q < NPRE; q++)
                if (q < n_c) centre_load(recbuf[q0 - r0 + q], q0 + q - base, v[q]);
#pragma unroll
            for (int part = 0; part < 5; part++) {
                const int s0 = max(base + rr[part], r0), s1 = min(base + rr[part + 1], r1);
                if (s1 <= s0) continue;
                const LineRec* fr = recbuf + (s0 - r0);
                if (part == 1) wing_run(fr, s1 - s0, true);
                else if (part == 3) wing_run(fr, s1 - s0, false);
                else pred_run(fr, s1 - s0);
            }
#pragma unroll
            for (int q = 0; q < NPRE; q++)
                if (q < n_c) centre_add(recbuf[q0 - r0 + q], v[q]);
            for (int qb = NPRE; qb < n_c; qb += NPRE) {   // further batches: loads first, then adds
#pragma unroll
                for (int q = 0; q < NPRE; q++)
                    if (qb + q < n_c) centre_load(recbuf[q0 - r0 + qb + q], q0 + qb + q - base, v[q]);
#pragma unroll
                for (int q = 0; q < NPRE; q++)
                    if (qb + q < n_c) centre_add(recbuf[q0 - r0 + qb + q], v[q]);
            }
        }
        if (r + 1 < n_rounds) {
            __syncthreads();          // everyone is done reading recbuf
            issue_round(r + 1);
        }
    }
    flush(cur_g);
    __syncthreads();

    // ---- phase 4: write the tile, coalesced.  Rows fed by this chunk only: plain store (every
    // element written exactly once).  Rows shared with other chunks: FP64 atomic add onto the
    // row zeroed by k_zero_rows.
    double* __restrict__ out = a.out + (size_t)cell * a.n_sets * 3 * a.n_grid;
    for (int slot = 0; slot < a.max_rows; slot++) {
        const int row = a.chunk_rows[chunk * a.max_rows + slot];
        if (row < 0) continue;
        const bool shared = a.chunk_shared[chunk * a.max_rows + slot] != 0;
#pragma unroll
        for (int k = 0; k < PPT; k++) {
            const int p = tid + k * NT;
            const long s = tile0 + p;
            if (s < a.n_grid) {
                double* dst = out + (size_t)row * a.n_grid + s;
                const double val = acc_s[slot * TP + p];
                if (shared) atomicAdd(dst, val); else __stcs(dst, val);
            }
        }
    }
}

// zero the output rows that no chunk owns exclusively (shared rows and rows without any line)
__global__ void k_zero_rows(double* __restrict__ out, const int* __restrict__ rows, int n_rows_z,
                            int n_rows_cell, long n_grid) {
    const int cell = blockIdx.z, row = rows[blockIdx.y];
    double* dst = out + ((size_t)cell * n_rows_cell + row) * n_grid;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_grid;
         i += (long)gridDim.x * blockDim.x)
        dst[i] = 0.0;
    (void)n_rows_z;
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
    sr::DevBuf<int> ind, grp_begin, grp_up, grp_lo, flags, il_min, ir_max;
    sr::DevBuf<LineCell> rec;
    sr::DevBuf<LineRec> lrec;
    sr::DevBuf<double> pt, facs, core;
    sr::DevBuf<int> chunk_gbeg, grp_slot, chunk_rows, chunk_shared, zero_rows;
    int n_chunks = 1, max_rows = 1, max_grp = 1, n_zero_rows = 0;
    std::vector<int> order;    // sorted position -> input line
    std::vector<int> ind_in;   // input line -> centre index (-1 dropped)
    int max_cells_per_batch = 1;
};

namespace {

size_t tile_smem(int tp, int max_rows, int max_grp) {
    return (size_t)max_rows * tp * sizeof(double) + REC_CAP * sizeof(LineRec) + 16 +
           (size_t)(GR * max_grp + max_grp + 1) * sizeof(int) + 16;
}

template <int NT, int PPT, int MINB>
int launch_tile(const TileArgs& ta, int n_cells, int max_grp, cudaStream_t st) {
    const size_t smem = tile_smem(NT * PPT, ta.max_rows, max_grp);
    SR_CUDA(cudaFuncSetAttribute(k_voigt_tile<NT, PPT, MINB>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int TP = NT * PPT;
    dim3 grid((unsigned)((ta.n_grid + TP - 1) / TP), (unsigned)(n_cells * ta.n_chunks));
    SR_LAUNCH((k_voigt_tile<NT, PPT, MINB>), grid, NT, smem, st, ta);
    return SR_OK;
}

// tile configurations (threads, points per thread, min CTAs per SM); SR_K1_CFG=<index> forces one
struct TileCfg { int nt, ppt, minb; };
constexpr TileCfg kTileCfgs[] = {{256, 4, 2}, {256, 2, 2}, {256, 4, 1}, {256, 2, 1}, {256, 1, 1},
                                 {512, 2, 1}, {128, 4, 2}};
constexpr int kNumCfgs = (int)(sizeof(kTileCfgs) / sizeof(kTileCfgs[0]));

bool cfg_fits(int i, int max_rows, int max_grp, size_t smem_max, size_t smem_sm) {
    const size_t need = tile_smem(kTileCfgs[i].nt * kTileCfgs[i].ppt, max_rows, max_grp);
    if (need > smem_max) return false;
    return kTileCfgs[i].minb * (need + 1024) <= smem_sm;
}

int pick_cfg(int max_rows, int max_grp, size_t smem_max, size_t smem_sm) {
    if (const char* e = getenv("SR_K1_CFG")) {
        int i = atoi(e);
        if (i >= 0 && i < kNumCfgs && cfg_fits(i, max_rows, max_grp, smem_max, smem_sm)) return i;
    }
    for (int i = 0; i < 5; i++)
        if (cfg_fits(i, max_rows, max_grp, smem_max, smem_sm)) return i;
    return -1;
}

int launch_cfg(int cfg, const TileArgs& ta, int n_cells, int max_grp, cudaStream_t st) {
    switch (cfg) {
        case 0: return launch_tile<256, 4, 2>(ta, n_cells, max_grp, st);
        case 1: return launch_tile<256, 2, 2>(ta, n_cells, max_grp, st);
        case 2: return launch_tile<256, 4, 1>(ta, n_cells, max_grp, st);
        case 3: return launch_tile<256, 2, 1>(ta, n_cells, max_grp, st);
        case 4: return launch_tile<256, 1, 1>(ta, n_cells, max_grp, st);
        case 5: return launch_tile<512, 2, 1>(ta, n_cells, max_grp, st);
        case 6: return launch_tile<128, 4, 2>(ta, n_cells, max_grp, st);
    }
    return sr::fail(SR_ERR_ARG, "bad tile configuration");
}

// Partition the (sorted) groups into n_chunks contiguous chunks of similar line count and build
// the compact row tables of every chunk.  Returns max rows per chunk.
struct ChunkPlan {
    int n_chunks = 1, max_rows = 0, max_grp = 0;
    std::vector<int> gbeg, slot, rows, shared, zero_rows;
};

ChunkPlan plan_chunks(int n_chunks, int n_sets, const std::vector<int>& gb,
                      const std::vector<int>& gu, const std::vector<int>& gl) {
    ChunkPlan P;
    const int n_groups = (int)gu.size();
    n_chunks = std::max(1, std::min(n_chunks, std::max(n_groups, 1)));
    P.n_chunks = n_chunks;
    const int n_lines = n_groups ? gb[n_groups] : 0;
    P.gbeg.assign(n_chunks + 1, n_groups);
    P.gbeg[0] = 0;
    {
        int g = 0;
        for (int c = 0; c < n_chunks; c++) {
            P.gbeg[c] = g;
            const long target = (long)n_lines * (c + 1) / n_chunks;
            const int g_min = g + 1, g_max = n_groups - (n_chunks - 1 - c);
            while (g < g_max && (g < g_min || gb[g + 1] <= target)) g++;
        }
        P.gbeg[n_chunks] = n_groups;
    }
    const int n_rows = n_sets * 3;
    std::vector<std::vector<int>> rowlist(n_chunks);
    std::vector<int> users(n_rows, 0);
    P.slot.assign(3 * std::max(n_groups, 1), 0);
    for (int c = 0; c < n_chunks; c++) {
        std::vector<int> map(n_rows, -1);
        auto get = [&](int row) {
            if (map[row] < 0) { map[row] = (int)rowlist[c].size(); rowlist[c].push_back(row); users[row]++; }
            return map[row];
        };
        for (int g = P.gbeg[c]; g < P.gbeg[c + 1]; g++) {
            P.slot[3 * g + 0] = get(gu[g] * 3 + 0);
            P.slot[3 * g + 1] = get(gu[g] * 3 + 1);
            P.slot[3 * g + 2] = get(gl[g] * 3 + 2);
        }
        P.max_rows = std::max(P.max_rows, (int)rowlist[c].size());
        P.max_grp = std::max(P.max_grp, P.gbeg[c + 1] - P.gbeg[c]);
    }
    P.max_rows = std::max(P.max_rows, 1);
    P.rows.assign((size_t)n_chunks * P.max_rows, -1);
    P.shared.assign((size_t)n_chunks * P.max_rows, 0);
    for (int c = 0; c < n_chunks; c++)
        for (size_t i = 0; i < rowlist[c].size(); i++) {
            P.rows[(size_t)c * P.max_rows + i] = rowlist[c][i];
            P.shared[(size_t)c * P.max_rows + i] = users[rowlist[c][i]] > 1;
        }
    for (int row = 0; row < n_rows; row++)
        if (users[row] != 1) P.zero_rows.push_back(row);   // shared rows and rows without lines
    return P;
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
            // smallest chunk count whose tile (1024 points) fits twice per SM; else fewest rows
            int want = 0;
            if (const char* e = getenv("SR_K1_CHUNKS")) want = atoi(e);
            ChunkPlan best;
            bool have = false;
            for (int c = 1; c <= MAX_CHUNKS; c++) {
                if (want > 0 && c != want) continue;
                ChunkPlan P = plan_chunks(c, n_sets, gb, gu, gl);
                if (!have || P.max_rows < best.max_rows) { best = P; have = true; }
                if (want > 0 || 2 * (tile_smem(1024, P.max_rows, P.max_grp) + 1024) <= (size_t)228 * 1024) {
                    best = P;
                    break;
                }
                if (P.n_chunks < c) break;
            }
            ls->n_chunks = best.n_chunks;
            ls->max_rows = best.max_rows;
            ls->max_grp = best.max_grp;
            ls->n_zero_rows = (int)best.zero_rows.size();
            SR_CUDA(ls->chunk_gbeg.upload(best.gbeg.data(), best.gbeg.size(), st));
            SR_CUDA(ls->grp_slot.upload(best.slot.data(), best.slot.size(), st));
            SR_CUDA(ls->chunk_rows.upload(best.rows.data(), best.rows.size(), st));
            SR_CUDA(ls->chunk_shared.upload(best.shared.data(), best.shared.size(), st));
            if (ls->n_zero_rows) SR_CUDA(ls->zero_rows.upload(best.zero_rows.data(), best.zero_rows.size(), st));
            SR_CUDA(cudaStreamSynchronize(st));
        }
        // ind / gc in sorted order
        SR_LAUNCH(k_closest_grid, (n_act + 255) / 256, 256, 0, st, ls->grid.p, n_grid,
                  ls->freq.p, n_act, ls->ind.p, ls->gc.p);
        SR_CUDA(cudaStreamSynchronize(st));
        return SR_OK;
    };
    int code = body();
    if (code != SR_OK) { delete ls; return code; }
    // cells per batch: keep the per-(line,cell) tables under ~6 GiB
    size_t per_cell = (size_t)std::max(n_act, 1) *
                      (sizeof(LineCell) + sizeof(LineRec) + CORE_STRIDE * sizeof(double));
    ls->max_cells_per_batch =
        (int)std::max<size_t>(1, std::min<size_t>(4096, ((size_t)6 << 30) / per_cell));
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
    SR_CUDA(ls->lrec.ensure((size_t)n_cells * ls->n_act));
    SR_CUDA(ls->il_min.ensure(n_cells));
    SR_CUDA(ls->ir_max.ensure(n_cells));
    SR_CUDA(cudaMemsetAsync(ls->il_min.p, 0x7f, sizeof(int) * n_cells, st));
    SR_CUDA(cudaMemsetAsync(ls->ir_max.p, 0, sizeof(int) * n_cells, st));
    ParamsArgs pa;
    pa.L = {ls->freq.p, ls->a_coeff.p, ls->air.p, ls->tdep.p, ls->e_lower.p, ls->g_up.p,
            ls->g_lo.p, ls->evu.p, ls->evl.p, ls->gc.p, ls->ind.p};
    pa.lin = ls->lin.p;
    pa.pt = ls->pt.p;
    pa.rec = ls->rec.p;
    pa.lrec = ls->lrec.p;
    pa.il_min = ls->il_min.p;
    pa.ir_max = ls->ir_max.p;
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
    int dev = 0, smem_max = 0, smem_sm = 0;
    SR_CUDA(cudaGetDevice(&dev));
    SR_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    SR_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    const int cfg = pick_cfg(ls->max_rows, ls->max_grp, (size_t)smem_max, (size_t)smem_sm);
    if (cfg < 0)
        return sr::fail(SR_ERR_LIMIT, "%d output rows per chunk need more than %d bytes of shared "
                        "memory", ls->max_rows, smem_max);
    for (int c0 = 0; c0 < n_cells; c0 += ls->max_cells_per_batch) {
        const int nb = std::min(ls->max_cells_per_batch, n_cells - c0);
        if (c0 > 0) SR_CUDA(cudaStreamSynchronize(st));  // per-batch tables are reused
        int code = run_params(ls, pt_host + 2 * c0, nb, st);
        if (code) return code;
        SR_CUDA(ls->core.ensure((size_t)nb * ls->n_act * CORE_STRIDE));
        {
            dim3 cgrid((ls->n_act + 7) / 8, nb);
            SR_LAUNCH(k_core_eval, cgrid, 256, 0, st, ls->rec.p, ls->freq.p, ls->gc.p, ls->lin.p,
                      ls->n_act, ls->core.p);
        }
        TileArgs ta;
        ta.rec = ls->rec.p;
        ta.lrec = ls->lrec.p;
        ta.nu0 = ls->freq.p;
        ta.gc = ls->gc.p;
        ta.ind = ls->ind.p;
        ta.grp_begin = ls->grp_begin.p;
        ta.grp_up = ls->grp_up.p;
        ta.grp_lo = ls->grp_lo.p;
        ta.lin = ls->lin.p;
        ta.core = ls->core.p;
        ta.il_min = ls->il_min.p;
        ta.ir_max = ls->ir_max.p;
        ta.out = out_dev + (size_t)c0 * cell_elems;
        ta.n_grid = ls->n_grid;
        ta.n_lines = ls->n_act;
        ta.n_sets = ls->n_sets;
        ta.n_groups = ls->n_groups;
        ta.chunk_gbeg = ls->chunk_gbeg.p;
        ta.grp_slot = ls->grp_slot.p;
        ta.chunk_rows = ls->chunk_rows.p;
        ta.chunk_shared = ls->chunk_shared.p;
        ta.n_chunks = ls->n_chunks;
        ta.max_rows = ls->max_rows;
        if (ls->n_zero_rows) {
            dim3 zgrid(148, ls->n_zero_rows, nb);
            SR_LAUNCH(k_zero_rows, zgrid, 256, 0, st, ta.out, ls->zero_rows.p, ls->n_zero_rows,
                      ls->n_sets * 3, ls->n_grid);
        }
        code = launch_cfg(cfg, ta, nb, ls->max_grp, st);
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
