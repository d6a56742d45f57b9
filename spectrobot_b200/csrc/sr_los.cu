// sr_los.cu -- K3 / K3a: line-of-sight radiative transfer, batched over LOS x wavenumber.
//
// Reference pieces restated here:
//   LutSet.calculate              spect_main_module.py:997-1066  (nearest-node bilinear rule)
//   SpectralGcoeff.interpolate    spect_classes.py:1349-1375
//   make_abscoeff_LUTS_fast       spect_main_module.py:2134-2299 (level populations, abs/emi)
//   CalcPartitionSum              spect_classes.py:1692-1710
//   float32 LUT                   spect_main_module.py:1676 / spect_classes.py:732
// The layer recursion itself (sbm.LineOfSight.radtran_fast) is NOT in the reference tree; the
// specification implemented here is DESIGN.md section 6 ("parity unpinned" against the original).
//
// Kernels
//   k_step_weights  one thread per (gas, LOS, step): LUT cell indices + combined weights
//                   W[set][cell] = w_cell * pop_set * iso_ratio * column
//   k_los_fused     K3a+K3: thread per grid point, sequential over steps; LUT rows are read as
//                   coalesced float32 streams (L2-resident), tau/J never touch HBM
//   k_los_tau_src   K3a alone: materialises tau and S = J/tau
//   k_los_layers    K3 alone: HBM-streaming recursion over materialised tau/S
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#include "sr_common.h"
#include "sr_device.cuh"

namespace {

constexpr int MAX_GAS = 8;
constexpr int TIPS_N = SR_TIPS_N;
enum : int { LFLAG_EXTRAP_P = 1, LFLAG_NO_CELL = 2, LFLAG_NONFINITE = 4 };

struct GasDev {
    const float* g32;       // [n_cells][n_sets][3][n_grid]
    const double* Ps;       // [nP] unique sorted pressures
    const double* Ts;       // [nT] unique sorted temperatures
    const int* cellmap;     // [nP][nT] -> cell index or -1
    const double* qtab;     // [119] TIPS Q(T) row
    const double* elev;     // [n_sets] level energies (cm-1)
    const int* rowlist;     // [3][n_sets] sets whose (set, ctype) LUT row is non-zero somewhere, per
                            // ctype (all-zero spectra are None in the reference and skipped,
                            // smm:1674-1679, 2244-2249)
    int n_rows[3];          // list lengths: sp_emission, ind_emission, absorption
    int nP, nT, n_sets, lte_unidentified;
    long row_stride;        // floats between consecutive LUT rows (>= n_grid)
    double iso_ratio;
};

struct StepArgs {
    GasDev gas[MAX_GAS];
    int n_gas, n_los, n_steps_max, n_sets_max;
    const int* n_steps;       // [n_los]
    const double* temp;       // [n_los][n_steps_max]
    const double* pres;
    const double* column;     // [n_gas][n_los][n_steps_max]
    const double* tvib;       // [n_gas][n_sets_max][n_los][n_steps_max] or nullptr
    int* cells;               // out [n_gas][n_los][n_steps_max][4]
    double* W;                // out [n_gas][n_los][n_steps_max][n_sets_max][4]
    int* flags;
    double c2;                // h c / k  (spect_classes.py:47)
};

// nearest and second-nearest node, ties -> lower index (np.argmin / stable argsort()[1])
__device__ void nearest_two(const double* __restrict__ nodes, int n, double v, int& i1, int& i2) {
    int a = 0;
    for (int i = 1; i < n; i++)
        if (fabs(nodes[i] - v) < fabs(nodes[a] - v)) a = i;
    int b = -1;
    for (int i = 0; i < n; i++) {
        if (i == a) continue;
        if (b < 0 || fabs(nodes[i] - v) < fabs(nodes[b] - v)) b = i;
    }
    i1 = a;
    i2 = b;
}

// CalcPartitionSum: Lagrange through the <=2 nodes at or below T and the <=2 nodes above
__device__ double partition_sum(const double* __restrict__ q, double temp) {
    int nle = 0;
    while (nle < TIPS_N && 60.0 + 25.0 * nle <= temp) nle++;
    const int lo = nle - 2 < 0 ? 0 : nle - 2;
    const int hi = nle + 2 > TIPS_N ? TIPS_N : nle + 2;
    double acc = 0.0;
    for (int a = lo; a < hi; a++) {
        double w = q[a];
        const double ta = 60.0 + 25.0 * a;
        for (int b = lo; b < hi; b++)
            if (b != a) { const double tb = 60.0 + 25.0 * b; w *= (temp - tb) / (ta - tb); }
        acc += w;
    }
    return acc;
}

__device__ void step_weights_one(const StepArgs& a, int k, int l, int m) {
    const GasDev& G = a.gas[m];
    const size_t sk = (size_t)l * a.n_steps_max + k;
    const size_t o = ((size_t)m * a.n_los + l) * a.n_steps_max + k;
    int* cells = a.cells + o * 4;
    double* W = a.W + o * a.n_sets_max * 4;
    cells[0] = cells[1] = cells[2] = cells[3] = -1;
    for (int i = 0; i < a.n_sets_max * 4; i++) W[i] = 0.0;
    if (k >= a.n_steps[l]) return;
    const double temp = a.temp[sk], pres = a.pres[sk];
    double wc[4] = {0.0, 0.0, 0.0, 0.0};
    int flags = 0;
    if (G.nT < 2) flags |= LFLAG_NO_CELL;
    else if (pres <= G.Ps[0]) {                                   // smm:1007-1025
        int ta, tb;
        nearest_two(G.Ts, G.nT, temp, ta, tb);
        cells[0] = G.cellmap[0 * G.nT + ta];
        cells[1] = G.cellmap[0 * G.nT + tb];
        if (cells[0] < 0 || cells[1] < 0) flags |= LFLAG_NO_CELL;
        wc[0] = (G.Ts[tb] - temp) / (G.Ts[tb] - G.Ts[ta]);        // sbm.weight, DESIGN 6.2
        wc[1] = (temp - G.Ts[ta]) / (G.Ts[tb] - G.Ts[ta]);
    } else if (pres <= G.Ps[G.nP - 1]) {                          // smm:1026-1056
        if (G.nP < 2) flags |= LFLAG_NO_CELL;
        else {
            int p1, p2, t1, t2;
            nearest_two(G.Ps, G.nP, pres, p1, p2);
            nearest_two(G.Ts, G.nT, temp, t1, t2);
            cells[0] = G.cellmap[p1 * G.nT + t1];
            cells[1] = G.cellmap[p1 * G.nT + t2];
            cells[2] = G.cellmap[p2 * G.nT + t1];
            cells[3] = G.cellmap[p2 * G.nT + t2];
            if (cells[0] < 0 || cells[1] < 0 || cells[2] < 0 || cells[3] < 0)
                flags |= LFLAG_NO_CELL;
            const double wp1 = (G.Ps[p2] - pres) / (G.Ps[p2] - G.Ps[p1]);
            const double wp2 = (pres - G.Ps[p1]) / (G.Ps[p2] - G.Ps[p1]);
            const double wt1 = (G.Ts[t2] - temp) / (G.Ts[t2] - G.Ts[t1]);
            const double wt2 = (temp - G.Ts[t1]) / (G.Ts[t2] - G.Ts[t1]);
            wc[0] = wt1 * wp1; wc[1] = wt2 * wp1; wc[2] = wt1 * wp2; wc[3] = wt2 * wp2;
        }
    } else {
        flags |= LFLAG_EXTRAP_P;                                  // smm:1058
    }
    if (flags) {
        atomicOr(a.flags, flags);
        cells[0] = cells[1] = cells[2] = cells[3] = -1;
        return;
    }
    const double q_part = partition_sum(G.qtab, temp);            // smm:2212
    const double col = G.iso_ratio * a.column[o];
    for (int s = 0; s < G.n_sets; s++) {
        double pop;
        if (G.lte_unidentified) pop = 1 / q_part;                 // smm:2218
        else {
            const double vibt = a.tvib
                ? a.tvib[(((size_t)m * a.n_sets_max + s) * a.n_los + l) * a.n_steps_max + k]
                : temp;                                           // smm:2231-2234
            pop = exp(-a.c2 * G.elev[s] / vibt) / q_part;         // smm:2241
        }
        for (int c = 0; c < 4; c++) {
            const double w = wc[c] * pop * col;
            if (!isfinite(w)) flags |= LFLAG_NONFINITE;
            W[s * 4 + c] = w;
        }
    }
    if (flags) atomicOr(a.flags, flags);
}

__global__ void k_step_weights(const __grid_constant__ StepArgs a) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_steps_max) return;
    for (int l = blockIdx.y; l < a.n_los; l += gridDim.y) step_weights_one(a, k, l, blockIdx.z);
}

// which LUT rows (set, ctype) hold any non-zero value (over all cells and grid points)
__global__ void k_row_nonzero(const float* __restrict__ g32, int n_cells, int n_sets, long n_grid,
                              long row_stride, int* __restrict__ rowmask) {
    const long n_rows = (long)n_cells * n_sets * 3;
    for (long r = blockIdx.y; r < n_rows; r += gridDim.y) {   // (cell, set, ctype) row
        const int s = (int)((r / 3) % n_sets), ct = (int)(r % 3);
        const float* __restrict__ p = g32 + (size_t)r * row_stride;
        bool nz = false;
        for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_grid;
             i += (long)gridDim.x * blockDim.x)
            nz |= (p[i] != 0.0f);
        if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(rowmask + s, 1 << ct);
    }
}

struct LosArgs {
    GasDev gas[MAX_GAS];
    int n_gas, n_los, n_steps_max, n_sets_max;
    const int* n_steps;
    const int* cells;
    const double* W;
    long n_grid, pt0, n_pts;
    const double* i0;     // [n_los][n_pts] or nullptr
    double* rad;          // [n_los][n_pts]
    double* tau_out;      // [n_los][n_steps_max][n_pts] (k_los_tau_src)
    double* src_out;
    int solo_absorption;
    int emit_j;           // k_los_tau_src: write J (emission coefficient x column) instead of S = J/tau
};

// tau and J of one step for PPT points of one thread
template <int PPT>
__device__ __forceinline__ void step_tau_j(const LosArgs& a, int l, int k, long p_first,
                                           const bool (&ok)[PPT], double (&tau)[PPT],
                                           double (&J)[PPT]) {
#pragma unroll
    for (int i = 0; i < PPT; i++) tau[i] = J[i] = 0.0;
    for (int m = 0; m < a.n_gas; m++) {
        const GasDev& G = a.gas[m];
        const size_t o = ((size_t)m * a.n_los + l) * a.n_steps_max + k;
        const int4 cells = __ldg(reinterpret_cast<const int4*>(a.cells + o * 4));
        const double* __restrict__ W = a.W + o * a.n_sets_max * 4;
        const int cl[4] = {cells.x, cells.y, cells.z, cells.w};
        const size_t row = (size_t)G.row_stride;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            if (cl[c] < 0) continue;
            const float* __restrict__ base =
                G.g32 + (size_t)cl[c] * G.n_sets * 3 * row + (size_t)(a.pt0 + p_first);
            // one branch-free loop per ctype over the non-zero rows (independent coalesced loads)
#pragma unroll
            for (int ct = 0; ct < 3; ct++) {
                const int* __restrict__ list = G.rowlist + ct * G.n_sets;
                const int n = G.n_rows[ct];
#pragma unroll 4
                for (int j = 0; j < n; j++) {
                    const int s = __ldg(list + j);
                    double w = __ldg(W + s * 4 + c);
                    if (ct == 1) w = -w;                 // abs_coeff -= G_ind*pop, smm:2246-2247
                    const float* __restrict__ r0 = base + (size_t)(s * 3 + ct) * row;
#pragma unroll
                    for (int i = 0; i < PPT; i++) {
                        if (!ok[i]) continue;
                        const double v = (double)__ldg(r0 + i * 256);
                        if (ct == 0) J[i] = fma(w, v, J[i]);        // smm:2248-2249
                        else tau[i] = fma(w, v, tau[i]);            // smm:2244-2247
                    }
                }
            }
        }
    }
}

// I <- I exp(-tau) + J phi(tau), phi = (1 - exp(-tau))/tau (1 at tau = 0)
__device__ __forceinline__ double layer_update(double I, double tau, double J, int solo) {
    double t, em;
    srdev::exp_pair(-tau, t, em);
    if (solo) return I * t;
    const double phi = (tau == 0.0) ? 1.0 : -em / tau;
    return fma(I, t, J * phi);
}

template <int PPT, bool MATERIALISE>
__global__ void __launch_bounds__(1024) k_los_fused(LosArgs a) {
    // blockDim = (256, G): G consecutive LOS (the 3 LOS of a pixel are neighbours in the batch and
    // cross nearly the same (P,T) cells) share one wavenumber tile, so that the LUT rows one of
    // them pulls from L2 are L1 hits for the others.
    // grid = (LOS groups, wavenumber tiles): all LOS of one wavenumber tile are scheduled next to
    // each other, so the LUT rows of that tile are read from HBM once and then served by L2
    const int l = blockIdx.x * blockDim.y + threadIdx.y;
    if (l >= a.n_los) return;
    const long p_first = (long)blockIdx.y * (256 * PPT) + threadIdx.x;
    bool ok[PPT];
    double I[PPT], tau[PPT], J[PPT];
#pragma unroll
    for (int i = 0; i < PPT; i++) {
        ok[i] = p_first + i * 256 < a.n_pts;
        I[i] = (ok[i] && a.i0) ? a.i0[(size_t)l * a.n_pts + p_first + i * 256] : 0.0;
    }
    const int ns = a.n_steps[l];
    for (int k = 0; k < ns; k++) {
        step_tau_j<PPT>(a, l, k, p_first, ok, tau, J);
        if (MATERIALISE) {
            const size_t o = ((size_t)l * a.n_steps_max + k) * a.n_pts + p_first;
#pragma unroll
            for (int i = 0; i < PPT; i++)
                if (ok[i]) {
                    __stcs(a.tau_out + o + i * 256, tau[i]);
                    __stcs(a.src_out + o + i * 256,
                           a.emit_j ? J[i] : (tau[i] == 0.0 ? 0.0 : J[i] / tau[i]));
                }
        } else {
#pragma unroll
            for (int i = 0; i < PPT; i++) I[i] = layer_update(I[i], tau[i], J[i], a.solo_absorption);
        }
    }
    if (!MATERIALISE) {
#pragma unroll
        for (int i = 0; i < PPT; i++)
            if (ok[i]) __stcs(a.rad + (size_t)l * a.n_pts + p_first + i * 256, I[i]);
    }
}

// One LUT row of a quad program (K3a, grouped product): the (LOS, step) pairs of a batch are
// grouped by the LUT cells their interpolation uses ("quad": up to 4 cells per gas); every needed
// LUT row is read once per CTA and used for all pairs of the chunk:
//   tau[pair][p] = sum_rows w_abs/ind[pair][row] * G[row][p],  J[pair][p] = sum_rows w_sp * G
// Output goes to the [los][step][point] layer arrays that k_los_layers then streams.
constexpr int GEMM_MAXJ = 8 * 4 * 3 * 16;   // rows of one quad program (capacity check on the host)

struct ProgEntry {
    long long roff;               // address of the row's first element (float*) on the device
    int gas;                      // which LUT (-1: zero-weight padding row)
    int widx;                     // set*4 + cell slot: index into the pair's weight block
    int neg;                      // 1: subtract (ind_emission row), 0: add
    int pad;
};

// K3: I <- I exp(-tau) + S (1 - exp(-tau)) over materialised layers; pure HBM streaming.
// A CTA is (LOS, tile of NT*PPT points): each thread owns PPT points (stride NT, so every
// warp load is one coalesced 256-byte row segment for any n_pts parity) and keeps UNROLL steps x
// PPT points x 2 arrays of loads in flight.
struct RecArgs {
    const double* tau;        // [n_los][n_steps_max][lay_stride]
    const double* src;
    const int* n_steps;       // [n_los]
    const double* i0;         // [n_los][io_stride] window at io_off, or nullptr
    double* rad;              // [n_los][io_stride] window at io_off
    long n_pts, io_stride, io_off, lay_stride;
    long n_work;              // n_los * n_tiles (0: nothing to do)
    int n_steps_max, n_tiles, solo, src_is_j;
    int keep;                 // layers were just written and are expected in L2: plain loads
    int f32;                  // the layer arrays hold float32 values (same element indexing)
};

template <int PPT, int UNROLL, int NT>
__device__ __forceinline__ void layers_item(const RecArgs& r, int l, int tile) {
    const long p0 = (long)tile * (NT * PPT) + threadIdx.x;
    if (p0 >= r.n_pts) return;
    bool ok[PPT];
    double I[PPT];
#pragma unroll
    for (int i = 0; i < PPT; i++) {
        ok[i] = p0 + i * NT < r.n_pts;
        I[i] = (r.i0 && ok[i]) ? r.i0[(size_t)l * r.io_stride + r.io_off + p0 + i * NT] : 0.0;
    }
    const int ns = r.n_steps[l];
    const long ls = r.lay_stride;
    const size_t col0 = (size_t)l * r.n_steps_max * ls + p0;
    const double* __restrict__ tp = r.tau + col0;
    const double* __restrict__ sp = r.src + col0;
    const float* __restrict__ tpf = reinterpret_cast<const float*>(r.tau) + col0;
    const float* __restrict__ spf = reinterpret_cast<const float*>(r.src) + col0;
    const int solo = r.solo, src_is_j = r.src_is_j;
    const bool keep = r.keep != 0, f32 = r.f32 != 0;
    // layer elements at offset o of this thread's column (float32 scratch: same element indexing)
    auto ld_tau = [&](size_t o) {
        if (f32) return (double)(keep ? __ldcg(tpf + o) : __ldcs(tpf + o));
        return keep ? __ldcg(tp + o) : __ldcs(tp + o);
    };
    auto ld_src = [&](size_t o) {
        if (f32) return (double)(keep ? __ldcg(spf + o) : __ldcs(spf + o));
        return keep ? __ldcg(sp + o) : __ldcs(sp + o);
    };
    auto update = [&](int i, double t, double s) {
        if (src_is_j) {   // s = J: I <- I e^-tau + J phi(tau), phi = (1 - e^-tau)/tau
            I[i] = srdev::layer_update_j(I[i], t, s, solo != 0);
            return;
        }
        double ex, em;
        srdev::exp_pair(-t, ex, em);
        {
            I[i] = solo ? I[i] * ex : fma(I[i], ex, -s * em);
        }
    };
    int k = 0;
    for (; k + UNROLL <= ns; k += UNROLL) {
        double t[UNROLL][PPT], s[UNROLL][PPT];
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int i = 0; i < PPT; i++) {
                const size_t o = (size_t)(k + u) * ls + i * NT;
                t[u][i] = ok[i] ? ld_tau(o) : 0.0;
                s[u][i] = ok[i] ? ld_src(o) : 0.0;
            }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
#pragma unroll
            for (int i = 0; i < PPT; i++) update(i, t[u][i], s[u][i]);
    }
    for (; k < ns; k++)
#pragma unroll
        for (int i = 0; i < PPT; i++)
            if (ok[i]) update(i, ld_tau((size_t)k * ls + i * NT), ld_src((size_t)k * ls + i * NT));
#pragma unroll
    for (int i = 0; i < PPT; i++)
        if (ok[i]) __stcs(r.rad + (size_t)l * r.io_stride + r.io_off + p0 + i * NT, I[i]);
}

template <int PPT, int UNROLL>
__global__ void __launch_bounds__(256, 6) k_los_layers(const __grid_constant__ RecArgs r) {
    layers_item<PPT, UNROLL, 256>(r, (int)blockIdx.y, (int)blockIdx.x);
}

// K3 over a float32 layer scratch (J form only): a thread owns two neighbouring points and reads
// them as one float2 per array and step, so that it keeps as many bytes in flight as the FP64
// kernel does (UNROLL steps x 2 arrays x 8 B).  Rows are 16-byte aligned (lay_stride % 4 == 0).
template <int UNROLL>
__global__ void __launch_bounds__(256, 6) k_los_layers_f32(const __grid_constant__ RecArgs r) {
    const int l = blockIdx.y;
    const long p0 = ((long)blockIdx.x * 256 + threadIdx.x) * 2;
    if (p0 >= r.n_pts) return;
    const bool ok1 = p0 + 1 < r.n_pts;
    double I0 = r.i0 ? r.i0[(size_t)l * r.io_stride + r.io_off + p0] : 0.0;
    double I1 = (r.i0 && ok1) ? r.i0[(size_t)l * r.io_stride + r.io_off + p0 + 1] : 0.0;
    const int ns = r.n_steps[l];
    const long ls = r.lay_stride;
    const size_t col0 = (size_t)l * r.n_steps_max * ls + p0;
    const float* __restrict__ tp = reinterpret_cast<const float*>(r.tau) + col0;
    const float* __restrict__ sp = reinterpret_cast<const float*>(r.src) + col0;
    const bool solo = r.solo != 0;
    auto ld2 = [&](const float* p) { return __ldcs(reinterpret_cast<const float2*>(p)); };
    const unsigned live = __activemask();   // (threads beyond the window have left; ns is uniform)
    // one (step, two points) update; the form of the exponential is chosen per warp
    auto upd = [&](float2 t, float2 s) {
        const float m = fmaxf(fabsf(t.x), fabsf(t.y));
        if (__all_sync(live, m < 1.0e-2f)) {
            I0 = srdev::layer_update_j_small(I0, (double)t.x, (double)s.x, solo);
            I1 = srdev::layer_update_j_small(I1, (double)t.y, (double)s.y, solo);
        } else if (__all_sync(live, m < 0.34f)) {
            I0 = srdev::layer_update_j_medium(I0, (double)t.x, (double)s.x, solo);
            I1 = srdev::layer_update_j_medium(I1, (double)t.y, (double)s.y, solo);
        } else {
            I0 = srdev::layer_update_j_f32in(I0, (double)t.x, (double)s.x, solo);
            I1 = srdev::layer_update_j_f32in(I1, (double)t.y, (double)s.y, solo);
        }
    };
    int k = 0;
    for (; k + UNROLL <= ns; k += UNROLL) {
        float2 t[UNROLL], s[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            t[u] = ld2(tp + (size_t)(k + u) * ls);
            s[u] = ld2(sp + (size_t)(k + u) * ls);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) upd(t[u], s[u]);
    }
    for (; k < ns; k++) upd(ld2(tp + (size_t)k * ls), ld2(sp + (size_t)k * ls));
    double* o = r.rad + (size_t)l * r.io_stride + r.io_off + p0;
    __stcs(o, I0);
    if (ok1) __stcs(o + 1, I1);
}

// ---------------------------------------------------------------------------------------------
// K3 with analytic Jacobians (callers spect_main_module.py:2758, 2837,
// 2867-2881: `calc_derivatives=True`, `par.hires_deriv`).  A retrieval parameter p (a VMR node of
// one gas, RetParam.maskgrid :600-656) changes step k only through the gas column:
//   d tau_k / dp = tau_g,k * f_kp,   d J_k / dp = J_g,k * f_kp,   f_kp = (d u_k/dp) / u_k
// (tau_g, J_g: the layers of the retrieved gas alone; T and P of the step are held fixed).
// Differentiating the layer update I_k = I_{k-1} e^-tau + J phi(tau):
//   D_p,k = D_p,k-1 e^-tau_k + f_kp * B_k
//   B_k   = -I_{k-1} e^-tau tau_g + J_g phi + J phi'(tau) tau_g,   phi' = (e^-tau - phi)/tau
// which collapses to B_k = e^-tau (J - I_{k-1} tau) when the retrieved gas is the only absorber.
// One thread per grid point keeps I and NP derivative accumulators in registers and streams the
// layers once; more than NP parameters -> blockIdx.z chunks, each re-streaming the layers.
// ---------------------------------------------------------------------------------------------
struct JacArgs {
    RecArgs r;                // forward part; src is J (emission coefficient x column)
    const double* tau_g;      // layers of the retrieved gas alone (MULTI only)
    const double* src_g;
    const double* dfrac;      // [n_los][n_steps_max][n_par]
    double* jac;              // [n_los][n_par][io_stride], window at io_off
    int n_par;
};

// phi'(t) for phi(t) = (1 - e^-t)/t, given ex = e^-t and phi
__device__ __forceinline__ double dphi(double t, double ex, double phi) {
    if (fabs(t) < 0.05)   // Taylor series: the closed form cancels like eps/t near 0 (tau < 0 under
                          // population inversion takes the closed form like tau > 0)
        return fma(t, fma(t, fma(t, fma(t, fma(t, 1.0 / 840.0, -1.0 / 144.0), 1.0 / 30.0), -0.125),
                          1.0 / 3.0), -0.5);
    return (ex - phi) / t;
}

template <int NP, bool MULTI, int U, int MINB>
__global__ void __launch_bounds__(256, MINB) k_los_layers_jac(const __grid_constant__ JacArgs a) {
    const RecArgs& r = a.r;
    const int l = blockIdx.y, pz = blockIdx.z * NP;
    const long p = (long)blockIdx.x * 256 + threadIdx.x;
    const bool live = p < r.n_pts;
    double I = (r.i0 && live) ? r.i0[(size_t)l * r.io_stride + r.io_off + p] : 0.0;
    double D[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) D[q] = 0.0;
    const int ns = min(max(r.n_steps[l], 0), r.n_steps_max), solo = r.solo;   // (table width)
    const long ls = r.lay_stride;
    const size_t lay0 = (size_t)l * r.n_steps_max * ls + p;
    const double* __restrict__ tp = r.tau + lay0;
    const double* __restrict__ sp = r.src + lay0;
    const double* __restrict__ tgp = MULTI ? a.tau_g + lay0 : nullptr;
    const double* __restrict__ sgp = MULTI ? a.src_g + lay0 : nullptr;
    const bool f32 = r.f32 != 0;   // float32 layer scratch: same element indexing, half the bytes
    auto ldv = [&](const double* base, size_t o) {
        if (f32) return (double)__ldcs(reinterpret_cast<const float*>(base - lay0) + lay0 + o);
        return __ldcs(base + o);
    };
    const int npar = min(NP, a.n_par - pz);
    // this LOS' rows of the derivative table, zero-padded to NP columns, in shared memory: the
    // inner loop reads them as broadcast LDS.128 (no per-parameter predicate, no global latency)
    extern __shared__ __align__(16) double fs[];
    {
        const double* __restrict__ fr = a.dfrac + (size_t)l * r.n_steps_max * a.n_par + pz;
        for (int i = threadIdx.x; i < ns * NP; i += 256) {
            const int k = i / NP, q = i - k * NP;
            fs[i] = q < npar ? fr[(size_t)k * a.n_par + q] : 0.0;
        }
    }
    __syncthreads();
    if (!live) return;
    for (int k0 = 0; k0 < ns; k0 += U) {
        double t[U], s[U], tg[U], sg[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const bool ok = k0 + u < ns;
            const size_t o = (size_t)(k0 + u) * ls;
            t[u] = ok ? ldv(tp, o) : 0.0;
            s[u] = ok ? ldv(sp, o) : 0.0;
            if (MULTI) {
                tg[u] = ok ? ldv(tgp, o) : 0.0;
                sg[u] = ok ? ldv(sgp, o) : 0.0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (k0 + u >= ns) break;
            double ex, phi;
            srdev::exp_phi(t[u], ex, phi);
            double B;
            if (MULTI) {
                const double a0 = -I * ex * tg[u];
                B = solo ? a0 : fma(s[u] * dphi(t[u], ex, phi), tg[u], fma(sg[u], phi, a0));
            } else {
                B = solo ? -I * ex * t[u] : ex * fma(-I, t[u], s[u]);
            }
            I = solo ? I * ex : fma(I, ex, s[u] * phi);
            const double2* __restrict__ f2 = reinterpret_cast<const double2*>(fs + (k0 + u) * NP);
#pragma unroll
            for (int q = 0; q < NP; q += 2) {
                const double2 fq = f2[q >> 1];
                D[q] = fma(D[q], ex, fq.x * B);
                D[q + 1] = fma(D[q + 1], ex, fq.y * B);
            }
        }
    }
    if (blockIdx.z == 0) __stcs(r.rad + (size_t)l * r.io_stride + r.io_off + p, I);
#pragma unroll
    for (int q = 0; q < NP; q++)
        if (q < npar)
            __stcs(a.jac + ((size_t)l * a.n_par + pz + q) * r.io_stride + r.io_off + p, D[q]);
}

// ---------------------------------------------------------------------------------------------
// K3a on the FP64 tensor path (v3).  Same grouping as v2, but the per-CTA product
//   C[16 pairs][points] += A[16 pairs][rows] * B[rows][points]
// is issued as DMMA.8x8x4 (mma.sync.m8n8k4.f64): one warp instruction does the work of 8 DFMA
// warp instructions on the same FP64 pipe (tools/ubench/ub_dmma.cu: 37 TFLOP/s with ILP 1), so the
// pipe is fed with ~1.4 issue slots per 16 pipe cycles instead of ~1.5 per 2.
//   A fragment  lane -> W[pair = lane>>2][row j0 + (lane&3)]: packed once per call in fragment order
//               by k_pack_wfrag, copied to shared memory per CTA (one conflict-free LDS.64 per use)
//   B fragment  lane -> G[row j0 + (lane&3)][point 8i + (lane>>2)]: one float per lane straight from
//               the LUT (4 rows x 32 B per warp load), converted to double in registers
//   C fragment  lane -> pair lane>>2, points 8i + 2(lane&3) + {0,1}
// One warp owns 16 pairs x 8*NB points; the second 8-pair M block is skipped when the chunk has
// at most 8 pairs.  Rows are processed in two passes (tau rows, then emission rows) over the same
// accumulators; program rows are padded to multiples of 4 with zero weights.
// ---------------------------------------------------------------------------------------------
constexpr int MMA_PB = 16;
constexpr int MMA_NT = 128;

struct MmaArgs {
    const long long* rowptr;      // [n_groups][max_jp] device address of each program row
    const int* grp_ntau;          // [n_groups] tau rows, padded to a multiple of 4
    const int* grp_ntot;          // [n_groups] all rows (padded tau rows + padded emission rows)
    const int* chunk_grp;         // [n_chunks]
    const int* chunk_pair;        // [n_chunks][16] global pair index los*n_steps_max+step, -1 = padding
    const double* wfrag;          // [n_chunks][max_jp/4][2][32] fragment-ordered weights
    int max_jp, chunk0;
    long pair_base;               // first pair of the LOS block (rows of tau_out/src_out are local)
    long pt0, n_pts;
    long ld_min;                  // smallest LUT row stride: loads are clamped below it
    long ld_out;                  // stride of the output rows (>= n_pts)
    double* tau_out;              // [pairs of the block][ld_out]
    double* src_out;
    int mode;                     // 0: src = S = J/tau, 1: src = J
    int keep;                     // 1: the layer rows are read back at once (L2-sized scratch):
                                  // default-policy stores instead of streaming ones
    int f32;                      // 1: the layer rows are stored as float32 (same element indexing)
    int tpc;                      // point tiles per CTA
};

struct PackArgs {
    const ProgEntry* prog;        // [n_groups][max_jp] (gas < 0: padding row)
    const int* grp_ntot;
    const int* chunk_grp;
    const int* chunk_pair;
    const double* W;              // [n_gas][n_pairs_tot][n_sets_max*4]
    double* wfrag;
    long n_pairs_tot;
    int n_sets_max, max_jp, n_chunks;
    unsigned gas_mask;            // LUTs whose weights are packed (the others get 0)
    const int* grp_ntau;          // rows [grp_ntau, grp_ntot) of a program are emission rows
    unsigned long long emis[MAX_GAS];   // per LUT: sets whose spontaneous emission is kept
};

__global__ void k_pack_wfrag(PackArgs a) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long per_chunk = (long)a.max_jp * MMA_PB;
    if (e >= per_chunk * a.n_chunks) return;
    const int chunk = (int)(e / per_chunk);
    const int r = (int)(e % per_chunk), j = r / MMA_PB, slot = r % MMA_PB;
    const int grp = a.chunk_grp[chunk];
    if (j >= a.grp_ntot[grp]) return;
    const ProgEntry pe = a.prog[(size_t)grp * a.max_jp + j];
    const int pr = a.chunk_pair[chunk * MMA_PB + slot];
    double w = 0.0;
    if (pr >= 0 && pe.gas >= 0 && ((a.gas_mask >> pe.gas) & 1u) &&
        (j < a.grp_ntau[grp] || ((a.emis[pe.gas] >> (pe.widx >> 2)) & 1ull)))
        w = a.W[((size_t)pe.gas * a.n_pairs_tot + pr) * ((size_t)a.n_sets_max * 4) + pe.widx];
    if (pe.neg) w = -w;
    const int kb = j >> 2, kq = j & 3, mb = slot >> 3, row = slot & 7;
    a.wfrag[(((size_t)chunk * (a.max_jp >> 2) + kb) * 2 + mb) * 32 + row * 4 + kq] = w;
}

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// VEC: every LUT row and the point window are 16-byte aligned (padded LUT rows, pt0 % 4 == 0) and
// the output rows have a stride that is a multiple of 4 -> one LDG.128 per lane brings 4 points of
// one row (a warp load = 4 rows x 128 contiguous bytes), and each lane ends up with 8 consecutive
// points of one pair, written as 16-byte vectors.  Point <-> fragment maps:
//   VEC    B(v,i): point 32v + pn(lane>>2) + i, pn(n) = 4(n>>1) + 16(n&1)
//          C(v,i)[e]: point 32v + 16e + 4(lane&3) + i   (n = 2(lane&3) + e)
//          -> a store instruction (v,e) writes 32 B per lane, 128 contiguous bytes per pair row
//   scalar B(i)  : point 8i + (lane>>2)               C(i)[e]  : point 8i + 2(lane&3) + e
// one 32-byte streaming store (a full sector per lane)
__device__ __forceinline__ void st_cs_v4(double* p, const double (&x)[4]) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]),
                 "d"(x[3]) : "memory");
}
__device__ __forceinline__ void st_cg_v4(double* p, const double (&x)[4]) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]),
                 "d"(x[3]) : "memory");
}

// LUT row loads: LD = 0 read-only path (ld.global.nc), 1 streaming (ld.global.cs), 2 no L1
// allocation (tuning aid, SR_MMA_LD)
template <int LD>
__device__ __forceinline__ float4 ld_row4(const float* p) {
    float4 q;
    if (LD == 1) q = __ldcs(reinterpret_cast<const float4*>(p));
    else if (LD == 2)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(p));
    else q = __ldg(reinterpret_cast<const float4*>(p));
    return q;
}

template <int NB, bool VEC, int MINB, int LD>
__global__ void __launch_bounds__(MMA_NT, MINB) k_los_mma(MmaArgs a) {
    extern __shared__ __align__(16) unsigned char msm[];
    double* wA = reinterpret_cast<double*>(msm);                                   // [kb][2][32]
    long long* rps = reinterpret_cast<long long*>(wA + (size_t)a.max_jp * MMA_PB);   // [max_jp]
    __shared__ int pair_s[MMA_PB];
    const int chunk = a.chunk0 + blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int grp = a.chunk_grp[chunk];
    const int n_tau = a.grp_ntau[grp], n_tot = a.grp_ntot[grp];
    if (tid < MMA_PB) pair_s[tid] = a.chunk_pair[chunk * MMA_PB + tid];
    {
        const double2* __restrict__ wsrc =
            reinterpret_cast<const double2*>(a.wfrag + (size_t)chunk * a.max_jp * MMA_PB);
        double2* wdst = reinterpret_cast<double2*>(wA);
        for (int e = tid; e < n_tot * (MMA_PB / 2); e += MMA_NT) wdst[e] = __ldg(wsrc + e);
        const long long* __restrict__ rsrc = a.rowptr + (size_t)grp * a.max_jp;
        for (int e = tid; e < n_tot; e += MMA_NT) rps[e] = __ldg(rsrc + e);
    }
    __syncthreads();
    const bool two = pair_s[8] >= 0;
    const int kq = lane & 3, nq = lane >> 2;
    struct Frag { float v[NB]; };
    // a CTA walks a.tpc consecutive point tiles with the weights it staged once
    for (int it = 0; it < a.tpc; it++) {
    const long p_warp = ((long)blockIdx.y * a.tpc + it) * ((MMA_NT / 32) * 8 * NB) + wid * (8 * NB);
    if (p_warp >= a.n_pts) return;
    // per-lane point offsets inside a LUT row; lanes beyond the point window re-read the last
    // valid vector of the row (their columns are never stored), so every load is unconditional
    long loff[VEC ? NB / 4 : NB];
    if (VEC) {
#pragma unroll
        for (int v = 0; v < NB / 4; v++)
            loff[v] = min(a.pt0 + p_warp + 32 * v + 4 * (nq >> 1) + 16 * (nq & 1), a.ld_min - 4);
    } else {
#pragma unroll
        for (int i = 0; i < NB; i++) loff[i] = min(a.pt0 + p_warp + 8 * i + nq, a.ld_min - 1);
    }
    auto load = [&](int j, Frag& f) {
        const float* __restrict__ r = reinterpret_cast<const float*>(rps[j + kq]);
        if (VEC) {
#pragma unroll
            for (int v = 0; v < NB / 4; v++) {
                const float4 q = ld_row4<LD>(r + loff[v]);
                f.v[4 * v + 0] = q.x; f.v[4 * v + 1] = q.y; f.v[4 * v + 2] = q.z; f.v[4 * v + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NB; i++) f.v[i] = __ldg(r + loff[i]);
        }
    };
    double C0[NB][2], C1[NB][2];
    // rows [j0, j1) into the accumulators.  Three B-fragment buffers rotate without register
    // copies: a buffer is refilled (rows j+12.., clamped to the last k-step of the pass) right
    // after its values have been converted, so every load has two full k-steps (>= 32 DMMA) to
    // arrive.  The main loop has no conditional code around the buffers (a merge point would make
    // the compiler copy the freshly loaded registers, i.e. wait for the load at once).
    auto mma_step = [&](int j, const double (&b)[NB]) {
        const double a0 = wA[(j >> 2) * 64 + lane];
        if (two) {
            const double a1 = wA[(j >> 2) * 64 + 32 + lane];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                dmma884(C0[i], a0, b[i]);
                dmma884(C1[i], a1, b[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < NB; i++) dmma884(C0[i], a0, b[i]);
        }
    };
#define SR_KSTEP(J, F, RELOAD)                                        \
    {                                                                 \
        double b_[NB];                                                \
        _Pragma("unroll") for (int i = 0; i < NB; i++) b_[i] = (double)F.v[i]; \
        if (RELOAD) load(min((J) + 12, j1 - 4), F);                   \
        mma_step((J), b_);                                            \
    }
    auto pass = [&](int j0, int j1) {
#pragma unroll
        for (int i = 0; i < NB; i++) C0[i][0] = C0[i][1] = C1[i][0] = C1[i][1] = 0.0;
        if (j0 >= j1) return;
        Frag f0, f1, f2;
        load(j0, f0);
        load(min(j0 + 4, j1 - 4), f1);
        load(min(j0 + 8, j1 - 4), f2);
        int j = j0;
#pragma unroll 1
        for (; j + 12 <= j1; j += 12) {
            SR_KSTEP(j, f0, true)
            SR_KSTEP(j + 4, f1, true)
            SR_KSTEP(j + 8, f2, true)
        }
        if (j < j1) SR_KSTEP(j, f0, false)
        if (j + 4 < j1) SR_KSTEP(j + 4, f1, false)
    };
#undef SR_KSTEP
    auto store = [&](double* __restrict__ out, bool as_src) {
#pragma unroll
        for (int mb = 0; mb < 2; mb++) {
            if (mb == 1 && !two) break;
            const int pr = pair_s[mb * 8 + nq];
            if (pr < 0) continue;
            double* __restrict__ o = out + (size_t)(pr - a.pair_base) * a.ld_out + p_warp;
            const double* __restrict__ t_in =
                a.tau_out + (size_t)(pr - a.pair_base) * a.ld_out + p_warp;
            const bool divide = as_src && a.mode == 0;   // S = J/tau (tau written by this very lane)
            if (VEC) {
#pragma unroll
                for (int v = 0; v < NB / 4; v++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int off = 32 * v + 16 * e + 4 * kq;
                        if (p_warp + off >= a.n_pts) continue;   // rows are padded to ld_out
                        double x[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            x[i] = mb ? C1[4 * v + i][e] : C0[4 * v + i][e];
                            if (divide) {
                                const double t = t_in[off + i];
                                x[i] = (t == 0.0) ? 0.0 : x[i] / t;
                            }
                        }
                        if (a.f32) {
                            float* of = reinterpret_cast<float*>(out) + (size_t)(pr - a.pair_base) * a.ld_out +
                                        p_warp + off;
                            __stcs(reinterpret_cast<float4*>(of),
                                   make_float4((float)x[0], (float)x[1], (float)x[2], (float)x[3]));
                        } else if (a.keep) st_cg_v4(o + off, x); else st_cs_v4(o + off, x);
                    }
            } else {
#pragma unroll
                for (int i = 0; i < NB; i++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int off = 8 * i + 2 * kq + e;
                        if (p_warp + off >= a.n_pts) continue;
                        double x = mb ? C1[i][e] : C0[i][e];
                        if (divide) {
                            const double t = t_in[off];
                            x = (t == 0.0) ? 0.0 : x / t;
                        }
                        if (a.f32)
                            __stcs(reinterpret_cast<float*>(out) + (size_t)(pr - a.pair_base) * a.ld_out + p_warp + off,
                                   (float)x);
                        else if (a.keep) __stcg(o + off, x); else __stcs(o + off, x);
                    }
            }
        }
    };
    pass(0, n_tau);
    store(a.tau_out, false);
    pass(n_tau, n_tot);
    store(a.src_out, true);
    }
}


// ---------------------------------------------------------------------------------------------
// K3a+K3 fused (v4): the tensor-path product AND the layer recursion in one kernel; tau and J
// never leave the SM.
//
// A CTA owns a LOS group (FZ_LG lines of sight that the planner sorted to be alike) and a tile of
// FZ_TP grid points, and keeps the running intensities I[LOS][point] of the whole group in shared
// memory.  The planner cuts the group's (LOS, step) pairs into level-synchronous rounds - round r
// holds step r of every LOS - and inside a round into <= 16-pair chunks of one cell quad, so a LOS
// occurs at most once per chunk and its steps arrive in order.  Per chunk:
//   stage    one thread moves the chunk's fragment-ordered weights (<= 12.8 KB, contiguous), the
//            quad's row pointers, the 16 LOS slots and the row counts into shared memory with
//            TMA bulk copies (cp.async.bulk, completion on an mbarrier), one chunk ahead;
//   product  tau[16 pairs][32 points per warp] = W_tau * G as DMMA.8x8x4, LUT rows straight from
//            global memory (L2) as in k_los_mma;
//   update   e^-tau and phi(tau) (srdev::exp_phi), I' = I e^-tau with I from shared memory;
//   product  J = W_J * G over the emission rows;
//   update   I = I' + J phi(tau) back to shared memory.
// After the last chunk the group's radiances go to global memory once.  DRAM traffic = LUT rows
// (through L2) + weights + 8 B per (LOS, point); the 16 B per (LOS, step, point) layer round trip
// of the k_los_mma -> k_los_layers pair is gone.
// ---------------------------------------------------------------------------------------------
constexpr int FZ_LG = 64;            // LOS per CTA
constexpr int FZ_NT = 128;           // threads per CTA (4 warps)
constexpr int FZ_NB = 4;             // 8-point N blocks per warp: 32 points per warp
constexpr int FZ_TP = (FZ_NT / 32) * 8 * FZ_NB;   // 128 points per CTA
constexpr int FZ_SP = FZ_TP + 4;     // state row stride (doubles): rows 32-byte aligned, banks shifted

struct FuseArgs {
    const long long* rowptr;      // [n_groups][max_jp] device address of each program row
    const int4* chunk_meta;       // [n_chunks] {group, n_tau, n_tot, 0}
    const int* chunk_slot;        // [n_chunks][16] LOS slot in the CTA's group, -1 = padding
    const double* wfrag;          // [n_chunks][max_jp/4][2][32] fragment-ordered weights
    const int* lg_chunk;          // [n_lg + 1] chunk range of each LOS group
    const int* lg_row;            // [n_lg][FZ_LG] output row of the slot in `rad` (-1: unused slot)
    const int* lg_i0row;          // [n_lg][FZ_LG] row of the slot in `i0`
    int max_jp, lg0;              // lg0: first LOS group of this launch
    long pt0, n_pts;              // window of the LUT grid
    long ld_min;                  // smallest LUT row stride: loads are clamped below it
    long io_stride, io_off;       // rad / i0 rows: [row][io_stride], this window at io_off
    const double* i0;
    double* rad;
    int solo;
    int dbg;                      // ablation switches (SR_LOS_FDBG; results are wrong when set)
    int* flags;                   // LFLAG_* word of the call (a stuck mbarrier is reported there)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// TMA 1-D bulk copy global -> shared, completion counted on the mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int LD>
__global__ void __launch_bounds__(FZ_NT, 2) k_los_fused2(const __grid_constant__ FuseArgs a) {
    constexpr int NB = FZ_NB;
    extern __shared__ __align__(128) unsigned char fsm[];
    // layout: state | 2 x {wA, rps, slots, meta} | mbarriers
    double* state = reinterpret_cast<double*>(fsm);                          // [FZ_LG][FZ_SP]
    const size_t buf_bytes = (size_t)a.max_jp * (MMA_PB * sizeof(double) + sizeof(long long)) + 64 + 16;
    unsigned char* bufs = fsm + (size_t)FZ_LG * FZ_SP * sizeof(double);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(bufs + 2 * buf_bytes);
    auto wA_of = [&](int b) { return reinterpret_cast<double*>(bufs + b * buf_bytes); };
    auto rps_of = [&](int b) {
        return reinterpret_cast<long long*>(bufs + b * buf_bytes + (size_t)a.max_jp * MMA_PB * sizeof(double));
    };
    auto slot_of = [&](int b) {
        return reinterpret_cast<int*>(bufs + b * buf_bytes +
                                      (size_t)a.max_jp * (MMA_PB * sizeof(double) + sizeof(long long)));
    };
    auto meta_of = [&](int b) { return reinterpret_cast<int4*>(slot_of(b) + MMA_PB); };

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int lg = a.lg0 + blockIdx.x;
    const int c_beg = __ldg(a.lg_chunk + lg), c_end = __ldg(a.lg_chunk + lg + 1);
    const long p_cta = (long)blockIdx.y * FZ_TP;
    const long p_warp = p_cta + wid * (8 * NB);
    const bool warp_live = p_warp < a.n_pts;

    auto stage = [&](int c, int b) {   // one thread: chunk c -> buffer b
        const int4 m = __ldg(a.chunk_meta + c);
        const unsigned wb = (unsigned)m.z * MMA_PB * sizeof(double), rb = (unsigned)m.z * sizeof(long long);
        mbar_expect_tx(bars + b, wb + rb + 64 + 16);
        bulk_g2s(wA_of(b), a.wfrag + (size_t)c * a.max_jp * MMA_PB, wb, bars + b);
        bulk_g2s(rps_of(b), a.rowptr + (size_t)m.x * a.max_jp, rb, bars + b);
        bulk_g2s(slot_of(b), a.chunk_slot + (size_t)c * MMA_PB, 64, bars + b);
        bulk_g2s(meta_of(b), a.chunk_meta + c, 16, bars + b);
    };
    if (tid == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // running intensities of the group: I0 or zero
    {
        const int* __restrict__ i0row = a.lg_i0row + (size_t)lg * FZ_LG;
        for (int idx = tid; idx < FZ_LG * FZ_TP; idx += FZ_NT) {
            const int sl = idx / FZ_TP, p = idx - sl * FZ_TP;
            double v = 0.0;
            if (a.i0 && p_cta + p < a.n_pts) {
                const int row = __ldg(i0row + sl);
                if (row >= 0) v = a.i0[(size_t)row * a.io_stride + a.io_off + p_cta + p];
            }
            state[sl * FZ_SP + p] = v;
        }
    }
    __syncthreads();
    if (tid == 0 && c_beg < c_end) stage(c_beg, 0);

    const int kq = lane & 3, nq = lane >> 2;
    struct Frag { float v[NB]; };
    // per-lane point offset inside a LUT row (one 16-byte vector: points 4(nq>>1) + 16(nq&1) + 0..3);
    // lanes beyond the window re-read the last valid vector of the row (never stored)
    const long loff = min(a.pt0 + p_warp + 4 * (nq >> 1) + 16 * (nq & 1), a.ld_min - 4);
    double C0[NB][2], C1[NB][2];

    for (int c = c_beg; c < c_end; c++) {
        const int b = (c - c_beg) & 1;
        if (tid == 0 && c + 1 < c_end) stage(c + 1, b ^ 1);   // buffer b^1: chunk c-1 is done (barrier below)
        {   // wait for chunk c (phase parity of buffer b: it completes once every two chunks)
            const unsigned parity = (unsigned)(((c - c_beg) >> 1) & 1);
            unsigned spins = 0;
            while (!mbar_try_wait(bars + b, parity)) {
                if (++spins > (1u << 26)) {   // never in a correct run: report instead of hanging
                    if (lane == 0) atomicOr(a.flags, LFLAG_NONFINITE);
                    break;
                }
            }
        }
        if (warp_live) {
            const double* __restrict__ wA = wA_of(b);
            const long long* __restrict__ rps = rps_of(b);
            const int* __restrict__ slot_s = slot_of(b);
            const int4 meta = *meta_of(b);
            const int n_tau = meta.y, n_tot = meta.z;
            const bool two = slot_s[8] >= 0;
            auto load = [&](int j, Frag& f) {
                const float* __restrict__ r = reinterpret_cast<const float*>(rps[j + kq]);
                const float4 q = ld_row4<LD>(r + loff);
                f.v[0] = q.x; f.v[1] = q.y; f.v[2] = q.z; f.v[3] = q.w;
            };
            auto mma_step = [&](int j, const double (&bv)[NB]) {
                const double a0 = wA[(j >> 2) * 64 + lane];
                if (two) {
                    const double a1 = wA[(j >> 2) * 64 + 32 + lane];
#pragma unroll
                    for (int i = 0; i < NB; i++) {
                        dmma884(C0[i], a0, bv[i]);
                        dmma884(C1[i], a1, bv[i]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NB; i++) dmma884(C0[i], a0, bv[i]);
                }
            };
#define SR_FSTEP(J, F, RELOAD)                                            \
    {                                                                     \
        double b_[NB];                                                    \
        _Pragma("unroll") for (int i = 0; i < NB; i++) b_[i] = (double)F.v[i]; \
        if (RELOAD) load(min((J) + 20, j1 - 4), F);                       \
        mma_step((J), b_);                                                \
    }
            // rows [j0, j1): five B-fragment buffers rotate, every load has four k-steps to arrive
            // (two warps per scheduler here, against four in k_los_mma)
            auto pass = [&](int j0, int j1) {
#pragma unroll
                for (int i = 0; i < NB; i++) C0[i][0] = C0[i][1] = C1[i][0] = C1[i][1] = 0.0;
                if (j0 >= j1) return;
                Frag f0, f1, f2, f3, f4;
                load(j0, f0);
                load(min(j0 + 4, j1 - 4), f1);
                load(min(j0 + 8, j1 - 4), f2);
                load(min(j0 + 12, j1 - 4), f3);
                load(min(j0 + 16, j1 - 4), f4);
                int j = j0;
#pragma unroll 1
                for (; j + 20 <= j1; j += 20) {
                    SR_FSTEP(j, f0, true)
                    SR_FSTEP(j + 4, f1, true)
                    SR_FSTEP(j + 8, f2, true)
                    SR_FSTEP(j + 12, f3, true)
                    SR_FSTEP(j + 16, f4, true)
                }
                if (j < j1) SR_FSTEP(j, f0, false)
                if (j + 4 < j1) SR_FSTEP(j + 4, f1, false)
                if (j + 8 < j1) SR_FSTEP(j + 8, f2, false)
                if (j + 12 < j1) SR_FSTEP(j + 12, f3, false)
            };
#undef SR_FSTEP
            // C fragment of m-block mb, half e: 4 consecutive points at 16 e + 4 kq of pair nq
            double Ip[2][2][4], ph[2][2][4];
            int sl[2];
            sl[0] = slot_s[nq];
            sl[1] = two ? slot_s[8 + nq] : -1;
            if (!(a.dbg & 2)) pass(0, n_tau);
#pragma unroll
            for (int mb = 0; mb < 2; mb++) {
                if (sl[mb] < 0) continue;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const double* st = state + sl[mb] * FZ_SP + wid * (8 * NB) + 16 * e + 4 * kq;
                    const double2 q0 = *reinterpret_cast<const double2*>(st);
                    const double2 q1 = *reinterpret_cast<const double2*>(st + 2);
                    const double Iv[4] = {q0.x, q0.y, q1.x, q1.y};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        double ex, phi;
                        if (a.dbg & 1) { ex = 0.5; phi = mb ? C1[i][e] : C0[i][e]; }
                        else srdev::exp_phi(mb ? C1[i][e] : C0[i][e], ex, phi);
                        Ip[mb][e][i] = Iv[i] * ex;
                        ph[mb][e][i] = phi;
                    }
                }
            }
            if (!a.solo && !(a.dbg & 2)) pass(n_tau, n_tot);
#pragma unroll
            for (int mb = 0; mb < 2; mb++) {
                if (sl[mb] < 0) continue;
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    double* st = state + sl[mb] * FZ_SP + wid * (8 * NB) + 16 * e + 4 * kq;
                    double o[4];
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        o[i] = a.solo ? Ip[mb][e][i]
                                      : fma(mb ? C1[i][e] : C0[i][e], ph[mb][e][i], Ip[mb][e][i]);
                    *reinterpret_cast<double2*>(st) = make_double2(o[0], o[1]);
                    *reinterpret_cast<double2*>(st + 2) = make_double2(o[2], o[3]);
                }
            }
        }
        __syncthreads();   // chunk c done by every warp: its buffer may be refilled, I is visible
    }
    // the group's radiances leave the SM once
    {
        const int* __restrict__ orow = a.lg_row + (size_t)lg * FZ_LG;
        for (int sl = 0; sl < FZ_LG; sl++) {
            const int row = __ldg(orow + sl);
            if (row < 0) continue;
            if (p_cta + tid < a.n_pts)
                __stcs(a.rad + (size_t)row * a.io_stride + a.io_off + p_cta + tid, state[sl * FZ_SP + tid]);
        }
    }
}

__global__ void k_scatter_rows(const double* __restrict__ src, const int* __restrict__ dst_row, int n_rows,
                               int n_cols, double* __restrict__ dst) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)n_rows * n_cols) return;
    const int r = (int)(e / n_cols), c = (int)(e - (long)r * n_cols);
    dst[(size_t)dst_row[r] * n_cols + c] = src[e];
}

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
struct sr_lut {
    const float* g32 = nullptr;
    int n_cells = 0, n_sets = 0, mol = 0, iso = 0, lte_unidentified = 0;
    long n_grid = 0, row_stride = 0;
    double iso_ratio = 1.0;
    unsigned long long emis_mask = ~0ull;   // sr_lut_set_emission_mask
    sr_consts c{};
    std::vector<double> pt, Ps, Ts;
    std::vector<int> cellmap;
    sr::DevBuf<double> dPs, dTs, dq, delev;
    sr::DevBuf<int> dmap, drowmask, drowlist;
    int n_rows[3] = {0, 0, 0};
    sr::DevBuf<double> ws_rad[2], ws_i0;   // workspace of the host-buffer entry point
    sr::DevBuf<double> ws_tau, ws_src;  // layer scratch of the grouped K3a -> K3 path
    sr::DevBuf<double> ws_tau_g, ws_src_g, ws_jac, g_wfrag_g, g_dfrac;   // Jacobian path
    sr::DevBuf<char> g_prog;            // quad row programs / chunk tables of the current call
    sr::DevBuf<long long> g_rowptr;
    sr::DevBuf<double> g_wfrag;
    sr::DevBuf<int> g_ntau, g_ntot, g_cgrp, g_cpair;
    sr::DevBuf<int> g_cslot, g_cmeta, g_lgchunk, g_lgrow, g_lgi0, g_scat;   // fused path (k_los_fused2)
    sr::DevBuf<double> ws_low;
    cudaStream_t copy_stream = nullptr; // device -> host copies of the host-buffer entry point
    // One LOS call at a time per handle: the per-call scratch below lives in luts[0].  `mtx`
    // serialises the host side of concurrent callers; `busy` is recorded on the caller's stream at
    // the end of a call and waited for at the start of the next one, so that a call on another
    // stream cannot overwrite scratch that queued kernels still read.
    std::mutex mtx;
    cudaEvent_t busy = nullptr;
    ~sr_lut() {
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (busy) cudaEventDestroy(busy);
    }
    // per-call scratch (owned by the first LUT of a call)
    sr::DevBuf<int> cells, nsteps, flags;
    sr::DevBuf<double> W, temp, pres, column, tvib;
};

namespace {

void unique_sorted(const double* v, int n, int stride, std::vector<double>& out) {
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = v[i * stride];
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
}

// nearest and second-nearest node by |node - v| (LutSet.calculate, spect_main_module.py:1007-1040:
// np.argsort of the distances): ties go to the lower index, as a linear scan with a strict '<'
// gives.  `nodes` is sorted ascending and unique, so the nearest node is one of the two that
// bracket v and the second nearest is adjacent to the nearest.
void host_nearest_two(const std::vector<double>& nodes, double v, int& i1, int& i2) {
    const int n = (int)nodes.size();
    int hi = (int)(std::lower_bound(nodes.begin(), nodes.end(), v) - nodes.begin());   // first >= v
    int a;
    if (hi <= 0) a = 0;
    else if (hi >= n) a = n - 1;
    else a = (std::fabs(nodes[hi] - v) < std::fabs(nodes[hi - 1] - v)) ? hi : hi - 1;
    int b = -1;
    if (a - 1 >= 0) b = a - 1;
    if (a + 1 < n && (b < 0 || std::fabs(nodes[a + 1] - v) < std::fabs(nodes[b] - v))) b = a + 1;
    i1 = a;
    i2 = b;
}

GasDev gas_dev(const sr_lut* L) {
    GasDev g;
    g.g32 = L->g32;
    g.Ps = L->dPs.p;
    g.Ts = L->dTs.p;
    g.cellmap = L->dmap.p;
    g.qtab = L->dq.p;
    g.elev = L->delev.p;
    g.rowlist = L->drowlist.p;
    for (int ct = 0; ct < 3; ct++) g.n_rows[ct] = L->n_rows[ct];
    g.nP = (int)L->Ps.size();
    g.nT = (int)L->Ts.size();
    g.n_sets = L->n_sets;
    g.lte_unidentified = L->lte_unidentified;
    g.row_stride = L->row_stride;
    g.iso_ratio = L->iso_ratio;
    return g;
}

int lflags_to_status(int f) {
    if (f & LFLAG_EXTRAP_P) return sr::fail(SR_ERR_LUT, "Extrapolating in P (spect_main_module.py:1058)");
    if (f & LFLAG_NO_CELL) return sr::fail(SR_ERR_LUT, "LUT couple not found (spect_main_module.py:989-991)");
    if (f & LFLAG_NONFINITE) return sr::fail(SR_ERR_LUT, "non-finite LOS step weight (T, T_vib or column)");
    return SR_OK;
}

// uploads the step tables, runs k_step_weights; scratch lives in luts[0]
int prepare_steps(sr_lut* const* luts, const sr_los_steps* S, cudaStream_t st, LosArgs& la) {
    if (!luts || !S || S->n_gas < 1 || S->n_gas > MAX_GAS || S->n_los < 1 || S->n_steps_max < 1 ||
        !S->n_steps || !S->temp || !S->pres || !S->column)
        return sr::fail(SR_ERR_ARG, "LOS step tables: bad argument (n_gas must be 1..%d)", MAX_GAS);
    sr_lut* L0 = luts[0];
    int n_sets_max = 0;
    for (int m = 0; m < S->n_gas; m++) {
        if (!luts[m]) return sr::fail(SR_ERR_ARG, "LOS: missing LUT %d", m);
        if (luts[m]->n_grid != L0->n_grid) return sr::fail(SR_ERR_ARG, "LOS: LUT grids differ");
        n_sets_max = std::max(n_sets_max, luts[m]->n_sets);
    }
    if (S->tvib && S->n_sets_max < n_sets_max)
        return sr::fail(SR_ERR_ARG, "LOS: tvib table has n_sets_max=%d < %d", S->n_sets_max, n_sets_max);
    if (S->tvib) n_sets_max = S->n_sets_max;
    const size_t nls = (size_t)S->n_los * S->n_steps_max;
    for (int l = 0; l < S->n_los; l++)
        if (S->n_steps[l] < 0 || S->n_steps[l] > S->n_steps_max)
            return sr::fail(SR_ERR_ARG, "LOS %d: n_steps out of range", l);
    SR_CUDA(cudaStreamSynchronize(st));  // scratch of a previous call on this LUT may be in use
    SR_CUDA(L0->nsteps.upload(S->n_steps, S->n_los, st));
    SR_CUDA(L0->temp.upload(S->temp, nls, st));
    SR_CUDA(L0->pres.upload(S->pres, nls, st));
    SR_CUDA(L0->column.upload(S->column, nls * S->n_gas, st));
    if (S->tvib) SR_CUDA(L0->tvib.upload(S->tvib, nls * S->n_gas * n_sets_max, st));
    SR_CUDA(L0->cells.ensure(nls * S->n_gas * 4));
    SR_CUDA(L0->W.ensure(nls * S->n_gas * n_sets_max * 4));
    if (!L0->flags.p) {
        SR_CUDA(L0->flags.alloc(1));
        SR_CUDA(cudaMemsetAsync(L0->flags.p, 0, sizeof(int), st));
    }
    StepArgs sa;
    for (int m = 0; m < S->n_gas; m++) sa.gas[m] = gas_dev(luts[m]);
    sa.n_gas = S->n_gas;
    sa.n_los = S->n_los;
    sa.n_steps_max = S->n_steps_max;
    sa.n_sets_max = n_sets_max;
    sa.n_steps = L0->nsteps.p;
    sa.temp = L0->temp.p;
    sa.pres = L0->pres.p;
    sa.column = L0->column.p;
    sa.tvib = S->tvib ? L0->tvib.p : nullptr;
    sa.cells = L0->cells.p;
    sa.W = L0->W.p;
    sa.flags = L0->flags.p;
    sa.c2 = L0->c.h_cgs * L0->c.c_cgs / L0->c.k_cgs;
    dim3 grid((S->n_steps_max + 63) / 64, std::min(S->n_los, 65535), S->n_gas);
    SR_LAUNCH(k_step_weights, grid, 64, 0, st, sa);
    for (int m = 0; m < S->n_gas; m++) la.gas[m] = sa.gas[m];
    la.n_gas = S->n_gas;
    la.n_los = S->n_los;
    la.n_steps_max = S->n_steps_max;
    la.n_sets_max = n_sets_max;
    la.n_steps = L0->nsteps.p;
    la.cells = L0->cells.p;
    la.W = L0->W.p;
    la.n_grid = L0->n_grid;
    return SR_OK;
}

int check_lflags(sr_lut* L0, cudaStream_t st) {
    int f = 0;
    SR_CUDA(cudaMemcpyAsync(&f, L0->flags.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SR_CUDA(cudaStreamSynchronize(st));
    if (f) SR_CUDA(cudaMemsetAsync(L0->flags.p, 0, sizeof(int), st));
    return lflags_to_status(f);
}

}  // namespace

extern "C" {

int sr_lut_create(const float* g32_dev, const double* pt_host, int n_cells, int n_sets,
                  long n_grid, const double* level_energy_host, int mol, int iso,
                  double iso_ratio, int lte_unidentified, const sr_consts* consts, sr_lut** out) {
    return sr_lut_create_ld(g32_dev, n_grid, pt_host, n_cells, n_sets, n_grid, level_energy_host,
                            mol, iso, iso_ratio, lte_unidentified, consts, out);
}

int sr_lut_create_ld(const float* g32_dev, long row_stride, const double* pt_host, int n_cells,
                     int n_sets, long n_grid, const double* level_energy_host, int mol, int iso,
                     double iso_ratio, int lte_unidentified, const sr_consts* consts,
                     sr_lut** out) {
    if (!g32_dev || !pt_host || n_cells < 1 || n_sets < 1 || n_grid < 1 || !out ||
        row_stride < n_grid || (!lte_unidentified && !level_energy_host))
        return sr::fail(SR_ERR_ARG, "sr_lut_create: bad argument");
    double gi, t[TIPS_N], q[TIPS_N];
    int rc = sr_bd_tips_2003(mol, iso, &gi, t, q);
    if (rc) return rc;
    sr_lut* L = new sr_lut();
    L->g32 = g32_dev;
    L->n_cells = n_cells;
    L->n_sets = n_sets;
    L->n_grid = n_grid;
    L->row_stride = row_stride;
    L->mol = mol;
    L->iso = iso;
    L->iso_ratio = iso_ratio;
    L->lte_unidentified = lte_unidentified;
    if (consts) L->c = *consts; else sr_default_consts(&L->c);
    L->pt.assign(pt_host, pt_host + 2 * n_cells);
    unique_sorted(pt_host, n_cells, 2, L->Ps);
    unique_sorted(pt_host + 1, n_cells, 2, L->Ts);
    const int nP = (int)L->Ps.size(), nT = (int)L->Ts.size();
    L->cellmap.assign((size_t)nP * nT, -1);
    for (int i = n_cells - 1; i >= 0; i--) {  // first occurrence wins, like list.index (smm:993)
        int ip = (int)(std::lower_bound(L->Ps.begin(), L->Ps.end(), pt_host[2 * i]) - L->Ps.begin());
        int it = (int)(std::lower_bound(L->Ts.begin(), L->Ts.end(), pt_host[2 * i + 1]) - L->Ts.begin());
        L->cellmap[(size_t)ip * nT + it] = i;
    }
    std::vector<double> elev(n_sets, 0.0);
    if (!lte_unidentified) elev.assign(level_energy_host, level_energy_host + n_sets);
    auto body = [&]() -> int {
        SR_CUDA(L->dPs.upload(L->Ps.data(), nP));
        SR_CUDA(L->dTs.upload(L->Ts.data(), nT));
        SR_CUDA(L->dmap.upload(L->cellmap.data(), L->cellmap.size()));
        SR_CUDA(L->dq.upload(q, TIPS_N));
        SR_CUDA(L->delev.upload(elev.data(), n_sets));
        SR_CUDA(L->drowmask.alloc(n_sets));
        SR_CUDA(cudaMemset(L->drowmask.p, 0, sizeof(int) * n_sets));
        {
            dim3 grid(8, (unsigned)std::min<long>((long)n_cells * n_sets * 3, 65535));
            SR_LAUNCH(k_row_nonzero, grid, 256, 0, 0, g32_dev, n_cells, n_sets, n_grid, row_stride,
                      L->drowmask.p);
        }
        std::vector<int> mask(n_sets), list(3 * (size_t)n_sets, 0);
        SR_CUDA(cudaMemcpy(mask.data(), L->drowmask.p, sizeof(int) * n_sets, cudaMemcpyDeviceToHost));
        for (int ct = 0; ct < 3; ct++) {
            int n = 0;
            for (int s2 = 0; s2 < n_sets; s2++)
                if (mask[s2] & (1 << ct)) list[(size_t)ct * n_sets + n++] = s2;
            L->n_rows[ct] = n;
        }
        SR_CUDA(L->drowlist.upload(list.data(), list.size()));
        SR_CUDA(cudaDeviceSynchronize());
        return SR_OK;
    };
    rc = body();
    if (rc) { delete L; return rc; }
    *out = L;
    return SR_OK;
}

int sr_lut_destroy(sr_lut* lut) {
    delete lut;
    return SR_OK;
}

int sr_lut_set_emission_mask(sr_lut* lut, unsigned long long mask) {
    if (!lut) return sr::fail(SR_ERR_ARG, "sr_lut_set_emission_mask: bad argument");
    if (lut->n_sets > 64 && mask != ~0ull)
        return sr::fail(SR_ERR_LIMIT, "emission masks cover at most 64 sets");
    lut->emis_mask = mask;
    return SR_OK;
}

int sr_lut_weights(const double* pt_host, int n_cells, double pres, double temp, int cell[4],
                   double w[4]) {
    if (!pt_host || n_cells < 1 || !cell || !w) return sr::fail(SR_ERR_ARG, "sr_lut_weights: bad argument");
    std::vector<double> Ps, Ts;
    unique_sorted(pt_host, n_cells, 2, Ps);
    unique_sorted(pt_host + 1, n_cells, 2, Ts);
    auto find = [&](double p, double t) {
        for (int i = 0; i < n_cells; i++)
            if (pt_host[2 * i] == p && pt_host[2 * i + 1] == t) return i;
        return -1;
    };
    for (int i = 0; i < 4; i++) { cell[i] = -1; w[i] = 0.0; }
    if (Ts.size() < 2) return sr::fail(SR_ERR_LUT, "LUT has fewer than two temperatures");
    if (pres <= Ps.front()) {
        int ta, tb;
        host_nearest_two(Ts, temp, ta, tb);
        cell[0] = find(Ps[0], Ts[ta]);
        cell[1] = find(Ps[0], Ts[tb]);
        if (cell[0] < 0 || cell[1] < 0) return lflags_to_status(LFLAG_NO_CELL);
        w[0] = (Ts[tb] - temp) / (Ts[tb] - Ts[ta]);
        w[1] = (temp - Ts[ta]) / (Ts[tb] - Ts[ta]);
    } else if (pres <= Ps.back()) {
        if (Ps.size() < 2) return lflags_to_status(LFLAG_NO_CELL);
        int p1, p2, t1, t2;
        host_nearest_two(Ps, pres, p1, p2);
        host_nearest_two(Ts, temp, t1, t2);
        cell[0] = find(Ps[p1], Ts[t1]);
        cell[1] = find(Ps[p1], Ts[t2]);
        cell[2] = find(Ps[p2], Ts[t1]);
        cell[3] = find(Ps[p2], Ts[t2]);
        for (int i = 0; i < 4; i++)
            if (cell[i] < 0) return lflags_to_status(LFLAG_NO_CELL);
        const double wp1 = (Ps[p2] - pres) / (Ps[p2] - Ps[p1]), wp2 = (pres - Ps[p1]) / (Ps[p2] - Ps[p1]);
        const double wt1 = (Ts[t2] - temp) / (Ts[t2] - Ts[t1]), wt2 = (temp - Ts[t1]) / (Ts[t2] - Ts[t1]);
        w[0] = wt1 * wp1; w[1] = wt2 * wp1; w[2] = wt1 * wp2; w[3] = wt2 * wp2;
    } else {
        return lflags_to_status(LFLAG_EXTRAP_P);
    }
    return SR_OK;
}

static RecArgs rec_args(const double* tau, const double* src, const int* n_steps, int n_los,
                        int n_steps_max, long n_pts, const double* i0, int solo_absorption,
                        double* rad, int src_is_j, long io_stride, long io_off, long lay_stride,
                        int tile_pts) {
    RecArgs r;
    r.tau = tau;
    r.src = src;
    r.n_steps = n_steps;
    r.i0 = i0;
    r.rad = rad;
    r.n_pts = n_pts;
    r.io_stride = io_stride < 0 ? n_pts : io_stride;   // rad / i0 rows: [n_los][io_stride], window at io_off
    r.io_off = io_off;
    r.lay_stride = lay_stride < 0 ? n_pts : lay_stride;   // tau / src rows: [n_los][n_steps_max][lay_stride]
    r.n_steps_max = n_steps_max;
    r.n_tiles = (int)((n_pts + tile_pts - 1) / tile_pts);
    r.n_work = (long)r.n_tiles * n_los;
    r.solo = solo_absorption;
    r.src_is_j = src_is_j;
    r.keep = 0;
    r.f32 = 0;
    return r;
}

static int layers_launch(const double* tau, const double* src, const int* n_steps, int n_los,
                         int n_steps_max, long n_pts, const double* i0, int solo_absorption,
                         double* rad, cudaStream_t st, int src_is_j, long io_stride = -1,
                         long io_off = 0, long lay_stride = -1, int keep = 0, int f32 = 0) {
    // measured on B200 (tools/tune.py): (PPT=1, UNROLL=4) at 38 registers streams at the
    // measured copy bandwidth; wider variants lose occupancy
    int cfg = 5;
    if (const char* e = getenv("SR_K3_CFG")) cfg = atoi(e);   // tuning aid
    if (f32 && src_is_j && (lay_stride < 0 ? n_pts : lay_stride) % 4 == 0) {
        RecArgs r = rec_args(tau, src, n_steps, n_los, n_steps_max, n_pts, i0, solo_absorption, rad,
                             src_is_j, io_stride, io_off, lay_stride, 512);
        r.keep = keep;
        r.f32 = 1;
        static const int u32 = getenv("SR_K3F_UNROLL") ? atoi(getenv("SR_K3F_UNROLL")) : 4;
        if (u32 == 8) SR_LAUNCH((k_los_layers_f32<8>), dim3((unsigned)r.n_tiles, (unsigned)n_los), 256, 0, st, r);
        else if (u32 == 2) SR_LAUNCH((k_los_layers_f32<2>), dim3((unsigned)r.n_tiles, (unsigned)n_los), 256, 0, st, r);
        else SR_LAUNCH((k_los_layers_f32<4>), dim3((unsigned)r.n_tiles, (unsigned)n_los), 256, 0, st, r);
        return SR_OK;
    }
#define SR_K3_LAUNCH(PPT, UNROLL)                                                             \
    {                                                                                          \
        RecArgs r = rec_args(tau, src, n_steps, n_los, n_steps_max, n_pts, i0,                 \
                             solo_absorption, rad, src_is_j, io_stride, io_off,                \
                             lay_stride, 256 * PPT);                                           \
        r.keep = keep;                                                                         \
        r.f32 = f32;                                                                           \
        SR_LAUNCH((k_los_layers<PPT, UNROLL>), dim3((unsigned)r.n_tiles, (unsigned)n_los), 256, \
                  0, st, r);                                                                   \
    }
    switch (cfg) {
        case 1: SR_K3_LAUNCH(1, 8) break;
        case 2: SR_K3_LAUNCH(2, 4) break;
        case 3: SR_K3_LAUNCH(1, 16) break;
        case 4: SR_K3_LAUNCH(4, 4) break;
        case 5: SR_K3_LAUNCH(1, 4) break;
        default: SR_K3_LAUNCH(2, 8) break;
    }
#undef SR_K3_LAUNCH
    return SR_OK;
}

int sr_los_rt_layers_dev(const double* tau, const double* src, const int* n_steps, int n_los,
                         int n_steps_max, long n_pts, const double* i0, int solo_absorption,
                         double* rad, void* stream) {
    sr::ProfScope ps(SR_PROF_LOS_LAYERS, 0.0, (cudaStream_t)stream);   // work: the caller knows n_steps
    if (!tau || !src || !n_steps || !rad || n_los < 1 || n_steps_max < 1 || n_pts < 1)
        return sr::fail(SR_ERR_ARG, "sr_los_rt_layers_dev: bad argument");
    return layers_launch(tau, src, n_steps, n_los, n_steps_max, n_pts, i0, solo_absorption, rad,
                         (cudaStream_t)stream, 0);
}

// K3 + Jacobians.  tau_g/src_g == nullptr: the retrieved gas is the only absorber.
static int layers_jac_launch(const double* tau, const double* src, const double* tau_g,
                             const double* src_g, const double* dfrac, int n_par,
                             const int* n_steps, int n_los, int n_steps_max, long n_pts,
                             const double* i0, int solo, double* rad, double* jac, cudaStream_t st,
                             long io_stride = -1, long io_off = 0, long lay_stride = -1, int f32 = 0) {
    JacArgs a;
    a.r = rec_args(tau, src, n_steps, n_los, n_steps_max, n_pts, i0, solo, rad, 1, io_stride,
                   io_off, lay_stride, 256);
    a.r.f32 = f32;
    a.tau_g = tau_g;
    a.src_g = src_g;
    a.dfrac = dfrac;
    a.jac = jac;
    a.n_par = n_par;
    const bool multi = tau_g != nullptr;
#define SR_JAC(NP)                                                                             \
    {                                                                                          \
        dim3 grid((unsigned)a.r.n_tiles, (unsigned)n_los, (unsigned)((n_par + NP - 1) / NP));  \
        const size_t smem = (size_t)n_steps_max * NP * sizeof(double);                         \
        if (smem > (size_t)200 * 1024)                                                         \
            return sr::fail(SR_ERR_LIMIT, "LOS Jacobians: %d steps per LOS exceed the shared-memory " \
                            "table of the derivative kernel", n_steps_max);                    \
        if (multi) {                                                                           \
            SR_CUDA(cudaFuncSetAttribute(k_los_layers_jac<NP, true, 4, 2>,                     \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            SR_LAUNCH((k_los_layers_jac<NP, true, 4, 2>), grid, 256, smem, st, a);             \
        } else if (jcfg == 1) {                                                                \
            SR_CUDA(cudaFuncSetAttribute(k_los_layers_jac<NP, false, 2, 3>,                    \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            SR_LAUNCH((k_los_layers_jac<NP, false, 2, 3>), grid, 256, smem, st, a);            \
        } else if (jcfg == 2) {                                                                \
            SR_CUDA(cudaFuncSetAttribute(k_los_layers_jac<NP, false, 4, 3>,                    \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            SR_LAUNCH((k_los_layers_jac<NP, false, 4, 3>), grid, 256, smem, st, a);            \
        } else {                                                                               \
            SR_CUDA(cudaFuncSetAttribute(k_los_layers_jac<NP, false, 4, 2>,                    \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            SR_LAUNCH((k_los_layers_jac<NP, false, 4, 2>), grid, 256, smem, st, a);            \
        }                                                                                      \
    }
    static const int jcfg = getenv("SR_JAC_CFG") ? atoi(getenv("SR_JAC_CFG")) : 1;   // tuning aid; 1 (2-step unroll, 3 CTAs/SM) measured best on B200
    // the smallest accumulator count that takes all parameters in one pass, else chunks of 16
    if (n_par <= 4) SR_JAC(4) else if (n_par <= 8) SR_JAC(8) else if (n_par <= 12) SR_JAC(12)
    else SR_JAC(16)
#undef SR_JAC
    return SR_OK;
}

int sr_los_rt_layers_jac_dev(const double* tau, const double* emi, const double* tau_g,
                             const double* emi_g, const double* dfrac, int n_par,
                             const int* n_steps, int n_los, int n_steps_max, long n_pts,
                             const double* i0, int solo_absorption, double* rad, double* jac,
                             void* stream) {
    if (!tau || !emi || !dfrac || !n_steps || !rad || !jac || n_los < 1 || n_los > 65535 ||
        n_steps_max < 1 || n_pts < 1 || n_par < 1 || (tau_g == nullptr) != (emi_g == nullptr))
        return sr::fail(SR_ERR_ARG, "sr_los_rt_layers_jac_dev: bad argument");
    return layers_jac_launch(tau, emi, tau_g, emi_g, dfrac, n_par, n_steps, n_los, n_steps_max,
                             n_pts, i0, solo_absorption, rad, jac, (cudaStream_t)stream);
}

// Host side of the grouped K3a: choose the LUT cells of every (LOS, step) pair with the
// reference's rule (same code as sr_lut_weights), group the pairs by their cells ("quads"),
// build one row program per quad and 16-pair chunks.  Large batches are cut into LOS blocks so
// that the layer scratch stays bounded; quads (row programs) are shared by all blocks.
constexpr int KEY_LEN = 4 * MAX_GAS;
struct QuadKey {
    int c[KEY_LEN];
    bool operator<(const QuadKey& o) const { return memcmp(c, o.c, sizeof(c)) < 0; }
    bool operator==(const QuadKey& o) const { return memcmp(c, o.c, sizeof(c)) == 0; }
};

struct GemmPlan {
    std::vector<ProgEntry> prog;        // [n_groups][max_jp], padding rows have gas = -1
    std::vector<long long> rowptr;      // [n_groups][max_jp]
    std::vector<int> ntau, ntot;        // per group, multiples of 4
    std::vector<int> chunk_grp, chunk_pair;
    std::vector<int> blk_los, blk_chunk;   // [n_blocks+1] LOS / chunk range of each LOS block
    std::vector<int> nreal;             // per group: rows that are not padding
    std::vector<double> blk_rowpairs;   // per block: sum over chunks of real rows x valid pairs
    int max_jp = 4, n_groups = 0, n_chunks = 0;
    // shared by both chunkers
    std::vector<int> pair_grp;          // [n_los][n_steps_max] group (cell quad) of the pair, -1 = none
    std::vector<int> grp_order;         // groups in key order
    // fused path (k_los_fused2): LOS sorted to be alike, groups of FZ_LG, level-synchronous rounds
    std::vector<int> perm;              // sorted position -> LOS
    std::vector<int> chunk_slot;        // [n_chunks][16] slot of the LOS in its group
    std::vector<int4> chunk_meta;       // [n_chunks] {group, n_tau, n_tot, 0}
    std::vector<int> lg_chunk;          // [n_lg + 1]
    std::vector<double> lg_rowpairs;    // per LOS group: real rows x pairs
    double slots = 0.0, pairs = 0.0;    // compute slots (8-pair granular) and valid pairs
    int n_lg = 0;
};

static int cells_of(const sr_lut* L, double pres, double temp, int cell[4]) {
    for (int i = 0; i < 4; i++) cell[i] = -1;
    const std::vector<double>&Ps = L->Ps, &Ts = L->Ts;
    const int nT = (int)Ts.size();
    if (nT < 2) return lflags_to_status(LFLAG_NO_CELL);
    if (pres <= Ps.front()) {
        int ta, tb;
        host_nearest_two(Ts, temp, ta, tb);
        cell[0] = L->cellmap[ta];
        cell[1] = L->cellmap[tb];
        if (cell[0] < 0 || cell[1] < 0) return lflags_to_status(LFLAG_NO_CELL);
    } else if (pres <= Ps.back()) {
        if (Ps.size() < 2) return lflags_to_status(LFLAG_NO_CELL);
        int p1, p2, t1, t2;
        host_nearest_two(Ps, pres, p1, p2);
        host_nearest_two(Ts, temp, t1, t2);
        cell[0] = L->cellmap[(size_t)p1 * nT + t1];
        cell[1] = L->cellmap[(size_t)p1 * nT + t2];
        cell[2] = L->cellmap[(size_t)p2 * nT + t1];
        cell[3] = L->cellmap[(size_t)p2 * nT + t2];
        for (int i = 0; i < 4; i++)
            if (cell[i] < 0) return lflags_to_status(LFLAG_NO_CELL);
    } else {
        return lflags_to_status(LFLAG_EXTRAP_P);
    }
    return SR_OK;
}

// cell quad of every (LOS, step) pair and one row program per quad
static int plan_groups(sr_lut* const* luts, const sr_los_steps* S, GemmPlan& P) {
    const int n_gas = S->n_gas;
    const size_t nmax = (size_t)S->n_steps_max;
    // row lists per gas and ctype (all-zero spectra are skipped like the reference's None entries)
    std::vector<std::vector<int>> rl((size_t)n_gas * 3);
    int max_j_tau = 0, max_j_src = 0;
    for (int m = 0; m < n_gas; m++) {
        std::vector<int> list(3 * (size_t)luts[m]->n_sets);
        if (cudaMemcpy(list.data(), luts[m]->drowlist.p, list.size() * sizeof(int),
                       cudaMemcpyDeviceToHost) != cudaSuccess)
            return sr::fail(SR_ERR_CUDA, "LOS: cannot read the LUT row lists");
        for (int ct = 0; ct < 3; ct++)
            rl[(size_t)m * 3 + ct].assign(list.begin() + (size_t)ct * luts[m]->n_sets,
                                          list.begin() + (size_t)ct * luts[m]->n_sets + luts[m]->n_rows[ct]);
        max_j_tau += 4 * (luts[m]->n_rows[1] + luts[m]->n_rows[2]);
        max_j_src += 4 * luts[m]->n_rows[0];
    }
    P.max_jp = std::max(4, (max_j_tau + 3) / 4 * 4 + (max_j_src + 3) / 4 * 4);

    // ---- pass 1: group (cell quad) of every pair, found once for the whole batch ---------------
    // consecutive steps of a LOS mostly stay in the same quad: a one-entry cache in front of the map
    // (host threads over LOS ranges: with N GPUs every rank plans the whole batch, so this is the part
    // of a step that does not shrink with N)
    std::vector<QuadKey> keys;                 // per group
    std::vector<int>& pair_grp = P.pair_grp;
    pair_grp.assign((size_t)S->n_los * nmax, -1);
    {
        long n_pairs = 0;
        for (int l = 0; l < S->n_los; l++) n_pairs += S->n_steps[l];
        int n_thr = 1;
        if (n_pairs > 50000) n_thr = (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        if (const char* e = getenv("SR_LOS_PLAN_THREADS")) n_thr = std::max(1, atoi(e));
        n_thr = std::min(n_thr, std::max(1, S->n_los));
        struct Part {
            std::map<QuadKey, int> ids;
            std::vector<QuadKey> keys;
            int rc = SR_OK, bad_l = -1, bad_k = -1;
        };
        std::vector<Part> parts(n_thr);
        auto work = [&](int t) {
            Part& pt = parts[t];
            const int l0 = (int)((long)S->n_los * t / n_thr), l1 = (int)((long)S->n_los * (t + 1) / n_thr);
            QuadKey last;
            memset(last.c, 0x7f, sizeof(last.c));
            int last_id = -1;
            for (int l = l0; l < l1 && pt.rc == SR_OK; l++)
                for (int k = 0; k < S->n_steps[l]; k++) {
                    QuadKey key;
                    memset(key.c, 0xff, sizeof(key.c));
                    for (int m = 0; m < n_gas; m++) {
                        int rc = cells_of(luts[m], S->pres[l * nmax + k], S->temp[l * nmax + k], key.c + 4 * m);
                        if (rc) { pt.rc = rc; pt.bad_l = l; pt.bad_k = k; break; }
                    }
                    if (pt.rc) break;
                    int id;
                    if (last_id >= 0 && key == last) id = last_id;
                    else {
                        auto f = pt.ids.find(key);
                        if (f != pt.ids.end()) id = f->second;
                        else {
                            id = (int)pt.keys.size();
                            pt.ids.emplace(key, id);
                            pt.keys.push_back(key);
                        }
                        last = key;
                        last_id = id;
                    }
                    pair_grp[l * nmax + k] = id;       // local id, made global below
                }
        };
        if (n_thr == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < n_thr; t++) th.emplace_back(work, t);
            for (auto& x : th) x.join();
        }
        for (int t = 0; t < n_thr; t++)
            if (parts[t].rc) {   // repeat the failing look-up on this thread: sr_last_error is per thread
                int dummy[4];
                for (int m = 0; m < n_gas; m++) {
                    int rc = cells_of(luts[m], S->pres[parts[t].bad_l * nmax + parts[t].bad_k],
                                      S->temp[parts[t].bad_l * nmax + parts[t].bad_k], dummy);
                    if (rc) return rc;
                }
                return parts[t].rc;
            }
        std::map<QuadKey, int> group_of;
        std::vector<std::vector<int>> remap(n_thr);
        for (int t = 0; t < n_thr; t++) {
            remap[t].resize(parts[t].keys.size());
            for (size_t i = 0; i < parts[t].keys.size(); i++) {
                auto f = group_of.find(parts[t].keys[i]);
                if (f != group_of.end()) remap[t][i] = f->second;
                else {
                    remap[t][i] = P.n_groups;
                    group_of.emplace(parts[t].keys[i], P.n_groups++);
                    keys.push_back(parts[t].keys[i]);
                }
            }
        }
        auto fix = [&](int t) {
            const int l0 = (int)((long)S->n_los * t / n_thr), l1 = (int)((long)S->n_los * (t + 1) / n_thr);
            for (int l = l0; l < l1; l++)
                for (int k = 0; k < S->n_steps[l]; k++) pair_grp[l * nmax + k] = remap[t][pair_grp[l * nmax + k]];
        };
        if (n_thr == 1) fix(0);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < n_thr; t++) th.emplace_back(fix, t);
            for (auto& x : th) x.join();
        }
    }
    // ---- row programs, one per group ---------------------------------------------------------
    P.prog.resize((size_t)P.n_groups * P.max_jp);
    for (int grp = 0; grp < P.n_groups; grp++) {
        ProgEntry* pr = P.prog.data() + (size_t)grp * P.max_jp;
        int nj = 0;
        for (int pass = 0; pass < 2; pass++) {      // pass 0: tau rows (abs +, ind -); 1: J rows
            const int nj0 = nj;
            for (int m = 0; m < n_gas; m++) {
                const sr_lut* L = luts[m];
                for (int c = 0; c < 4; c++) {
                    const int cell = keys[grp].c[m * 4 + c];
                    if (cell < 0) continue;
                    for (int ct = (pass == 0 ? 1 : 0); ct <= (pass == 0 ? 2 : 0); ct++)
                        for (int s : rl[(size_t)m * 3 + ct]) {
                            ProgEntry pe;
                            pe.roff = (long long)(L->g32 + (((size_t)cell * L->n_sets + s) * 3 + ct) *
                                                               (size_t)L->row_stride);
                            pe.gas = m;
                            pe.widx = s * 4 + c;
                            pe.neg = (ct == 1);
                            pe.pad = 0;
                            pr[nj++] = pe;
                        }
                }
            }
            while ((nj - nj0) % 4) {                // zero-weight padding rows
                ProgEntry pe;
                pe.roff = (long long)luts[0]->g32;
                pe.gas = -1;
                pe.widx = 0;
                pe.neg = 0;
                pe.pad = 0;
                pr[nj++] = pe;
            }
            if (pass == 0) P.ntau.push_back(nj);
        }
        P.ntot.push_back(nj);
        int real = 0;
        for (int j = 0; j < nj; j++) real += pr[j].gas >= 0;
        P.nreal.push_back(real);
    }
    // groups in key order: neighbouring quads share cells, so their chunks share LUT rows in L2
    std::vector<int>& grp_order = P.grp_order;
    grp_order.resize(P.n_groups);
    for (int g = 0; g < P.n_groups; g++) grp_order[g] = g;
    std::sort(grp_order.begin(), grp_order.end(), [&](int a, int b) { return keys[a] < keys[b]; });
    P.rowptr.resize(P.prog.size());
    for (size_t q = 0; q < P.prog.size(); q++) P.rowptr[q] = P.prog[q].roff;
    return SR_OK;
}

// nl_block: LOS per block (the last block may be shorter)
static int build_plan(sr_lut* const* luts, const sr_los_steps* S, int nl_block, GemmPlan& P) {
    int rc = plan_groups(luts, S, P);
    if (rc) return rc;
    const size_t nmax = (size_t)S->n_steps_max;
    const std::vector<int>&pair_grp = P.pair_grp, &grp_order = P.grp_order;
    // ---- per LOS block a counting sort of its pairs by group, then 16-pair chunks --------------
    std::vector<int> cnt(P.n_groups + 1), fill(P.n_groups), sorted;
    P.blk_los.push_back(0);
    P.blk_chunk.push_back(0);
    for (int l0 = 0; l0 < S->n_los; l0 += nl_block) {
        const int l1 = std::min(S->n_los, l0 + nl_block);
        std::fill(cnt.begin(), cnt.end(), 0);
        size_t n_items = 0;
        for (int l = l0; l < l1; l++)
            for (int k = 0; k < S->n_steps[l]; k++) { cnt[pair_grp[l * nmax + k] + 1]++; n_items++; }
        for (int g = 0; g < P.n_groups; g++) cnt[g + 1] += cnt[g];
        for (int g = 0; g < P.n_groups; g++) fill[g] = cnt[g];
        sorted.resize(n_items);
        for (int l = l0; l < l1; l++)           // ascending pair index inside every group
            for (int k = 0; k < S->n_steps[l]; k++) sorted[fill[pair_grp[l * nmax + k]]++] = (int)(l * nmax + k);
        double rowpairs = 0.0;
        for (int gi = 0; gi < P.n_groups; gi++) {
            const int grp = grp_order[gi];
            const int i = cnt[grp], e = cnt[grp + 1];
            if (e <= i) continue;
            rowpairs += (double)P.nreal[grp] * (double)(e - i);
            for (int q = i; q < e; q += MMA_PB) {
                P.chunk_grp.push_back(grp);
                for (int t = 0; t < MMA_PB; t++) P.chunk_pair.push_back(q + t < e ? sorted[q + t] : -1);
                P.n_chunks++;
            }
        }
        P.blk_los.push_back(l1);
        P.blk_chunk.push_back(P.n_chunks);
        P.blk_rowpairs.push_back(rowpairs);
    }
    return SR_OK;
}

// Chunks for k_los_fused2.  LOS are sorted by (step count, quad sequence) so that the FZ_LG
// members of a LOS group walk through the same cell quads at the same step index; round r of a
// group holds step r of every member, cut into <= 16-pair chunks of one quad (a LOS at most once
// per chunk, its steps in order).  A chunk with <= 8 pairs costs one 8-row M block, not two.
static int build_fused_plan(sr_lut* const* luts, const sr_los_steps* S, GemmPlan& P) {
    int rc = plan_groups(luts, S, P);
    if (rc) return rc;
    const size_t nmax = (size_t)S->n_steps_max;
    const int n_los = S->n_los;
    const std::vector<int>& pair_grp = P.pair_grp;
    std::vector<int> rank_of(P.n_groups);               // group -> position in key order
    for (int i = 0; i < P.n_groups; i++) rank_of[P.grp_order[i]] = i;
    P.perm.resize(n_los);
    for (int l = 0; l < n_los; l++) P.perm[l] = l;
    std::sort(P.perm.begin(), P.perm.end(), [&](int x, int y) {
        const int nx = S->n_steps[x], ny = S->n_steps[y];
        if (nx != ny) return nx < ny;
        const int* px = pair_grp.data() + (size_t)x * nmax;
        const int* py = pair_grp.data() + (size_t)y * nmax;
        for (int k = 0; k < nx; k++)
            if (px[k] != py[k]) return rank_of[px[k]] < rank_of[py[k]];
        return x < y;
    });
    P.n_lg = (n_los + FZ_LG - 1) / FZ_LG;
    P.lg_chunk.assign(1, 0);
    std::vector<std::pair<int, int>> ready;             // (group rank, slot)
    for (int g = 0; g < P.n_lg; g++) {
        const int s0 = g * FZ_LG, ns = std::min(FZ_LG, n_los - s0);
        int max_r = 0;
        for (int sl = 0; sl < ns; sl++) max_r = std::max(max_r, S->n_steps[P.perm[s0 + sl]]);
        double rowpairs = 0.0;
        for (int r = 0; r < max_r; r++) {
            ready.clear();
            for (int sl = 0; sl < ns; sl++) {
                const int l = P.perm[s0 + sl];
                if (S->n_steps[l] > r) ready.emplace_back(rank_of[pair_grp[(size_t)l * nmax + r]], sl);
            }
            std::sort(ready.begin(), ready.end());
            size_t i = 0;
            while (i < ready.size()) {
                size_t e = i;
                while (e < ready.size() && ready[e].first == ready[i].first) e++;
                const int grp = P.grp_order[ready[i].first];
                rowpairs += (double)P.nreal[grp] * (double)(e - i);
                for (size_t q = i; q < e; q += MMA_PB) {
                    const int nv = (int)std::min<size_t>(MMA_PB, e - q);
                    P.chunk_grp.push_back(grp);
                    P.chunk_meta.push_back(make_int4(grp, P.ntau[grp], P.ntot[grp], 0));
                    for (int t = 0; t < MMA_PB; t++) {
                        const bool ok = t < nv;
                        const int sl = ok ? ready[q + t].second : -1;
                        P.chunk_slot.push_back(sl);
                        P.chunk_pair.push_back(ok ? (int)((size_t)P.perm[s0 + sl] * nmax + r) : -1);
                    }
                    P.slots += nv <= 8 ? 8.0 : 16.0;
                    P.pairs += nv;
                    P.n_chunks++;
                }
                i = e;
            }
        }
        P.lg_chunk.push_back(P.n_chunks);
        P.lg_rowpairs.push_back(rowpairs);
    }
    return SR_OK;
}

}  // extern "C"

constexpr int MMA_NB = 8;   // 8-point N blocks per warp: 64 points per warp, 256 per CTA

template <bool VEC, int MINB, int LD>
static int mma_launch_t(dim3 grid, size_t smem, cudaStream_t st, const MmaArgs& ma) {
    SR_CUDA(cudaFuncSetAttribute(k_los_mma<MMA_NB, VEC, MINB, LD>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SR_LAUNCH((k_los_mma<MMA_NB, VEC, MINB, LD>), grid, MMA_NT, smem, st, ma);
    return SR_OK;
}

static int mma_launch(bool vec, dim3 grid, size_t smem, cudaStream_t st, const MmaArgs& ma) {
    static const int minb = getenv("SR_MMA_MINB") ? atoi(getenv("SR_MMA_MINB")) : 4;   // tuning aids
    static const int ld = getenv("SR_MMA_LD") ? atoi(getenv("SR_MMA_LD")) : 0;
    if (!vec) return mma_launch_t<false, 3, 0>(grid, smem, st, ma);
    if (minb == 3) {
        if (ld == 1) return mma_launch_t<true, 3, 1>(grid, smem, st, ma);
        if (ld == 2) return mma_launch_t<true, 3, 2>(grid, smem, st, ma);
        return mma_launch_t<true, 3, 0>(grid, smem, st, ma);
    }
    if (ld == 1) return mma_launch_t<true, 4, 1>(grid, smem, st, ma);
    if (ld == 2) return mma_launch_t<true, 4, 2>(grid, smem, st, ma);
    return mma_launch_t<true, 4, 0>(grid, smem, st, ma);
}

extern "C" {

// host-buffer sink of los_launch: radiances leave the device block by block, chunk by chunk, on a
// second stream while the next chunk is computed
struct HostSink {
    double* rad_host;
    const double* i0_host;
};

// low-resolution sink: every LOS block is reduced to the instrument channels on the device
// (k_convolve_lowres over the block's full point window) and only [n_los][n_chan] survives
struct LowSink {
    const double* grid_dev;     // the spectral grid points of the window [pt0, pt0+n_pts)
    const double* centre_dev;
    const double* width_dev;
    int n_chan;
    double n_sigma;
    double* low_dev;            // [n_los][n_chan]
    int units;                  // SR_CHAN_* (sr_channels)
};

// analytic Jacobians of the batch (k_los_layers_jac)
struct JacSpec {
    int n_par;
    unsigned gas_mask;          // bit m: LUT m belongs to the retrieved gas
    const double* dfrac_host;   // [n_los][n_steps_max][n_par]
    double* jac_dev;            // hi-res [n_los][n_par][n_pts], or
    double* jac_low_dev;        // low-res [n_los][n_par][n_chan] (with a LowSink)
};

static int los_launch_locked(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                             const double* i0_dev, int solo, double* rad_dev, double* tau_dev,
                             double* src_dev, cudaStream_t st, int emit_j, HostSink* sink,
                             LowSink* low, JacSpec* jac);

}  // extern "C"

template <int LD>
static int fused_launch_t(dim3 grid, size_t smem, cudaStream_t st, const FuseArgs& fa) {
    SR_CUDA(cudaFuncSetAttribute(k_los_fused2<LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SR_LAUNCH((k_los_fused2<LD>), grid, FZ_NT, smem, st, fa);
    return SR_OK;
}

// v4 (k_los_fused2).  *done = false: not applicable (alignment, shared memory, or chunks too empty
// to pay: the batch is small or its LOS too unlike) - the caller takes the v3 path.
static int los_launch_fused(sr_lut* const* luts, const sr_los_steps* steps, const LosArgs& la, long pt0,
                            long n_pts, const double* i0_dev, int solo, double* rad_dev, cudaStream_t st,
                            LowSink* low, bool force, bool* done) {
    *done = false;
    sr_lut* L0 = luts[0];
    const int n_los = steps->n_los;
    const size_t nmax = (size_t)steps->n_steps_max;
    long ld_min = luts[0]->row_stride;
    bool aligned = pt0 % 4 == 0;
    for (int m = 0; m < steps->n_gas; m++) {
        aligned = aligned && luts[m]->row_stride % 4 == 0 && (size_t)luts[m]->g32 % 16 == 0;
        ld_min = std::min(ld_min, luts[m]->row_stride);
    }
    if (!aligned || ld_min < 4 || (!low && !rad_dev)) return SR_OK;
    if (!force && n_los < 2 * FZ_LG) return SR_OK;   // small batches: chunks cannot fill
    GemmPlan P;
    static const bool timing = getenv("SR_LOS_TIMING") != nullptr;
    const auto t_plan0 = std::chrono::steady_clock::now();
    int rc = build_fused_plan(luts, steps, P);
    if (rc) return rc;
    const double fill = P.slots > 0 ? P.pairs / P.slots : 0.0;
    static const double min_fill = getenv("SR_LOS_FUSE_FILL") ? atof(getenv("SR_LOS_FUSE_FILL")) : 0.6;
    if (timing)
        fprintf(stderr, "[sr_los] fused plan: %d LOS, %d LOS groups, %d quads, %d chunks, fill %.3f: %.3f ms\n",
                n_los, P.n_lg, P.n_groups, P.n_chunks, fill,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count());
    if (!force && fill < min_fill) return SR_OK;
    if (P.max_jp > GEMM_MAXJ)
        return sr::fail(SR_ERR_LIMIT, "LOS: %d LUT rows per cell quad (limit %d)", P.max_jp, GEMM_MAXJ);
    const size_t buf_bytes = (size_t)P.max_jp * (MMA_PB * sizeof(double) + sizeof(long long)) + 64 + 16;
    const size_t smem = (size_t)FZ_LG * FZ_SP * sizeof(double) + 2 * buf_bytes + 16;
    int dev = 0, smem_max = 0;
    SR_CUDA(cudaGetDevice(&dev));
    SR_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (smem > (size_t)smem_max) return SR_OK;
    // LOS blocks (whole LOS groups): only the low-res sink needs a radiance workspace
    int lg_block = P.n_lg;
    if (low) {
        size_t mem_free = 0, mem_total = 0;
        if (cudaMemGetInfo(&mem_free, &mem_total) != cudaSuccess) mem_free = (size_t)32 << 30;
        mem_free += L0->ws_rad[0].n * 8;
        size_t rad_cap = std::min<size_t>((size_t)8 << 30, std::max<size_t>((size_t)1 << 30, mem_free / 8));
        if (L0->ws_rad[0].n * 8 >= rad_cap / 2) rad_cap = std::min(L0->ws_rad[0].n * 8, (size_t)8 << 30);
        lg_block = (int)std::max<size_t>(1, rad_cap / ((size_t)n_pts * 8 * FZ_LG));
        if (const char* e = getenv("SR_LOS_BLOCK")) lg_block = std::max(1, atoi(e) / FZ_LG);
        lg_block = std::min(lg_block, P.n_lg);
    }
    lg_block = std::min(lg_block, 65535);
    // slot tables
    std::vector<int> lg_row((size_t)P.n_lg * FZ_LG, -1), lg_i0((size_t)P.n_lg * FZ_LG, -1);
    for (int pos = 0; pos < n_los; pos++) {
        const int g = pos / FZ_LG;
        lg_i0[pos] = P.perm[pos];
        lg_row[pos] = low ? (g % lg_block) * FZ_LG + pos % FZ_LG : P.perm[pos];
    }
    SR_CUDA(L0->g_lgrow.upload(lg_row.data(), lg_row.size(), st));
    SR_CUDA(L0->g_lgi0.upload(lg_i0.data(), lg_i0.size(), st));
    SR_CUDA(L0->g_lgchunk.upload(P.lg_chunk.data(), P.lg_chunk.size(), st));
    FuseArgs fa;
    if (P.n_chunks > 0) {
        SR_CUDA(L0->g_prog.upload(reinterpret_cast<const char*>(P.prog.data()),
                                  P.prog.size() * sizeof(ProgEntry), st));
        SR_CUDA(L0->g_rowptr.upload(P.rowptr.data(), P.rowptr.size(), st));
        SR_CUDA(L0->g_ntau.upload(P.ntau.data(), P.ntau.size(), st));
        SR_CUDA(L0->g_ntot.upload(P.ntot.data(), P.ntot.size(), st));
        SR_CUDA(L0->g_cgrp.upload(P.chunk_grp.data(), P.chunk_grp.size(), st));
        SR_CUDA(L0->g_cpair.upload(P.chunk_pair.data(), P.chunk_pair.size(), st));
        SR_CUDA(L0->g_cslot.upload(P.chunk_slot.data(), P.chunk_slot.size(), st));
        SR_CUDA(L0->g_cmeta.upload(reinterpret_cast<const int*>(P.chunk_meta.data()), P.chunk_meta.size() * 4, st));
        SR_CUDA(L0->g_wfrag.ensure((size_t)P.n_chunks * P.max_jp * MMA_PB));
        PackArgs pa;
        pa.prog = reinterpret_cast<const ProgEntry*>(L0->g_prog.p);
        pa.grp_ntot = L0->g_ntot.p;
        pa.chunk_grp = L0->g_cgrp.p;
        pa.chunk_pair = L0->g_cpair.p;
        pa.W = la.W;
        pa.wfrag = L0->g_wfrag.p;
        pa.n_pairs_tot = (long)n_los * (long)nmax;
        pa.n_sets_max = la.n_sets_max;
        pa.max_jp = P.max_jp;
        pa.n_chunks = P.n_chunks;
        pa.gas_mask = ~0u;
        pa.grp_ntau = L0->g_ntau.p;
        for (int m = 0; m < MAX_GAS; m++) pa.emis[m] = m < steps->n_gas ? luts[m]->emis_mask : ~0ull;
        const long n_el = (long)P.n_chunks * P.max_jp * MMA_PB;
        SR_LAUNCH(k_pack_wfrag, (unsigned)((n_el + 255) / 256), 256, 0, st, pa);
    }
    fa.rowptr = L0->g_rowptr.p;
    fa.chunk_meta = reinterpret_cast<const int4*>(L0->g_cmeta.p);
    fa.chunk_slot = L0->g_cslot.p;
    fa.wfrag = L0->g_wfrag.p;
    fa.lg_chunk = L0->g_lgchunk.p;
    fa.lg_row = L0->g_lgrow.p;
    fa.lg_i0row = L0->g_lgi0.p;
    fa.max_jp = P.max_jp;
    fa.pt0 = pt0;
    fa.n_pts = n_pts;
    fa.ld_min = ld_min;
    fa.io_stride = n_pts;
    fa.io_off = 0;
    fa.i0 = i0_dev;
    fa.solo = solo;
    fa.dbg = getenv("SR_LOS_FDBG") ? atoi(getenv("SR_LOS_FDBG")) : 0;
    fa.flags = L0->flags.p;
    static const int ld = getenv("SR_MMA_LD") ? atoi(getenv("SR_MMA_LD")) : 0;
    const unsigned n_tiles = (unsigned)((n_pts + FZ_TP - 1) / FZ_TP);
    if (low) {
        SR_CUDA(L0->ws_rad[0].ensure((size_t)lg_block * FZ_LG * n_pts));
        SR_CUDA(L0->ws_low.ensure((size_t)lg_block * FZ_LG * low->n_chan));
        SR_CUDA(L0->g_scat.upload(P.perm.data(), P.perm.size(), st));
    }
    for (int g0 = 0; g0 < P.n_lg; g0 += lg_block) {
        const int ng = std::min(lg_block, P.n_lg - g0);
        const int nl = std::min(n_los - g0 * FZ_LG, ng * FZ_LG);
        fa.lg0 = g0;
        fa.rad = low ? L0->ws_rad[0].p : rad_dev;
        double rowpairs = 0.0;
        for (int g = g0; g < g0 + ng; g++) rowpairs += P.lg_rowpairs[g];
        {
            sr::ProfScope ps(SR_PROF_LOS_FUSED, 2.0 * rowpairs * (double)n_pts, st);
            dim3 grid((unsigned)ng, n_tiles);
            rc = ld == 1 ? fused_launch_t<1>(grid, smem, st, fa)
                         : ld == 2 ? fused_launch_t<2>(grid, smem, st, fa) : fused_launch_t<0>(grid, smem, st, fa);
            if (rc) return rc;
        }
        if (low) {
            const sr_channels ch{low->n_chan, low->centre_dev, low->width_dev, low->n_sigma, low->units};
            {
                sr::ProfScope ps(SR_PROF_CONV, 8.0 * (double)nl * (double)n_pts, st);
                rc = sr_convolve_channels_dev(low->grid_dev, n_pts, L0->ws_rad[0].p, nl, &ch, L0->ws_low.p, st);
                if (rc) return rc;
            }
            const long n_el = (long)nl * low->n_chan;
            SR_LAUNCH(k_scatter_rows, (unsigned)((n_el + 255) / 256), 256, 0, st, L0->ws_low.p,
                      L0->g_scat.p + (size_t)g0 * FZ_LG, nl, low->n_chan, low->low_dev);
        }
    }
    *done = true;
    return SR_OK;
}

extern "C" {

static int los_launch(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                      const double* i0_dev, int solo, double* rad_dev, double* tau_dev,
                      double* src_dev, cudaStream_t st, int emit_j = 0, HostSink* sink = nullptr,
                      LowSink* low = nullptr, JacSpec* jac = nullptr) {
    if (!luts || !luts[0]) return sr::fail(SR_ERR_ARG, "LOS: missing LUT");
    sr_lut* L0 = luts[0];
    std::lock_guard<std::mutex> guard(L0->mtx);
    if (L0->busy) SR_CUDA(cudaEventSynchronize(L0->busy));
    else SR_CUDA(cudaEventCreateWithFlags(&L0->busy, cudaEventDisableTiming));
    const int rc = los_launch_locked(luts, steps, pt0, n_pts, i0_dev, solo, rad_dev, tau_dev, src_dev,
                                     st, emit_j, sink, low, jac);
    // a call that failed on the host must not leave its device flag word behind for the next one
    if (rc != SR_OK && L0->flags.p) cudaMemsetAsync(L0->flags.p, 0, sizeof(int), st);
    cudaEventRecord(L0->busy, st);
    return rc;
}

static int los_launch_locked(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                             const double* i0_dev, int solo, double* rad_dev, double* tau_dev,
                             double* src_dev, cudaStream_t st, int emit_j, HostSink* sink,
                             LowSink* low, JacSpec* jac) {
    LosArgs la;
    int rc = prepare_steps(luts, steps, st, la);
    if (rc) return rc;
    if (pt0 < 0 || n_pts < 1 || pt0 + n_pts > la.n_grid)
        return sr::fail(SR_ERR_ARG, "LOS: point range [%ld,%ld) outside the LUT grid", pt0, pt0 + n_pts);
    sr_lut* L0 = luts[0];
    const int n_los = steps->n_los;
    const size_t nmax = (size_t)steps->n_steps_max;
    int ver = 3;
    if (const char* e = getenv("SR_LOS_VER")) ver = atoi(e);   // 1 = thread-per-point fused kernel
    if (jac && (ver == 1 || sink || tau_dev))
        return sr::fail(SR_ERR_ARG, "LOS Jacobians need the grouped LOS path with device outputs");
    const unsigned all_gas = (steps->n_gas >= 32) ? ~0u : ((1u << steps->n_gas) - 1u);
    const bool jac_multi = jac && (jac->gas_mask & all_gas) != all_gas;
    if (ver == 1) {
        for (int m = 0; m < steps->n_gas; m++)
            if (luts[m]->emis_mask != ~0ull)
                return sr::fail(SR_ERR_ARG, "emission masks need the grouped LOS path");
        if (sink) {   // plain path: whole batch on the device, then one copy
            const size_t n = (size_t)n_los * n_pts;
            SR_CUDA(L0->ws_rad[0].ensure(n));
            if (sink->i0_host) SR_CUDA(L0->ws_i0.upload(sink->i0_host, n, st));
            i0_dev = sink->i0_host ? L0->ws_i0.p : nullptr;
            rad_dev = L0->ws_rad[0].p;
        }
        la.pt0 = pt0;
        la.n_pts = n_pts;
        la.i0 = i0_dev;
        la.rad = rad_dev;
        la.tau_out = tau_dev;
        la.src_out = src_dev;
        la.solo_absorption = solo;
        la.emit_j = emit_j;
        int G = 2, ppt = 4;   // measured on B200 (tools/tune.py fused)
        if (const char* e = getenv("SR_LOS_G")) G = std::max(1, std::min(4, atoi(e)));
        if (const char* e = getenv("SR_LOS_PPT")) ppt = atoi(e);
        dim3 block(256, G);
#define SR_FUSED(PPT)                                                                          \
    {                                                                                          \
        dim3 grid((unsigned)((n_los + G - 1) / G),                                             \
                  (unsigned)((n_pts + 256 * PPT - 1) / (256 * PPT)));                          \
        if (tau_dev) SR_LAUNCH((k_los_fused<PPT, true>), grid, block, 0, st, la);              \
        else SR_LAUNCH((k_los_fused<PPT, false>), grid, block, 0, st, la);                     \
    }
        if (ppt == 1) SR_FUSED(1) else if (ppt == 4) SR_FUSED(4) else SR_FUSED(2)
#undef SR_FUSED
        if (sink) {
            SR_CUDA(cudaStreamSynchronize(st));
            SR_CUDA(cudaMemcpy(sink->rad_host, rad_dev, (size_t)n_los * n_pts * sizeof(double),
                               cudaMemcpyDeviceToHost));
        }
        return SR_OK;
    }
    // ---- v4: product and recursion fused in one kernel (k_los_fused2) when the batch is large
    // enough to fill its chunks; otherwise (and for Jacobians, host sinks, materialised layers) v3
    if (!tau_dev && !jac && !sink && ver != 3) {
        bool done = false;
        rc = los_launch_fused(luts, steps, la, pt0, n_pts, i0_dev, solo, rad_dev, st, low, ver == 4, &done);
        if (rc) return rc;
        if (done) return SR_OK;
    }
    // ---- v3: grouped tensor-path product into layer arrays, then the streaming recursion ------
    // blocking: the layer scratch (16 B per pair and point) stays below the budget
    // Bigger LOS blocks mean more (LOS, step) pairs per LUT cell quad, i.e. fuller 16-pair chunks
    // in the tensor-path product and fewer passes over the LUT, so the budgets follow the free
    // device memory (a quarter of it for the layer scratch, capped at 32 GiB).
    size_t mem_free = 0, mem_total = 0;
    if (cudaMemGetInfo(&mem_free, &mem_total) != cudaSuccess) mem_free = (size_t)32 << 30;
    mem_free += L0->ws_tau.n * 8 + L0->ws_src.n * 8 + L0->ws_tau_g.n * 8 + L0->ws_src_g.n * 8 +
                L0->ws_jac.n * 8 + L0->ws_rad[0].n * 8 + L0->ws_rad[1].n * 8;   // our own, reusable
    size_t budget = std::min<size_t>((size_t)32 << 30, std::max<size_t>((size_t)1 << 30, mem_free / 4));
    // radiance workspace of a LOS block: with the float32 layer scratch the block size is limited by
    // this buffer, and bigger blocks fill the 16-pair chunks better
    static const size_t rad_max = (size_t)(getenv("SR_LOS_RADCAP_MB") ? atol(getenv("SR_LOS_RADCAP_MB")) : 16384) << 20;
    size_t rad_cap = std::min<size_t>(rad_max, std::max<size_t>((size_t)1 << 30, mem_free / 6));
    size_t jac_cap = std::min<size_t>((size_t)48 << 30, std::max<size_t>((size_t)1 << 30, mem_free / 3));
    if (L0->ws_rad[0].n * 8 >= rad_cap / 2) rad_cap = std::min(L0->ws_rad[0].n * 8, rad_max);
    if (L0->ws_jac.n * 8 >= jac_cap / 2) jac_cap = std::min(L0->ws_jac.n * 8, (size_t)48 << 30);
    // sticky: a workspace that already holds at least half of what a fresh budget would give is
    // used as it is, so that calls do not reallocate tens of GiB whenever the free memory moves
    const size_t have = std::min(L0->ws_tau.n, L0->ws_src.n) * 16;
    if (have >= budget / 2) budget = std::min(have, (size_t)32 << 30);
    if (const char* e = getenv("SR_LOS_SCRATCH_MB")) budget = (size_t)std::max(1L, atol(e)) << 20;
    long chunk_pts = n_pts;
    int nl_block = n_los;
    // float32 layer scratch: tau and J are rounded to the LUT's own storage precision between the
    // product and the recursion; every sum, exponential and the recursion itself stay FP64.  It
    // halves the HBM round trip of the layers: k_los_mma is no longer slowed down by its own stores
    // (0.66 -> 0.78 of the FP64 peak) and k_los_layers becomes FP64-bound.  Hi-res radiances move
    // by <= ~1e-7 relative, channel integrals by ~1e-9 (tests/test_gpu_los.py).  Default: on for the
    // low-resolution forward sink (results are channel integrals), off wherever hi-res radiances
    // or derivative spectra are returned (finite-difference checks and retrievals want a forward
    // model that is smooth to rounding).  SR_LOS_F32=0 / 1 forces it off / on for every path.
    const int f32_env = getenv("SR_LOS_F32") ? atoi(getenv("SR_LOS_F32")) : -1;   // (per call: tests switch it)
    const bool lay_f32 = !tau_dev && (f32_env > 0 || (f32_env < 0 && low && !jac));
    if (!tau_dev) {
        const long min_chunk = std::min<long>(n_pts, 65536);
        const size_t lay_b = (jac_multi ? 32 : 16) / (lay_f32 ? 2 : 1);   // scratch bytes per (pair, point)
        if ((size_t)n_los * nmax * lay_b * (size_t)min_chunk <= budget) {
            chunk_pts = (long)(budget / ((size_t)n_los * nmax * lay_b));
        } else {
            chunk_pts = min_chunk;
            nl_block = (int)std::max<size_t>(1, budget / (nmax * lay_b * (size_t)chunk_pts));
        }
        if (const char* e = getenv("SR_LOS_CHUNK")) chunk_pts = std::max(256L, atol(e));
        if (const char* e = getenv("SR_LOS_BLOCK")) nl_block = std::max(1, atoi(e));
        chunk_pts = std::min(chunk_pts, n_pts);
        if (chunk_pts < n_pts) chunk_pts = std::max(256L, chunk_pts / 256 * 256);
        if (low || sink)   // the block's radiances [nl_block][n_pts] live in a workspace (<= 8 GiB)
            nl_block = (int)std::min<size_t>((size_t)nl_block,
                                             std::max<size_t>(1, rad_cap / ((size_t)n_pts * 8)));
        if (low && jac)    // ... and so do its derivatives [nl_block][n_par][n_pts] (<= 48 GiB)
            nl_block = (int)std::min<size_t>(
                (size_t)nl_block,
                std::max<size_t>(1, jac_cap / ((size_t)n_pts * 8 * (size_t)jac->n_par)));
        nl_block = std::min(std::min(nl_block, n_los), 65535);   // blockIdx.y of the recursion
    }
    GemmPlan P;
    static const bool timing = getenv("SR_LOS_TIMING") != nullptr;   // host-side timing to stderr
    const auto t_plan0 = std::chrono::steady_clock::now();
    rc = build_plan(luts, steps, nl_block, P);
    if (rc) return rc;
    if (timing)
        fprintf(stderr, "[sr_los] plan: %d LOS, %d groups, %d chunks, nl_block %d, chunk_pts %ld: %.3f ms\n",
                n_los, P.n_groups, P.n_chunks, nl_block, chunk_pts,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_plan0).count());
    const int n_blocks = (int)P.blk_los.size() - 1;
    if (P.max_jp > GEMM_MAXJ)
        return sr::fail(SR_ERR_LIMIT, "LOS: %d LUT rows per cell quad (limit %d)", P.max_jp, GEMM_MAXJ);
    const size_t smem = (size_t)P.max_jp * (MMA_PB * sizeof(double) + sizeof(long long)) + 16;
    constexpr int TILE = (MMA_NT / 32) * 8 * MMA_NB;
    MmaArgs ma;
    bool rows_aligned = true;   // every LUT row starts on a 16-byte boundary
    ma.ld_min = luts[0]->row_stride;
    for (int m = 0; m < steps->n_gas; m++) {
        rows_aligned = rows_aligned && luts[m]->row_stride % 4 == 0 && (size_t)luts[m]->g32 % 16 == 0;
        ma.ld_min = std::min(ma.ld_min, luts[m]->row_stride);
    }
    if (const char* e = getenv("SR_LOS_NOVEC")) rows_aligned = rows_aligned && atoi(e) == 0;
    if (P.n_chunks > 0) {
        SR_CUDA(L0->g_prog.upload(reinterpret_cast<const char*>(P.prog.data()),
                                  P.prog.size() * sizeof(ProgEntry), st));
        SR_CUDA(L0->g_rowptr.upload(P.rowptr.data(), P.rowptr.size(), st));
        SR_CUDA(L0->g_ntau.upload(P.ntau.data(), P.ntau.size(), st));
        SR_CUDA(L0->g_ntot.upload(P.ntot.data(), P.ntot.size(), st));
        SR_CUDA(L0->g_cgrp.upload(P.chunk_grp.data(), P.chunk_grp.size(), st));
        SR_CUDA(L0->g_cpair.upload(P.chunk_pair.data(), P.chunk_pair.size(), st));
        SR_CUDA(L0->g_wfrag.ensure((size_t)P.n_chunks * P.max_jp * MMA_PB));
        PackArgs pa;
        pa.prog = reinterpret_cast<const ProgEntry*>(L0->g_prog.p);
        pa.grp_ntot = L0->g_ntot.p;
        pa.chunk_grp = L0->g_cgrp.p;
        pa.chunk_pair = L0->g_cpair.p;
        pa.W = la.W;
        pa.wfrag = L0->g_wfrag.p;
        pa.n_pairs_tot = (long)n_los * (long)nmax;
        pa.n_sets_max = la.n_sets_max;
        pa.max_jp = P.max_jp;
        pa.n_chunks = P.n_chunks;
        pa.gas_mask = ~0u;
        pa.grp_ntau = L0->g_ntau.p;
        for (int m = 0; m < MAX_GAS; m++) pa.emis[m] = m < steps->n_gas ? luts[m]->emis_mask : ~0ull;
        const long n_el = (long)P.n_chunks * P.max_jp * MMA_PB;
        SR_LAUNCH(k_pack_wfrag, (unsigned)((n_el + 255) / 256), 256, 0, st, pa);
        if (jac_multi) {   // second weight set: the retrieved gas alone
            SR_CUDA(L0->g_wfrag_g.ensure((size_t)P.n_chunks * P.max_jp * MMA_PB));
            pa.wfrag = L0->g_wfrag_g.p;
            pa.gas_mask = jac->gas_mask;
            SR_LAUNCH(k_pack_wfrag, (unsigned)((n_el + 255) / 256), 256, 0, st, pa);
        }
        ma.rowptr = L0->g_rowptr.p;
        ma.grp_ntau = L0->g_ntau.p;
        ma.grp_ntot = L0->g_ntot.p;
        ma.chunk_grp = L0->g_cgrp.p;
        ma.chunk_pair = L0->g_cpair.p;
        ma.wfrag = L0->g_wfrag.p;
        ma.max_jp = P.max_jp;
    }
    if (tau_dev) {   // materialise for the caller: [los][step][n_pts]; rows without a step stay as they are
        if (P.n_chunks == 0) return SR_OK;
        ma.chunk0 = 0;
        ma.pair_base = 0;
        ma.pt0 = pt0;
        ma.n_pts = n_pts;
        ma.ld_out = n_pts;
        ma.tau_out = tau_dev;
        ma.src_out = src_dev;
        ma.mode = emit_j;
        ma.keep = 0;
        ma.f32 = 0;
        ma.tpc = 1;
        dim3 grid((unsigned)P.n_chunks, (unsigned)((n_pts + TILE - 1) / TILE));
        const bool vec = rows_aligned && pt0 % 4 == 0 && n_pts % 4 == 0 &&
                         ((size_t)tau_dev | (size_t)src_dev) % 32 == 0;
        return mma_launch(vec, grid, smem, st, ma);
    }
    // Per (LOS block, wavenumber chunk): product into the layer scratch, then the streaming
    // recursion.  (Running the two concurrently - second stream, persistent recursion CTAs, or
    // recursion items inside the product CTAs - was measured slower on B200: the product keeps
    // the whole register file busy and the recursion needs >= 4 CTAs per SM of loads in flight.)
    const size_t blk_pairs = (size_t)nl_block * nmax;
    const long ld_lay = (chunk_pts + 3) / 4 * 4;   // scratch rows: 32-byte aligned for any chunk size
    SR_CUDA(L0->ws_tau.ensure(lay_f32 ? (blk_pairs * ld_lay + 1) / 2 : blk_pairs * ld_lay));
    SR_CUDA(L0->ws_src.ensure(lay_f32 ? (blk_pairs * ld_lay + 1) / 2 : blk_pairs * ld_lay));
    cudaEvent_t buf_free[2] = {nullptr, nullptr};   // host sink: copies of a radiance buffer done
    if (sink) {
        if (!L0->copy_stream)
            SR_CUDA(cudaStreamCreateWithFlags(&L0->copy_stream, cudaStreamNonBlocking));
        const int n_buf = n_blocks > 1 ? 2 : 1;
        for (int b = 0; b < n_buf; b++) SR_CUDA(L0->ws_rad[b].ensure((size_t)nl_block * n_pts));
        if (sink->i0_host) SR_CUDA(L0->ws_i0.ensure((size_t)nl_block * n_pts));
    }
    if (low) SR_CUDA(L0->ws_rad[0].ensure((size_t)nl_block * n_pts));
    if (jac) {
        SR_CUDA(L0->g_dfrac.upload(jac->dfrac_host, (size_t)n_los * nmax * jac->n_par, st));
        if (jac_multi) {
            SR_CUDA(L0->ws_tau_g.ensure(blk_pairs * ld_lay));
            SR_CUDA(L0->ws_src_g.ensure(blk_pairs * ld_lay));
        }
        if (low) SR_CUDA(L0->ws_jac.ensure((size_t)nl_block * jac->n_par * n_pts));
    }
    int status = SR_OK;
    static const int l2keep = getenv("SR_LOS_L2KEEP") ? atoi(getenv("SR_LOS_L2KEEP")) : 0;
    auto body = [&]() -> int {
        for (int b = 0; b < n_blocks; b++) {
            const int l0 = P.blk_los[b], nl = P.blk_los[b + 1] - l0;
            const int c_lo = P.blk_chunk[b], n_ch = P.blk_chunk[b + 1] - c_lo;
            double* rad_blk = rad_dev ? rad_dev + (size_t)l0 * n_pts : nullptr;
            const double* i0_blk = i0_dev ? i0_dev + (size_t)l0 * n_pts : nullptr;
            if (low) rad_blk = L0->ws_rad[0].p;
            if (sink) {
                rad_blk = L0->ws_rad[b & 1].p;
                if (buf_free[b & 1]) {   // the copies of block b-2 must have left this buffer
                    SR_CUDA(cudaStreamWaitEvent(st, buf_free[b & 1], 0));
                    SR_CUDA(cudaEventDestroy(buf_free[b & 1]));
                    buf_free[b & 1] = nullptr;
                }
                i0_blk = nullptr;
                if (sink->i0_host) {
                    SR_CUDA(cudaMemcpyAsync(L0->ws_i0.p, sink->i0_host + (size_t)l0 * n_pts,
                                            (size_t)nl * n_pts * sizeof(double),
                                            cudaMemcpyHostToDevice, st));
                    i0_blk = L0->ws_i0.p;
                }
            }
            for (long c0 = 0; c0 < n_pts; c0 += chunk_pts) {
                const long np = std::min(chunk_pts, n_pts - c0);
                if (n_ch > 0) {
                    ma.chunk0 = c_lo;
                    ma.pair_base = (long)l0 * (long)nmax;
                    ma.pt0 = pt0 + c0;
                    ma.n_pts = np;
                    ma.ld_out = ld_lay;
                    ma.tau_out = L0->ws_tau.p;
                    ma.src_out = L0->ws_src.p;
                    ma.mode = 1;
                    ma.keep = l2keep;
                    ma.f32 = lay_f32 ? 1 : 0;
                    // point tiles per CTA (the weights are staged once per CTA): up to 8, as long as
                    // the launch still has several waves of CTAs
                    static const int tpc_env = getenv("SR_MMA_TPC") ? std::max(1, atoi(getenv("SR_MMA_TPC"))) : 0;
                    const long n_t = (np + TILE - 1) / TILE;
                    ma.tpc = tpc_env ? tpc_env
                                     : (int)std::max<long>(1, std::min<long>(8, (long)n_ch * n_t / (148L * 4 * 4)));
                    dim3 grid((unsigned)n_ch, (unsigned)((n_t + ma.tpc - 1) / ma.tpc));
                    ma.wfrag = L0->g_wfrag.p;
                    int code;
                    {
                        sr::ProfScope ps(SR_PROF_LOS_MMA, 2.0 * P.blk_rowpairs[b] * (double)np, st);
                        code = mma_launch(rows_aligned && (pt0 + c0) % 4 == 0, grid, smem, st, ma);
                    }
                    if (code) return code;
                    if (jac_multi) {
                        ma.wfrag = L0->g_wfrag_g.p;
                        ma.tau_out = L0->ws_tau_g.p;
                        ma.src_out = L0->ws_src_g.p;   // (float32 like the first pair when lay_f32)
                        code = mma_launch(rows_aligned && (pt0 + c0) % 4 == 0, grid, smem, st, ma);
                        if (code) return code;
                    }
                }
                // radiances (and i0) rows have stride n_pts; this chunk is the window [c0, c0+np)
                int code;
                if (jac) {
                    double* jac_blk = low ? L0->ws_jac.p
                                          : jac->jac_dev + (size_t)l0 * jac->n_par * n_pts;
                    code = layers_jac_launch(L0->ws_tau.p, L0->ws_src.p,
                                             jac_multi ? L0->ws_tau_g.p : nullptr,
                                             jac_multi ? L0->ws_src_g.p : nullptr,
                                             L0->g_dfrac.p + (size_t)l0 * nmax * jac->n_par,
                                             jac->n_par, la.n_steps + l0, nl, steps->n_steps_max, np,
                                             i0_blk, solo, rad_blk, jac_blk, st, n_pts, c0, ld_lay,
                                             lay_f32 ? 1 : 0);
                } else {
                    double pairs = 0.0;
                    for (int l = l0; l < l0 + nl; l++) pairs += steps->n_steps[l];
                    sr::ProfScope ps(SR_PROF_LOS_LAYERS, ((lay_f32 ? 8.0 : 16.0) * pairs + 8.0 * nl) * (double)np, st);
                    code = layers_launch(L0->ws_tau.p, L0->ws_src.p, la.n_steps + l0, nl,
                                         steps->n_steps_max, np, i0_blk, solo, rad_blk, st, 1, n_pts,
                                         c0, ld_lay, l2keep, lay_f32 ? 1 : 0);
                }
                if (code) return code;
                if (sink) {
                    cudaEvent_t ev;
                    SR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    SR_CUDA(cudaEventRecord(ev, st));
                    SR_CUDA(cudaStreamWaitEvent(L0->copy_stream, ev, 0));
                    SR_CUDA(cudaEventDestroy(ev));
                    SR_CUDA(cudaMemcpy2DAsync(sink->rad_host + (size_t)l0 * n_pts + c0,
                                              (size_t)n_pts * sizeof(double), rad_blk + c0,
                                              (size_t)n_pts * sizeof(double), (size_t)np * sizeof(double),
                                              (size_t)nl, cudaMemcpyDeviceToHost, L0->copy_stream));
                }
            }
            if (sink && n_blocks > 1) {
                SR_CUDA(cudaEventCreateWithFlags(&buf_free[b & 1], cudaEventDisableTiming));
                SR_CUDA(cudaEventRecord(buf_free[b & 1], L0->copy_stream));
            }
            if (low) {
                const sr_channels ch{low->n_chan, low->centre_dev, low->width_dev, low->n_sigma,
                                     low->units};
                sr::ProfScope ps(SR_PROF_CONV, 8.0 * (double)nl * (double)n_pts, st);
                int code = sr_convolve_channels_dev(low->grid_dev, n_pts, rad_blk, nl, &ch,
                                                    low->low_dev + (size_t)l0 * low->n_chan, st);
                if (code) return code;
                if (jac) {
                    code = sr_convolve_channels_dev(low->grid_dev, n_pts, L0->ws_jac.p, nl * jac->n_par,
                                                    &ch,
                                                    jac->jac_low_dev + (size_t)l0 * jac->n_par * low->n_chan,
                                                    st);
                    if (code) return code;
                }
            }
        }
        if (sink) SR_CUDA(cudaStreamSynchronize(L0->copy_stream));
        return SR_OK;
    };
    status = body();
    for (int b = 0; b < 2; b++)
        if (buf_free[b]) cudaEventDestroy(buf_free[b]);
    return status;
}

int sr_los_rt_lut_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                      const double* i0_dev, int solo_absorption, double* rad_dev, void* stream) {
    if (!rad_dev) return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_dev: bad argument");
    return los_launch(luts, steps, pt0, n_pts, i0_dev, solo_absorption, rad_dev, nullptr, nullptr,
                      (cudaStream_t)stream);
}

int sr_los_rt_lut_lowres_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                             const double* grid_dev, const double* centre_dev,
                             const double* width_dev, int n_chan, double n_sigma,
                             const double* i0_dev, int solo_absorption, double* low_dev,
                             void* stream) {
    const sr_channels ch{n_chan, centre_dev, width_dev, n_sigma, SR_CHAN_SAME_UNITS};
    return sr_los_rt_lut_channels_dev(luts, steps, pt0, n_pts, grid_dev, &ch, i0_dev,
                                      solo_absorption, low_dev, stream);
}

int sr_los_rt_lut_channels_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                               const double* grid_dev, const sr_channels* ch, const double* i0_dev,
                               int solo_absorption, double* low_dev, void* stream) {
    if (!grid_dev || !ch || !ch->centre_dev || !ch->width_dev || !low_dev || ch->n_chan < 1 ||
        !(ch->n_sigma > 0.0))
        return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_channels_dev: bad argument");
    if (getenv("SR_LOS_VER") && atoi(getenv("SR_LOS_VER")) == 1)
        return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_channels_dev needs the grouped LOS path");
    LowSink low{grid_dev, ch->centre_dev, ch->width_dev, ch->n_chan, ch->n_sigma, low_dev, ch->units};
    return los_launch(luts, steps, pt0, n_pts, i0_dev, solo_absorption, nullptr, nullptr, nullptr,
                      (cudaStream_t)stream, 0, nullptr, &low);
}

static int jac_spec(const sr_los_steps* steps, int n_par, const int* gas_in_jac,
                    const double* dfrac_host, JacSpec& j) {
    if (!steps || n_par < 1 || !dfrac_host || steps->n_gas < 1 || steps->n_gas > MAX_GAS)
        return sr::fail(SR_ERR_ARG, "LOS Jacobians: bad argument");
    j.n_par = n_par;
    j.dfrac_host = dfrac_host;
    j.gas_mask = 0;
    for (int m = 0; m < steps->n_gas; m++)
        if (!gas_in_jac || gas_in_jac[m]) j.gas_mask |= 1u << m;
    if (!j.gas_mask) return sr::fail(SR_ERR_ARG, "LOS Jacobians: no LUT belongs to the retrieved gas");
    j.jac_dev = j.jac_low_dev = nullptr;
    return SR_OK;
}

int sr_los_rt_lut_jac_dev(sr_lut* const* luts, const sr_los_steps* steps, int n_par,
                          const int* gas_in_jac, const double* dfrac_host, long pt0, long n_pts,
                          const double* i0_dev, int solo_absorption, double* rad_dev,
                          double* jac_dev, void* stream) {
    if (!rad_dev || !jac_dev) return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_jac_dev: bad argument");
    JacSpec j;
    int rc = jac_spec(steps, n_par, gas_in_jac, dfrac_host, j);
    if (rc) return rc;
    j.jac_dev = jac_dev;
    return los_launch(luts, steps, pt0, n_pts, i0_dev, solo_absorption, rad_dev, nullptr, nullptr,
                      (cudaStream_t)stream, 0, nullptr, nullptr, &j);
}

int sr_los_rt_lut_jac_lowres_dev(sr_lut* const* luts, const sr_los_steps* steps, int n_par,
                                 const int* gas_in_jac, const double* dfrac_host, long pt0,
                                 long n_pts, const double* grid_dev, const double* centre_dev,
                                 const double* width_dev, int n_chan, double n_sigma,
                                 const double* i0_dev, int solo_absorption, double* low_dev,
                                 double* jac_low_dev, void* stream) {
    const sr_channels ch{n_chan, centre_dev, width_dev, n_sigma, SR_CHAN_SAME_UNITS};
    return sr_los_rt_lut_jac_channels_dev(luts, steps, n_par, gas_in_jac, dfrac_host, pt0, n_pts,
                                          grid_dev, &ch, i0_dev, solo_absorption, low_dev,
                                          jac_low_dev, stream);
}

int sr_los_rt_lut_jac_channels_dev(sr_lut* const* luts, const sr_los_steps* steps, int n_par,
                                   const int* gas_in_jac, const double* dfrac_host, long pt0,
                                   long n_pts, const double* grid_dev, const sr_channels* ch,
                                   const double* i0_dev, int solo_absorption, double* low_dev,
                                   double* jac_low_dev, void* stream) {
    if (!grid_dev || !ch || !ch->centre_dev || !ch->width_dev || !low_dev || !jac_low_dev ||
        ch->n_chan < 1 || !(ch->n_sigma > 0.0))
        return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_jac_channels_dev: bad argument");
    JacSpec j;
    int rc = jac_spec(steps, n_par, gas_in_jac, dfrac_host, j);
    if (rc) return rc;
    j.jac_low_dev = jac_low_dev;
    LowSink low{grid_dev, ch->centre_dev, ch->width_dev, ch->n_chan, ch->n_sigma, low_dev, ch->units};
    return los_launch(luts, steps, pt0, n_pts, i0_dev, solo_absorption, nullptr, nullptr, nullptr,
                      (cudaStream_t)stream, 0, nullptr, &low, &j);
}

int sr_los_tau_src_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                       double* tau_dev, double* src_dev, void* stream) {
    if (!tau_dev || !src_dev) return sr::fail(SR_ERR_ARG, "sr_los_tau_src_dev: bad argument");
    return los_launch(luts, steps, pt0, n_pts, nullptr, 0, nullptr, tau_dev, src_dev,
                      (cudaStream_t)stream);
}

int sr_los_abs_emi_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                       double* abs_dev, double* emi_dev, void* stream) {
    if (!abs_dev || !emi_dev) return sr::fail(SR_ERR_ARG, "sr_los_abs_emi_dev: bad argument");
    return los_launch(luts, steps, pt0, n_pts, nullptr, 0, nullptr, abs_dev, emi_dev,
                      (cudaStream_t)stream, 1);
}

int sr_los_check(sr_lut* const* luts, void* stream) {
    if (!luts || !luts[0] || !luts[0]->flags.p) return SR_OK;
    return check_lflags(luts[0], (cudaStream_t)stream);
}

int sr_los_rt_lut_host(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                       const double* i0_host, int solo_absorption, double* rad_host) {
    if (!rad_host || !steps) return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_host: bad argument");
    if (!luts || !luts[0]) return sr::fail(SR_ERR_ARG, "sr_los_rt_lut_host: missing LUT");
    sr_lut* L0 = luts[0];
    // grow-only workspaces inside the LUT handle: no cudaMalloc/cudaFree per call; the radiances
    // are copied out per (LOS block, wavenumber chunk) while the next chunk is computed
    HostSink sink{rad_host, i0_host};
    int rc = los_launch(luts, steps, pt0, n_pts, nullptr, solo_absorption, nullptr, nullptr, nullptr,
                        0, 0, &sink);
    if (rc) return rc;
    return check_lflags(L0, 0);
}

}  // extern "C"
