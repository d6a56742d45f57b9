// sr_steps.cu -- LOS geometry and radtran-step construction for whole batches on the device
// (SURVEY 8f row 4).  Replaces, for a batch of lines of sight, the per-LOS host chain
//   LineOfSight.calc_atm_intersections   (ray / atmosphere-shell intersections, samples every delta_x)
//   LineOfSight.calc_radtran_steps       (adaptive merge limited by max_T_variation / max_Plog_variation,
//                                         Curtis-Godson T, P, gas columns, per-level vibrational
//                                         temperatures; callers spect_main_module.py:2746-2767,
//                                         3133-3147; radtran_3D_ch4.py:200-202, 311)
//   curgod_fort_1..4                     curgods.f:2-98
// The reference's own code for the first two is in the missing spect_base_module; DESIGN.md 6.1 is
// the specification, spectrobot_b200/spect_base_module.py the host implementation this kernel
// pair must agree with.
//
// Kernels
//   k_steps_points     one thread per LOS: sample points (far end -> observer), T, P, number
//                      density per point, greedy merge into steps (bounds per step)
//   k_steps_integrals  one thread per (LOS, step): air column, Curtis-Godson T and P, gas columns,
//                      column-weighted vibrational temperatures, parameter derivative columns
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>
#include "sr_common.h"
#include "sr_device.cuh"

namespace {

constexpr double KB_HPA = 1.38065e-19;   // spect_classes.py:34 (P in hPa, n in cm-3)
constexpr int MAX_GAS_ST = 8;

struct AtmDev {
    int n_band, n_z, n_gas, n_sets_max, n_par, jac_gas, n_sza;
    const double* sza_nodes;   // [n_sza] degrees, ascending (n_sza > 1)
    const double* lat_edges;   // [n_band+1] (n_band > 1)
    const double* z;           // [n_z]
    const double* temp;        // [n_band][n_z]
    const double* lnpres;      // [n_band][n_z] log of P (hPa): P is log-linear in z
    const double* vmr;         // [n_gas][n_band][n_z]
    const double* tvib;        // [n_gas][n_sets_max][n_band][n_sza][n_z]
    const int* tvib_on;        // [n_gas][n_sets_max]: 1 own profile, 0 T_vib = step T, -1 no level
    const double* masks;       // [n_par][n_band][n_z]
    double radius, top;
};

// np.interp(x, xp, fp) on an ascending grid, clamped outside
__device__ __forceinline__ int interp_index(const double* __restrict__ xp, int n, double x) {
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (xp[m] <= x) lo = m; else hi = m;
    }
    return lo;
}
__device__ __forceinline__ double interp_at(const double* __restrict__ xp, const double* __restrict__ fp,
                                            int n, int j, double x) {
    if (x <= xp[0]) return fp[0];
    if (x >= xp[n - 1]) return fp[n - 1];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    return slope * (x - xp[j]) + fp[j];
}

struct PtArgs {
    AtmDev A;
    const double* origin;      // [n_los][3]
    const double* dir;         // [n_los][3]
    const double* sun;         // [n_los][3] unit vectors towards the Sun, or nullptr
    const double* sza_fixed;   // [n_los] degrees (use_tangent_sza), or nullptr
    int n_los, n_pts_max, n_steps_max, photon_order;
    double delta_x, max_dT, max_dlnP, max_tau;   // max_tau <= 0: no optical-depth limit
    double sigma[MAX_GAS_ST];  // peak cross-section estimate per gas (cm2 / molecule)
    // per-point scratch [n_los][n_pts_max]
    double *T, *P, *nd, *x, *alt, *sza;
    int *band, *jz;
    int* n_pts;                // [n_los]
    int* n_steps;              // [n_los] (true count, may exceed n_steps_max)
    int* bounds;               // [n_los][n_steps_max][2]
};

__global__ void k_steps_points(PtArgs a) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= a.n_los) return;
    const double ox = a.origin[3 * l], oy = a.origin[3 * l + 1], oz = a.origin[3 * l + 2];
    const double dx = a.dir[3 * l], dy = a.dir[3 * l + 1], dz = a.dir[3 * l + 2];
    const double st = -(ox * dx + oy * dy + oz * dz);
    const double tx = ox + st * dx, ty = oy + st * dy, tz = oz + st * dz;
    const double rt = sqrt(tx * tx + ty * ty + tz * tz);
    const double r_top = a.A.radius + a.A.top;
    a.n_pts[l] = 0;
    a.n_steps[l] = 0;
    if (rt >= r_top) return;
    const double half = sqrt(r_top * r_top - rt * rt);
    const double s_near = st - half;
    double s_far = st + half;
    if (rt < a.A.radius) s_far = st - sqrt(a.A.radius * a.A.radius - rt * rt);   // hits the surface
    const int kmax = (int)floor(half / a.delta_x - 1e-9);
    const size_t o = (size_t)l * a.n_pts_max;
    int n = 0;
    double s0 = 0.0;
    double sunx = 0.0, suny = 0.0, sunz = 0.0;
    if (a.sun) { sunx = a.sun[3 * l]; suny = a.sun[3 * l + 1]; sunz = a.sun[3 * l + 2]; }
    // greedy merge state: current step starts at point i0; running extrema over [i0, i]
    int i0 = 0, ns = 0;
    double tmin = 0, tmax = 0, pmin = 0, pmax = 0, t_prev = 0, lp_prev = 0;
    // optical-depth limit (max_opt_depth): running gas columns of the current step [i0, i] and of
    // its last segment, weighted by the peak cross-section estimates
    const bool lim_tau = a.max_tau > 0.0;
    double tau_run = 0.0, nd_prev = 0.0, x_prev = 0.0, v_prev[MAX_GAS_ST];
    auto close_step = [&](int last) {
        if (ns < a.n_steps_max) {
            a.bounds[((size_t)l * a.n_steps_max + ns) * 2] = i0;
            a.bounds[((size_t)l * a.n_steps_max + ns) * 2 + 1] = last;
        }
        ns++;
    };
    auto add_point = [&](double s) {
        const double px = ox + s * dx, py = oy + s * dy, pz = oz + s * dz;
        const double r = sqrt(px * px + py * py + pz * pz);
        const double alt = r - a.A.radius;
        int band = 0;
        if (a.A.n_band > 1) {
            const double lat = asin(pz / r) * (180.0 / M_PI);
            int b = 0;   // searchsorted(edges, lat, 'right') - 1, clipped to [0, n_band-1]
            while (b + 1 <= a.A.n_band && a.A.lat_edges[b + 1] <= lat) b++;
            band = min(max(b, 0), a.A.n_band - 1);
        }
        const int j = interp_index(a.A.z, a.A.n_z, alt);
        const double T = interp_at(a.A.z, a.A.temp + (size_t)band * a.A.n_z, a.A.n_z, j, alt);
        const double lnP = interp_at(a.A.z, a.A.lnpres + (size_t)band * a.A.n_z, a.A.n_z, j, alt);
        const double P = exp(lnP);
        const double nd = P / (KB_HPA * T);
        if (n == 0) s0 = s;
        const double x = fabs(s0 - s) * 1.e5;     // path length from the first sample, cm
        // solar zenith angle of the sample (LineOfSight.calc_SZA_along_los, or the tangent-point
        // value everywhere with use_tangent_sza, smm:3138-3141)
        double sza = 0.0;
        if (a.sza_fixed) sza = a.sza_fixed[l];
        else if (a.sun) {
            const double c = (px * sunx + py * suny + pz * sunz) / r;
            sza = acos(fmin(1.0, fmax(-1.0, c))) * (180.0 / M_PI);
        }
        if (n < a.n_pts_max) {
            a.T[o + n] = T;
            a.P[o + n] = P;
            a.nd[o + n] = nd;
            a.x[o + n] = x;
            a.alt[o + n] = alt;
            a.sza[o + n] = sza;
            a.band[o + n] = band;
            a.jz[o + n] = j;
        }
        // merge rule of calc_radtran_steps: close the step at i-1 when the range over [i0, i]
        // exceeds a limit and the step has at least two segments
        const double lp = log(P);
        const int i = n;
        double tau_seg = 0.0;
        if (lim_tau) {
            for (int m = 0; m < a.A.n_gas; m++) {
                const double v = interp_at(a.A.z, a.A.vmr + ((size_t)m * a.A.n_band + band) * a.A.n_z,
                                           a.A.n_z, j, alt);
                if (i > 0) tau_seg += a.sigma[m] * srdev::curgod_seg2(nd_prev, nd, v_prev[m], v, x - x_prev);
                v_prev[m] = v;
            }
            tau_run += tau_seg;
        }
        if (i == 0) { tmin = tmax = T; pmin = pmax = lp; }
        else {
            tmin = fmin(tmin, T); tmax = fmax(tmax, T);
            pmin = fmin(pmin, lp); pmax = fmax(pmax, lp);
            if ((tmax - tmin > a.max_dT || pmax - pmin > a.max_dlnP || (lim_tau && tau_run > a.max_tau)) &&
                i - i0 >= 2) {
                close_step(i - 1);
                i0 = i - 1;
                tmin = fmin(t_prev, T); tmax = fmax(t_prev, T);
                pmin = fmin(lp_prev, lp); pmax = fmax(lp_prev, lp);
                tau_run = tau_seg;
            }
        }
        t_prev = T;
        lp_prev = lp;
        nd_prev = nd;
        x_prev = x;
        n++;
    };
    if (!a.photon_order) {   // far end -> observer: the order of the layer recursion
        add_point(s_far);
        for (int k = kmax; k >= -kmax; k--) {
            const double s = st + a.delta_x * (double)k;
            if (s < s_far - 1e-6 && s > s_near + 1e-6) add_point(s);
        }
        add_point(s_near);
    } else {                 // LOS_order = 'photon' (invert_LOS_direction): observer side first
        add_point(s_near);
        for (int k = -kmax; k <= kmax; k++) {
            const double s = st + a.delta_x * (double)k;
            if (s < s_far - 1e-6 && s > s_near + 1e-6) add_point(s);
        }
        add_point(s_far);
    }
    if (n >= 2) close_step(n - 1);
    a.n_pts[l] = n;
    a.n_steps[l] = ns;
}

struct IntArgs {
    AtmDev A;
    int n_los, n_pts_max, n_steps_max;
    const double *T, *P, *nd, *x, *alt, *sza;
    const int *band, *jz, *n_steps, *bounds;
    // outputs, sr_los_steps layout
    double* temp;     // [n_los][n_steps_max]
    double* pres;
    double* column;   // [n_gas][n_los][n_steps_max]
    double* tvib;     // [n_gas][n_sets_max][n_los][n_steps_max]
    double* dfrac;    // [n_los][n_steps_max][n_par]
};

// Threads = (step k, task lane ty) of one LOS.  The integrals of a step are independent tasks -
// air column, Curtis-Godson T, Curtis-Godson P, one column per gas, then (they need the columns)
// one vibrational temperature per (gas, level) and one derivative per parameter - and the lanes of
// a step share them round-robin, with the columns and T passed through shared memory.
constexpr int INT_TX = 32;    // steps per block
constexpr int INT_TY = 8;     // task lanes per step

__global__ void __launch_bounds__(INT_TX * INT_TY) k_steps_integrals(IntArgs a) {
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int k = blockIdx.x * INT_TX + tx;
    const int l = blockIdx.y;
    const AtmDev& A = a.A;
    __shared__ double s_col[MAX_GAS_ST][INT_TX];
    __shared__ double s_air[INT_TX], s_ct[INT_TX], s_cp[INT_TX];
    const bool in_table = k < a.n_steps_max;
    const size_t sk = (size_t)l * a.n_steps_max + (in_table ? k : 0);
    const size_t nls = (size_t)a.n_los * a.n_steps_max;
    const bool real = in_table && k < min(a.n_steps[l], a.n_steps_max);
    if (in_table && !real && ty == 0) {   // padding, as LineOfSight.step_tables fills it
        a.temp[sk] = 100.0;
        a.pres[sk] = 1.e-6;
        for (int m = 0; m < A.n_gas; m++) {
            a.column[(size_t)m * nls + sk] = 0.0;
            for (int s = 0; s < A.n_sets_max; s++)
                a.tvib[((size_t)m * A.n_sets_max + s) * nls + sk] = 100.0;
        }
        for (int q = 0; q < A.n_par; q++) a.dfrac[sk * A.n_par + q] = 0.0;
    }
    const int ia = real ? a.bounds[sk * 2] : 0, ie = real ? a.bounds[sk * 2 + 1] : 0;
    const size_t o = (size_t)l * a.n_pts_max;
    const double* __restrict__ nd = a.nd + o;
    const double* __restrict__ x = a.x + o;
    auto prof = [&](const double* __restrict__ table, int i) {   // profile value at sample point i
        return interp_at(A.z, table + (size_t)a.band[o + i] * A.n_z, A.n_z, a.jz[o + i], a.alt[o + i]);
    };
    // vibrational temperature table [band][sza][z]: linear in altitude, then linear between the two
    // SZA nodes that bracket the sample's solar zenith angle (clamped outside the node range)
    auto prof_sza = [&](const double* __restrict__ table, int i) {
        const double* __restrict__ tb = table + (size_t)a.band[o + i] * A.n_sza * A.n_z;
        if (A.n_sza <= 1) return interp_at(A.z, tb, A.n_z, a.jz[o + i], a.alt[o + i]);
        const double sz = a.sza[o + i];
        const int js = interp_index(A.sza_nodes, A.n_sza, sz);
        const double f0 = interp_at(A.z, tb + (size_t)js * A.n_z, A.n_z, a.jz[o + i], a.alt[o + i]);
        const double f1 = interp_at(A.z, tb + (size_t)(js + 1) * A.n_z, A.n_z, a.jz[o + i], a.alt[o + i]);
        if (sz <= A.sza_nodes[0]) return interp_at(A.z, tb, A.n_z, a.jz[o + i], a.alt[o + i]);
        if (sz >= A.sza_nodes[A.n_sza - 1])
            return interp_at(A.z, tb + (size_t)(A.n_sza - 1) * A.n_z, A.n_z, a.jz[o + i], a.alt[o + i]);
        const double w = (sz - A.sza_nodes[js]) / (A.sza_nodes[js + 1] - A.sza_nodes[js]);
        return (1.0 - w) * f0 + w * f1;
    };
    // ---- phase A: air column (curgod_fort_1), Curtis-Godson T and P (curgod_fort_4 with vmr = 1),
    // gas columns (curgod_fort_2) ---------------------------------------------------------------
    const int n_a = 3 + A.n_gas;
    if (real)
        for (int t = ty; t < n_a; t += INT_TY) {
            double acc = 0.0;
            if (t == 0) {
                for (int i = ia; i < ie; i++) acc += srdev::curgod_seg1(nd[i], nd[i + 1], x[i + 1] - x[i]);
                s_air[tx] = acc;
            } else if (t == 1) {
                for (int i = ia; i < ie; i++)
                    acc += srdev::curgod_seg4(nd[i], nd[i + 1], 1.0, 1.0, a.T[o + i], a.T[o + i + 1], x[i + 1] - x[i]);
                s_ct[tx] = acc;
            } else if (t == 2) {
                for (int i = ia; i < ie; i++)
                    acc += srdev::curgod_seg4(nd[i], nd[i + 1], 1.0, 1.0, a.P[o + i], a.P[o + i + 1], x[i + 1] - x[i]);
                s_cp[tx] = acc;
            } else {
                const int m = t - 3;
                const double* vt = A.vmr + (size_t)m * A.n_band * A.n_z;
                double v0 = prof(vt, ia);
                for (int i = ia; i < ie; i++) {
                    const double v1 = prof(vt, i + 1);
                    acc += srdev::curgod_seg2(nd[i], nd[i + 1], v0, v1, x[i + 1] - x[i]);
                    v0 = v1;
                }
                s_col[m][tx] = acc;
                a.column[(size_t)m * nls + sk] = acc;
            }
        }
    __syncthreads();
    if (!real) return;
    const double t_cg = s_ct[tx] / s_air[tx];
    if (ty == 0) {
        a.temp[sk] = t_cg;
        a.pres[sk] = s_cp[tx] / s_air[tx];
    }
    // ---- phase B: vibrational temperatures (curgod_fort_3) and parameter derivatives -----------
    const int n_tv = A.n_gas * A.n_sets_max;
    const bool jac_ok = A.jac_gas >= 0 && A.jac_gas < A.n_gas;
    for (int t = ty; t < n_tv + A.n_par; t += INT_TY) {
        if (t < n_tv) {
            const int m = t / A.n_sets_max, sidx = t - m * A.n_sets_max;
            const double* vt = A.vmr + (size_t)m * A.n_band * A.n_z;
            const int on = A.tvib_on[m * A.n_sets_max + sidx];
            double tv = 100.0;
            if (on == 0) tv = t_cg;
            else if (on > 0) {
                const double* tt = A.tvib + ((size_t)m * A.n_sets_max + sidx) * A.n_band * A.n_sza * A.n_z;
                double acc = 0.0, w0 = prof(vt, ia), f0 = prof_sza(tt, ia);
                for (int i = ia; i < ie; i++) {
                    const double w1 = prof(vt, i + 1), f1 = prof_sza(tt, i + 1);
                    acc += srdev::curgod_seg3(nd[i], nd[i + 1], w0, w1, f0, f1, x[i + 1] - x[i]);
                    w0 = w1;
                    f0 = f1;
                }
                tv = acc / s_col[m][tx];
            }
            a.tvib[((size_t)m * A.n_sets_max + sidx) * nls + sk] = tv;
        } else {
            const int q = t - n_tv;
            double out = 0.0;
            if (jac_ok) {
                const double col = s_col[A.jac_gas][tx];
                const double* mk = A.masks + (size_t)q * A.n_band * A.n_z;   // per latitude box
                double d = 0.0;
                double m0 = prof(mk, ia);
                bool any = m0 != 0.0;
                for (int i = ia; i < ie; i++) {
                    const double m1 = prof(mk, i + 1);
                    any = any || m1 != 0.0;
                    d += srdev::curgod_seg2(nd[i], nd[i + 1], m0, m1, x[i + 1] - x[i]);
                    m0 = m1;
                }
                out = (any && col != 0.0) ? d / col : 0.0;
            }
            a.dfrac[sk * A.n_par + q] = out;
        }
    }
}

}  // namespace

extern "C" {

int sr_los_steps_build(const sr_atmosphere* atm, int n_los, const double* origin,
                       const double* direction, double delta_x_km, double max_T_variation,
                       double max_Plog_variation, int n_par, const double* masks, int jac_gas,
                       int n_steps_max, int* n_steps, double* temp, double* pres, double* column,
                       double* tvib, double* dfrac, int* n_steps_needed) {
    sr_los_rays rays{n_los, origin, direction, nullptr, nullptr};
    sr_steps_opt opt{delta_x_km, max_T_variation, max_Plog_variation, 0.0, nullptr, 0};
    return sr_los_steps_build_rays(atm, &rays, &opt, n_par, masks, jac_gas, n_steps_max, n_steps, temp,
                                   pres, column, tvib, dfrac, n_steps_needed);
}

int sr_los_steps_build_rays(const sr_atmosphere* atm, const sr_los_rays* rays, const sr_steps_opt* opt,
                            int n_par, const double* masks, int jac_gas, int n_steps_max,
                            int* n_steps, double* temp, double* pres, double* column, double* tvib,
                            double* dfrac, int* n_steps_needed) {
    if (!rays || !opt) return sr::fail(SR_ERR_ARG, "sr_los_steps_build_rays: bad argument");
    const int n_los = rays->n_los;
    const double* origin = rays->origin;
    const double* direction = rays->direction;
    const double delta_x_km = opt->delta_x_km, max_T_variation = opt->max_T_variation,
                 max_Plog_variation = opt->max_Plog_variation;
    if (opt->max_opt_depth > 0.0 && !opt->sigma_peak)
        return sr::fail(SR_ERR_ARG, "sr_los_steps_build_rays: max_opt_depth needs the per-gas peak "
                                    "cross-sections (sigma_peak)");
    const int nsz = atm ? std::max(atm->n_sza, 1) : 1;
    if (atm && atm->n_sza > 1) {
        if (!atm->sza_nodes) return sr::fail(SR_ERR_ARG, "sr_los_steps_build_rays: n_sza > 1 without sza_nodes");
        for (int i = 1; i < atm->n_sza; i++)
            if (!(atm->sza_nodes[i] > atm->sza_nodes[i - 1]))
                return sr::fail(SR_ERR_ARG, "sr_los_steps_build_rays: sza_nodes must ascend");
        if (!rays->sun && !rays->sza_fixed)
            return sr::fail(SR_ERR_ARG, "sr_los_steps_build_rays: SZA-dependent vibrational temperatures "
                                        "need the Sun direction or a fixed SZA per LOS");
    }
    if (!atm || n_los < 1 || !origin || !direction || !(delta_x_km > 0.0) || n_steps_max < 1 ||
        !n_steps || !temp || !pres || !column || atm->n_band < 1 || atm->n_z < 2 ||
        atm->n_gas < 1 || atm->n_gas > MAX_GAS_ST || !atm->z || !atm->temp || !atm->pres ||
        !atm->vmr || (atm->n_band > 1 && !atm->lat_edges) || n_par < 0 ||
        (n_par > 0 && (!masks || !dfrac)) || (atm->n_sets_max > 0 && (!tvib || !atm->tvib_on)))
        return sr::fail(SR_ERR_ARG, "sr_los_steps_build: bad argument");
    const int nb = atm->n_band, nz = atm->n_z, ng = atm->n_gas, nsx = atm->n_sets_max;
    for (int m = 0; m < ng * nsx; m++)
        if (atm->tvib_on[m] > 0 && !atm->tvib)
            return sr::fail(SR_ERR_ARG, "sr_los_steps_build: tvib_on set but no tvib table");
    cudaStream_t st = 0;
    std::vector<double> lnp((size_t)nb * nz);
    for (size_t i = 0; i < lnp.size(); i++) lnp[i] = std::log(atm->pres[i]);
    sr::PoolBuf<double> d_edges, d_z, d_temp, d_lnp, d_vmr, d_tvib, d_masks, d_org, d_dir, d_sun,
        d_fsza, d_nodes, ssza;
    sr::PoolBuf<int> d_on;
    if (nb > 1) SR_CUDA(d_edges.upload(atm->lat_edges, nb + 1, st));
    SR_CUDA(d_z.upload(atm->z, nz, st));
    SR_CUDA(d_temp.upload(atm->temp, (size_t)nb * nz, st));
    SR_CUDA(d_lnp.upload(lnp.data(), lnp.size(), st));
    SR_CUDA(d_vmr.upload(atm->vmr, (size_t)ng * nb * nz, st));
    if (atm->tvib) SR_CUDA(d_tvib.upload(atm->tvib, (size_t)ng * nsx * nb * nsz * nz, st));
    if (nsz > 1) SR_CUDA(d_nodes.upload(atm->sza_nodes, nsz, st));
    if (nsx > 0) SR_CUDA(d_on.upload(atm->tvib_on, (size_t)ng * nsx, st));
    if (n_par > 0) SR_CUDA(d_masks.upload(masks, (size_t)n_par * nb * nz, st));
    AtmDev A;
    A.n_band = nb; A.n_z = nz; A.n_gas = ng; A.n_sets_max = nsx; A.n_par = n_par; A.jac_gas = jac_gas;
    A.n_sza = nsz; A.sza_nodes = d_nodes.p;
    A.lat_edges = d_edges.p; A.z = d_z.p; A.temp = d_temp.p; A.lnpres = d_lnp.p; A.vmr = d_vmr.p;
    A.tvib = d_tvib.p; A.tvib_on = d_on.p; A.masks = d_masks.p;
    A.radius = atm->radius_km; A.top = atm->top_km;
    const double r_top = atm->radius_km + atm->top_km;
    const int n_pts_max = 2 * (int)std::floor(r_top / delta_x_km) + 3;
    // LOS blocks bound the per-point scratch (56 B per point) to ~1 GiB
    int blk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_los, ((size_t)1 << 30) / ((size_t)n_pts_max * 56)));
    if (const char* e = getenv("SR_STEPS_BLOCK")) blk = std::max(1, std::min(n_los, atoi(e)));   // test aid
    sr::PoolBuf<double> sT, sP, snd, sx, salt, o_temp, o_pres, o_col, o_tvib, o_dfrac;
    sr::PoolBuf<int> sband, sjz, d_npts, d_nsteps, d_bounds;
    const size_t np = (size_t)blk * n_pts_max;
    SR_CUDA(sT.alloc(np, st)); SR_CUDA(sP.alloc(np, st)); SR_CUDA(snd.alloc(np, st));
    SR_CUDA(sx.alloc(np, st)); SR_CUDA(salt.alloc(np, st)); SR_CUDA(sband.alloc(np, st));
    SR_CUDA(ssza.alloc(np, st));
    if (rays->sun) SR_CUDA(d_sun.alloc((size_t)3 * blk, st));
    if (rays->sza_fixed) SR_CUDA(d_fsza.alloc((size_t)blk, st));
    SR_CUDA(sjz.alloc(np, st));
    SR_CUDA(d_npts.alloc(blk, st)); SR_CUDA(d_nsteps.alloc(blk, st));
    SR_CUDA(d_bounds.alloc((size_t)blk * n_steps_max * 2, st));
    const size_t bls = (size_t)blk * n_steps_max;
    SR_CUDA(o_temp.alloc(bls, st)); SR_CUDA(o_pres.alloc(bls, st)); SR_CUDA(o_col.alloc(bls * ng, st));
    SR_CUDA(o_tvib.alloc(bls * ng * nsx, st));
    SR_CUDA(o_dfrac.alloc(bls * n_par, st));
    SR_CUDA(d_org.alloc((size_t)3 * blk, st));
    SR_CUDA(d_dir.alloc((size_t)3 * blk, st));
    int needed = 0;
    std::vector<double> tmp;
    for (int l0 = 0; l0 < n_los; l0 += blk) {
        const int nl = std::min(blk, n_los - l0);
        SR_CUDA(cudaMemcpyAsync(d_org.p, origin + (size_t)3 * l0, sizeof(double) * 3 * nl,
                                cudaMemcpyHostToDevice, st));
        SR_CUDA(cudaMemcpyAsync(d_dir.p, direction + (size_t)3 * l0, sizeof(double) * 3 * nl,
                                cudaMemcpyHostToDevice, st));
        if (rays->sun)
            SR_CUDA(cudaMemcpyAsync(d_sun.p, rays->sun + (size_t)3 * l0, sizeof(double) * 3 * nl,
                                    cudaMemcpyHostToDevice, st));
        if (rays->sza_fixed)
            SR_CUDA(cudaMemcpyAsync(d_fsza.p, rays->sza_fixed + l0, sizeof(double) * nl,
                                    cudaMemcpyHostToDevice, st));
        PtArgs pa;
        pa.A = A; pa.origin = d_org.p; pa.dir = d_dir.p;
        pa.sun = rays->sun ? d_sun.p : nullptr;
        pa.sza_fixed = rays->sza_fixed ? d_fsza.p : nullptr;
        pa.photon_order = opt->photon_order ? 1 : 0;
        pa.max_tau = opt->max_opt_depth;
        for (int m = 0; m < MAX_GAS_ST; m++)
            pa.sigma[m] = (opt->sigma_peak && m < ng) ? opt->sigma_peak[m] : 0.0;
        pa.sza = ssza.p;
        pa.n_los = nl; pa.n_pts_max = n_pts_max; pa.n_steps_max = n_steps_max;
        pa.delta_x = delta_x_km; pa.max_dT = max_T_variation; pa.max_dlnP = max_Plog_variation;
        pa.T = sT.p; pa.P = sP.p; pa.nd = snd.p; pa.x = sx.p; pa.alt = salt.p;
        pa.band = sband.p; pa.jz = sjz.p; pa.n_pts = d_npts.p; pa.n_steps = d_nsteps.p;
        pa.bounds = d_bounds.p;
        SR_LAUNCH(k_steps_points, (nl + 63) / 64, 64, 0, st, pa);
        IntArgs ia;
        ia.A = A; ia.n_los = nl; ia.n_pts_max = n_pts_max; ia.n_steps_max = n_steps_max;
        ia.T = sT.p; ia.P = sP.p; ia.nd = snd.p; ia.x = sx.p; ia.alt = salt.p; ia.sza = ssza.p;
        ia.band = sband.p; ia.jz = sjz.p; ia.n_steps = d_nsteps.p; ia.bounds = d_bounds.p;
        ia.temp = o_temp.p; ia.pres = o_pres.p; ia.column = o_col.p; ia.tvib = o_tvib.p;
        ia.dfrac = o_dfrac.p;
        SR_LAUNCH(k_steps_integrals, dim3((n_steps_max + INT_TX - 1) / INT_TX, nl), dim3(INT_TX, INT_TY), 0, st, ia);
        // copy the block into the caller's [..][n_los][n_steps_max] tables
        SR_CUDA(cudaMemcpyAsync(n_steps + l0, d_nsteps.p, sizeof(int) * nl, cudaMemcpyDeviceToHost, st));
        const size_t row = (size_t)nl * n_steps_max;
        SR_CUDA(cudaMemcpyAsync(temp + (size_t)l0 * n_steps_max, o_temp.p, 8 * row, cudaMemcpyDeviceToHost, st));
        SR_CUDA(cudaMemcpyAsync(pres + (size_t)l0 * n_steps_max, o_pres.p, 8 * row, cudaMemcpyDeviceToHost, st));
        for (int m = 0; m < ng; m++) {
            SR_CUDA(cudaMemcpyAsync(column + ((size_t)m * n_los + l0) * n_steps_max,
                                    o_col.p + (size_t)m * row, 8 * row, cudaMemcpyDeviceToHost, st));
            for (int s = 0; s < nsx; s++)
                SR_CUDA(cudaMemcpyAsync(tvib + (((size_t)m * nsx + s) * n_los + l0) * n_steps_max,
                                        o_tvib.p + ((size_t)m * nsx + s) * row, 8 * row,
                                        cudaMemcpyDeviceToHost, st));
        }
        if (n_par > 0)
            SR_CUDA(cudaMemcpyAsync(dfrac + (size_t)l0 * n_steps_max * n_par, o_dfrac.p,
                                    8 * row * n_par, cudaMemcpyDeviceToHost, st));
        SR_CUDA(cudaStreamSynchronize(st));
        for (int l = l0; l < l0 + nl; l++) needed = std::max(needed, n_steps[l]);
    }
    if (n_steps_needed) *n_steps_needed = needed;
    if (needed > n_steps_max)
        return sr::fail(SR_ERR_LIMIT, "sr_los_steps_build: a line of sight needs %d steps (n_steps_max = %d)",
                        needed, n_steps_max);
    return SR_OK;
}

}  // extern "C"
