/*
 * spectrobot.h -- C ABI of libspectrobot.so: the B200 (sm_100a) implementation of the
 * SpectRobot line-by-line hot path (Voigt/G-coefficient cross-sections, (P,T) LUT build,
 * line-of-sight radiative transfer).
 *
 * The reference reaches this path through three f2py extension modules (`lineshape`,
 * `fparts_mod`, `curgods`) plus Python loops around them; file:line citations below are
 * relative to the reference tree.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain C types only; every function returns an int status (0 = SR_OK) instead of the
 *     Fortran `stop` that kills the reference's worker process (lineshape.f:255,263);
 *     sr_last_error() gives the message for the calling thread.
 *   - "_host" / Tier-1 entry points take HOST pointers and are synchronous (they copy in,
 *     run the CUDA kernels, copy out).  Tier-2 entry points take DEVICE pointers plus a
 *     cudaStream_t (passed as void*), are asynchronous and never allocate output.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     SR_ERR_CUDA.  Pure table look-ups (sr_bd_tips_2003, sr_partition_sum) are host-only.
 */
#ifndef SPECTROBOT_H
#define SPECTROBOT_H

#ifdef __cplusplus
extern "C" {
#endif

#define SR_OK            0
#define SR_ERR_ARG       1   /* bad argument (incl. humliv_bb i1 > i2, lineshape.f:253) */
#define SR_ERR_DW        2   /* humliv_bb called with dw <= 0 (lineshape.f:260-264) */
#define SR_ERR_CUDA      3   /* CUDA runtime error / no device */
#define SR_ERR_GEOMETRY  4   /* a line's Voigt window has an unsupported region geometry */
#define SR_ERR_TABLE     5   /* no TIPS table for (mol, iso) (fparts_mod.f:294 falls through) */
#define SR_ERR_LUT       6   /* LUT interpolation: cell not found / extrapolating in P
                                (spect_main_module.py:989-991, 1058) */
#define SR_ERR_LIMIT     7   /* size limit exceeded (shared memory for n_sets, imxlines ...) */

#define SR_IMXSIG      13010     /* parameters.inc:65  - Voigt window length */
#define SR_IMXLINES    40000     /* parameters.inc:64 */
#define SR_IMXSIG_LONG 2000000   /* parameters.inc:64 */
#define SR_IMXSTP      8000      /* parameters.inc:64 */
#define SR_TIPS_N      119       /* fparts_mod.f:58 */

int         sr_version(void);
const char* sr_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long   sr_kernel_launch_count(void);
int         sr_device_count(void);
int         sr_set_device(int device);

/* ===========================================================================================
 * Tier 1 -- drop-ins for the f2py modules (HOST pointers, synchronous, run on the GPU)
 * =========================================================================================*/

/* lineshape.humliv_bb(x,i1,i2,x0,lw,dw) -> y        [lineshape.f:226-569; spect_classes.py:1999]
 * x, y: n doubles (the reference fixes n = 13010); i1,i2 1-based inclusive.  y is written on
 * [i1,i2] only, exactly where the Fortran writes.  dw is Doppler HWHM / sqrt(ln 2). */
int sr_humliv_bb(const double* x, int n, int i1, int i2, double x0, double lw, double dw,
                 double* y);

/* lineshape.sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe) -> spe_fin
 * [lineshape.f:2-25; spect_classes.py:1092].  matrix is Fortran-ordered with leading dimension
 * ld_lines (imxlines in the reference) and n_win columns (imxsig); init/fin 1-based inclusive;
 * spe_ini/spe_fin have n_spe elements (the reference pads to imxsig_long). */
int sr_sum_all_lines(const double* spe_ini, const double* matrix_colmajor, const int* init,
                     const int* fin, int n_lines, int ld_lines, int n_win, int n_spe,
                     double* spe_fin);

/* fparts_mod.bd_tips_2003(mol, iso) -> gi, t_grid[119], QT_grid[119]
 * [fparts_mod.f:33-295; spect_classes.py:1687].  Host table look-up. */
int sr_bd_tips_2003(int mol, int iso, double* gi, double* t119, double* q119);
/* CalcPartitionSum(mol, iso, temp) [spect_classes.py:1692-1710]: 4-point (3 below 85 K)
 * Lagrange interpolation of the table above.  Host. */
int sr_partition_sum(int mol, int iso, double temp, double* q);

/* curgods.curgod_fort_k(...) -> res   [curgods.f:2-98].  n_batch independent integrals, each
 * over n_p points; arrays are [n_batch][n_p] row-major; res[n_batch].  vmr / f may be NULL for
 * the variants that do not take them (k=1: nd,x; k=2: nd,vmr,x; k=3,4: nd,vmr,f,x). */
int sr_curgod(int k, const double* nd, const double* vmr, const double* f, const double* x,
              int n_p, int n_batch, double* res);
/* The same four integrals under the f2py names, one integral per call, n_p valid points:
 * curgod_fort_1(nd,x,n_p) [curgods.f:2-21], _2(nd,vmr,x,n_p) [:24-45], _3 / _4(nd,vmr,f,x,n_p)
 * [:48-73, :76-97] -> *res. */
int sr_curgod_1(const double* nd, const double* x, int n_p, double* res);
int sr_curgod_2(const double* nd, const double* vmr, const double* x, int n_p, double* res);
int sr_curgod_3(const double* nd, const double* vmr, const double* f, const double* x, int n_p,
                double* res);
int sr_curgod_4(const double* nd, const double* vmr, const double* f, const double* x, int n_p,
                double* res);

/* ===========================================================================================
 * Tier 2 -- fused cross-section path (K1/K2): calc_shapes_lines + Calc_Gcoeffs + BuildCoeff +
 * add_lines_to_spectrum + sum_all_lines for whole (P,T) cells
 * [spect_classes.py:1378-1462, 312-343, 1277-1337, 1016-1147; spect_main_module.py:718-788,
 *  1122-1168]
 * =========================================================================================*/

/* Physical constants as the host reads them from scipy (spect_classes.py:44-47, 1984) and the
 * libm constants the Python side forms with math.log/sqrt (spect_classes.py:1984,1997,1999). */
typedef struct {
    double h_cgs, c_cgs, k_cgs, avogadro;
    double ln2;          /* math.log(2.0) */
    double sqrt_ln2;     /* math.sqrt(math.log(2.0)) */
    double sqrt_pi_ln2;  /* math.sqrt(np.pi/math.log(2.0)) */
} sr_consts;
void sr_default_consts(sr_consts* c); /* CODATA-2018 exact SI values */

/* Line table, HOST arrays of n_lines entries (HITRAN fields of SpectLine, spect_classes.py:50-53,
 * with the level link of LinkToMolec :122-150 resolved to integers by the caller):
 * up_set / lo_set = index of the LUT set (vibrational level) the line feeds as upper level
 * (sp_emission, ind_emission) and as lower level (absorption); a negative value on either
 * drops the line (spect_classes.py:1384-1388).  LTE isotopologue: n_sets = 1, both 0, e_vib 0. */
typedef struct {
    int n_lines;
    const double *freq, *a_coeff, *air_broad, *t_dep, *e_lower, *g_up, *g_lo;
    const double *e_vib_up, *e_vib_lo;
    const int *up_set, *lo_set;
} sr_lines;

typedef struct sr_lineset sr_lineset; /* device-resident line table bound to one spectral grid */

/* grid: the spectral grid EXACTLY as prepare_spe_grid builds it (np.arange), n_grid
 * points; lin_grid: the SR_IMXSIG window offsets of PrepareCalcShapes (spect_classes.py:1446);
 * mm: molar mass of the isotopologue (isomolec.MM).  Uploads everything, computes closest_grid
 * (spect_classes.py:1937) per line on the device and groups lines by (upper, lower) set. */
int sr_lineset_create(const sr_lines* lines, const double* grid, long n_grid,
                      const double* lin_grid, int n_sets, double mm, const sr_consts* consts,
                      sr_lineset** out);
int sr_lineset_destroy(sr_lineset* ls);
long sr_lineset_n_active(const sr_lineset* ls);   /* lines kept after the level-link filter */
int sr_lineset_centres(const sr_lineset* ls, int* ind_host); /* closest_grid index per INPUT line
                                                                (-1 for dropped lines) */

/* G-coefficient spectra of n_cells (P,T) cells.  pt_host: [n_cells][2] = (P_hPa, T_K) on the
 * host.  (Numerics: every point of regions 2/3/4 and of the near far wings is evaluated with the
 * formulas of lineshape.f:443-562; in launches of two waves or more, region-1 wings that cover a
 * whole 512-point tile from more than two tile lengths away are evaluated at 12 Chebyshev nodes
 * per tile and interpolated, <= 4e-11 of the line's own value - DESIGN.md 4, K1; the environment
 * variable SR_K1_FAR=0 evaluates every point.)  out: [n_cells][n_sets][3][n_grid] doubles, ctype order sp_emission, ind_emission,
 * absorption; every element is written (no zero-fill needed).
 * _dev: out is a DEVICE pointer, asynchronous on `stream`.  _host: out is a HOST pointer. */
int sr_gcoeff_cells_dev(sr_lineset* ls, const double* pt_host, int n_cells, double* out_dev,
                        void* stream);
int sr_gcoeff_cells_host(sr_lineset* ls, const double* pt_host, int n_cells, double* out_host);
/* Synchronise `stream` and report the humliv_bb STOP conditions / geometry errors that the
 * asynchronous _dev calls on this lineset raised since the last check (SR_OK if none). */
int sr_lineset_check(sr_lineset* ls, void* stream);
/* Same, stored as float32 (the reference's split_and_compress_LUTS casts the LUT to float32,
 * spect_main_module.py:1676 via spect_classes.py:732). out32_dev: [n_cells][n_sets][3][n_grid]. */
int sr_gcoeff_cells_dev_f32(sr_lineset* ls, const double* pt_host, int n_cells, float* out32_dev,
                            double* scratch_dev, void* stream);
/* Same with padded rows: out32_dev is [n_cells][n_sets][3][row_stride] floats, row_stride >=
 * n_grid; only the first n_grid elements of a row are written.  A row stride that is a multiple
 * of 32 floats keeps every LUT row 128-byte aligned, which lets the LOS kernels read the table
 * with 16-byte vector loads. */
int sr_gcoeff_cells_dev_f32_ld(sr_lineset* ls, const double* pt_host, int n_cells,
                               float* out32_dev, long row_stride, void* stream);

/* The same cells restricted to the grid points [pt0, pt0+n_pts) of the lineset's grid (a
 * wavenumber slab: the multi-GPU LUT build gives every rank all cells of one slab).
 * The window positions of the lines are still those of closest_grid over the WHOLE grid
 * (spect_classes.py:1937-1943), so the slab is bit-identical to the same points of a full build;
 * the lineset only needs the lines whose 13010-point window reaches the slab.  out_dev:
 * [n_cells][n_sets][3][row_stride] doubles (f32 == 0) or floats (f32 != 0), row_stride >= n_pts
 * (0 = n_pts); pt0 must be a multiple of sr_lineset_tile_points(). */
int sr_gcoeff_cells_window_dev(sr_lineset* ls, const double* pt_host, int n_cells, void* out_dev,
                               int f32, long row_stride, long pt0, long n_pts, void* stream);
int sr_lineset_tile_points(const sr_lineset* ls);

/* Per-line normalised shapes of one cell (MakeShapeLine with keep_memory, spect_classes.py:174):
 * shapes_dev [n_active][SR_IMXSIG] in the lineset's internal (sorted) line order, and the three
 * G coefficients g_dev [n_active][3]; order_host[n_active] maps sorted position -> input line. */
int sr_line_shapes_dev(sr_lineset* ls, double pres_hpa, double temp, double* shapes_dev,
                       double* g_dev, void* stream);
int sr_lineset_order(const sr_lineset* ls, int* order_host);

/* ===========================================================================================
 * Tier 2 -- line-of-sight radiative transfer (K3 / K3a)
 * [callers spect_main_module.py:2506-2545, 3133-3221; coefficient assembly :2134-2299;
 *  LUT interpolation :997-1066.  The integral itself lives in the reference's missing
 *  spect_base_module; DESIGN.md section 6 is the specification implemented here.]
 * =========================================================================================*/

/* K3: recursion over materialised layers.  tau, src: [n_los][n_steps_max][n_pts] doubles
 * (device), n_steps[n_los] (device ints), steps ordered from the far end to the observer.
 *   I <- I*exp(-tau) + src*(1-exp(-tau)),  I_0 = i0[los][pt] or 0 when i0 == NULL
 * solo_absorption != 0 drops the emission term.  rad: [n_los][n_pts]. */
int sr_los_rt_layers_dev(const double* tau, const double* src, const int* n_steps, int n_los,
                         int n_steps_max, long n_pts, const double* i0, int solo_absorption,
                         double* rad, void* stream);

/* A float32 LUT of one isotopologue resident on the device (the compressed LUT of
 * split_and_compress_LUTS).  g32_dev: [n_cells][n_sets][3][n_grid] floats (device, caller-owned,
 * must outlive the handle); pt_host [n_cells][2]; level_energy_host[n_sets] (cm-1; ignored when
 * lte_unidentified); mol/iso select the TIPS table for Q(T); iso_ratio multiplies the columns. */
typedef struct sr_lut sr_lut;
int sr_lut_create(const float* g32_dev, const double* pt_host, int n_cells, int n_sets,
                  long n_grid, const double* level_energy_host, int mol, int iso,
                  double iso_ratio, int lte_unidentified, const sr_consts* consts, sr_lut** out);
/* Same for a table with padded rows, [n_cells][n_sets][3][row_stride] (see
 * sr_gcoeff_cells_dev_f32_ld); the padding elements must be finite. */
int sr_lut_create_ld(const float* g32_dev, long row_stride, const double* pt_host, int n_cells,
                     int n_sets, long n_grid, const double* level_energy_host, int mol, int iso,
                     double iso_ratio, int lte_unidentified, const sr_consts* consts,
                     sr_lut** out);
int sr_lut_destroy(sr_lut* lut);
/* Bit s of mask: the spontaneous emission of set (level) s contributes to the source function of
 * the LOS calls that follow (default: all bits set).  Absorption and induced emission are not
 * affected, so a run with one emitter switched on gives the radiance the observer receives from
 * that emitter through the whole mixture - the per-gas / per-level `single_rads` of radtrans
 * (spect_main_module.py:3176-3186, 3242-3246; track_levels :2250-2290).  <= 64 sets. */
int sr_lut_set_emission_mask(sr_lut* lut, unsigned long long mask);

/* Step tables of a LOS batch (HOST arrays, steps ordered far end -> observer):
 *   n_steps[n_los]; temp, pres [n_los][n_steps_max] (Curtis-Godson T in K, P in hPa);
 *   column [n_gas][n_los][n_steps_max] (molecules cm-2 of the gas, before the isotopic ratio);
 *   tvib   [n_gas][n_sets_max][n_los][n_steps_max] vibrational temperatures (K); NULL = LTE
 *          (T_vib = T, spect_main_module.py:2231-2234). */
typedef struct {
    int n_los, n_steps_max, n_gas, n_sets_max;
    const int* n_steps;
    const double *temp, *pres, *column, *tvib;
} sr_los_steps;

/* Fused K3a+K3: radiances of the batch on grid points [pt0, pt0+n_pts) from the LUTs of n_gas
 * isotopologues.  rad_dev: [n_los][n_pts] doubles (device).  i0_dev as above (may be NULL). */
int sr_los_rt_lut_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                      const double* i0_dev, int solo_absorption, double* rad_dev, void* stream);
int sr_los_rt_lut_host(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                       const double* i0_host, int solo_absorption, double* rad_host);
/* Same, reduced to the instrument's low-resolution channels on the device (radtrans ->
 * hires_to_lowres, spect_main_module.py:3342-3374, spect_classes.py:1180-1191): the batch is
 * processed in LOS blocks, each block's hi-res radiances are convolved (see
 * sr_convolve_lowres_dev) and only low_dev [n_los][n_chan] is kept, so a 10^4-pixel batch never
 * materialises its 288 GB of hi-res output.  grid_dev: the n_pts spectral grid points of the
 * window [pt0, pt0+n_pts); centre_dev / width_dev: [n_chan] channel centres and Gaussian widths
 * in the units of the grid (device).  i0_dev: [n_los][n_pts] or NULL. */
int sr_los_rt_lut_lowres_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                             const double* grid_dev, const double* centre_dev,
                             const double* width_dev, int n_chan, double n_sigma,
                             const double* i0_dev, int solo_absorption, double* low_dev,
                             void* stream);
/* Synchronise `stream` and report LUT-interpolation errors (SR_ERR_LUT) raised by the
 * asynchronous LOS calls that used luts[0] since the last check. */
int sr_los_check(sr_lut* const* luts, void* stream);
/* K3a alone: materialise tau/src [n_los][n_steps_max][n_pts] (device) for sr_los_rt_layers_dev. */
int sr_los_tau_src_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                       double* tau_dev, double* src_dev, void* stream);

/* -------------------------------------------------------------------------------------------
 * Analytic Jacobians: d radiance / d retrieval parameter, the `hires_deriv`
 * spectra the reference's retrieval collects per LOS and parameter
 * [callers spect_main_module.py:2758, 2837, 2867-2881; RetParam / maskgrid :600-656; the
 *  computation itself is in the missing spect_base_module.radtran_fast - DESIGN.md 6.5 is the
 *  specification implemented here].
 * A parameter p is a VMR node of ONE gas: it changes step k only through that gas' column u_k,
 *   dfrac[los][step][p] = (d u_k / d p) / u_k      (0 where the node's mask does not reach)
 * with the step's Curtis-Godson T and P held fixed.
 * ----------------------------------------------------------------------------------------- */

/* K3 + Jacobians over materialised layers, all DEVICE pointers.  tau, emi: [n_los][n_steps_max]
 * [n_pts] optical depth and emission (coefficient x column, the outputs of sr_los_abs_emi_dev)
 * of the whole mixture; tau_g, emi_g: the same for the retrieved gas alone, or both NULL when
 * that gas is the only absorber; dfrac [n_los][n_steps_max][n_par]; rad [n_los][n_pts];
 * jac [n_los][n_par][n_pts].  n_los <= 65535. */
int sr_los_rt_layers_jac_dev(const double* tau, const double* emi, const double* tau_g,
                             const double* emi_g, const double* dfrac, int n_par,
                             const int* n_steps, int n_los, int n_steps_max, long n_pts,
                             const double* i0, int solo_absorption, double* rad, double* jac,
                             void* stream);
/* Fused from the LUTs.  gas_in_jac[n_gas] (HOST ints, NULL = all): which LUTs (isotopologues)
 * belong to the retrieved gas; dfrac_host [n_los][n_steps_max][n_par] (HOST).
 * rad_dev [n_los][n_pts], jac_dev [n_los][n_par][n_pts] (device). */
int sr_los_rt_lut_jac_dev(sr_lut* const* luts, const sr_los_steps* steps, int n_par,
                          const int* gas_in_jac, const double* dfrac_host, long pt0, long n_pts,
                          const double* i0_dev, int solo_absorption, double* rad_dev,
                          double* jac_dev, void* stream);
/* Same, radiances and derivatives reduced to the instrument channels on the device
 * (par_mod.hires_deriv.hires_to_lowres, spect_main_module.py:2874): low_dev [n_los][n_chan],
 * jac_low_dev [n_los][n_par][n_chan]; other arguments as sr_los_rt_lut_lowres_dev. */
int sr_los_rt_lut_jac_lowres_dev(sr_lut* const* luts, const sr_los_steps* steps, int n_par,
                                 const int* gas_in_jac, const double* dfrac_host, long pt0,
                                 long n_pts, const double* grid_dev, const double* centre_dev,
                                 const double* width_dev, int n_chan, double n_sigma,
                                 const double* i0_dev, int solo_absorption, double* low_dev,
                                 double* jac_low_dev, void* stream);

/* make_abscoeff_LUTS_fast [spect_main_module.py:2134-2299]: per (LOS, step, point)
 * abs = sum_gas column*ratio*sum_lev (G_abs - G_ind)*pop and emi = sum_gas column*ratio*sum_lev
 * G_sp*pop, [n_los][n_steps_max][n_pts] each (device).  With column = 1/ratio these are the
 * reference's absorption / emission coefficients of the isotopologue. */
int sr_los_abs_emi_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                       double* abs_dev, double* emi_dev, void* stream);

/* LutSet.calculate(P,T) [spect_main_module.py:997-1066]: the (up to) 4 cells and weights the
 * reference's nearest-node bilinear rule picks.  Host helper (index logic only). */
int sr_lut_weights(const double* pt_host, int n_cells, double pres, double temp, int cell[4],
                   double w[4]);

/* ===========================================================================================
 * Tier 2 -- instrument convolution of hi-res spectra to low-res channels
 * SpectralIntensity.hires_to_lowres -> convolve_to_grid_from_irregular
 * [spect_classes.py:1180-1191, 883-918, gaussian :1926, conv_single :1162]
 * =========================================================================================*/

/* Low-resolution channels of an instrument.  units says how the channel centres / widths relate to
 * the hi-res grid: SR_CHAN_SAME_UNITS - same units as the grid; SR_CHAN_NM_FROM_CM1 - grid in cm-1,
 * channels in nm: the reference first converts the hi-res spectrum to the wavelength axis
 * (convert_grid_to, spect_classes.py:771-778: grid -> 1e7/grid, spectrum -> spectrum*grid^2*1e-7,
 * reversed) and convolves there (hires_to_lowres :1180-1191); the kernels do the same per point. */
#define SR_CHAN_SAME_UNITS  0
#define SR_CHAN_NM_FROM_CM1 1
typedef struct {
    int n_chan;
    const double *centre_dev, *width_dev;   /* [n_chan], DEVICE pointers */
    double n_sigma;                         /* window half-width in widths (reference: 5) */
    int units;
} sr_channels;

/* spec: [n_spec][n_pts] spectra on the common ascending (possibly irregular) grid[n_pts];
 * channel c integrates spec * N(centre[c], width[c]) with the trapezoid rule over the grid points
 * within +-n_sigma*width[c] (reference default n_sigma = 5); out: [n_spec][n_chan].
 * _dev: every pointer is a DEVICE pointer, asynchronous on `stream`.  _host: HOST pointers. */
int sr_convolve_lowres_dev(const double* grid_dev, long n_pts, const double* spec_dev, int n_spec,
                           const double* centre_dev, const double* width_dev, int n_chan,
                           double n_sigma, double* out_dev, void* stream);
int sr_convolve_lowres_host(const double* grid, long n_pts, const double* spec, int n_spec,
                            const double* centre, const double* width, int n_chan,
                            double n_sigma, double* out);
/* Same with a channel set that carries its units (see sr_channels). */
int sr_convolve_channels_dev(const double* grid_dev, long n_pts, const double* spec_dev, int n_spec,
                             const sr_channels* ch, double* out_dev, void* stream);
int sr_convolve_channels_host(const double* grid, long n_pts, const double* spec, int n_spec,
                              const double* centre, const double* width, int n_chan,
                              double n_sigma, int units, double* out);

/* sr_los_rt_lut_lowres_dev / sr_los_rt_lut_jac_lowres_dev with a channel set that carries its units
 * (an observation in nm over a cm-1 hi-res grid, as the reference's VIMS pixels are,
 * spect_main_module.py:2372): the radiances (and derivative spectra) of every LOS block are
 * converted per point and convolved on the wavelength axis on the device. */
int sr_los_rt_lut_channels_dev(sr_lut* const* luts, const sr_los_steps* steps, long pt0, long n_pts,
                               const double* grid_dev, const sr_channels* ch, const double* i0_dev,
                               int solo_absorption, double* low_dev, void* stream);
int sr_los_rt_lut_jac_channels_dev(sr_lut* const* luts, const sr_los_steps* steps, int n_par,
                                   const int* gas_in_jac, const double* dfrac_host, long pt0,
                                   long n_pts, const double* grid_dev, const sr_channels* ch,
                                   const double* i0_dev, int solo_absorption, double* low_dev,
                                   double* jac_low_dev, void* stream);

/* ===========================================================================================
 * Tier 2 -- LOS geometry and radtran steps for a whole batch
 * LineOfSight.calc_atm_intersections + calc_radtran_steps [callers spect_main_module.py:2746-2767,
 * 3133-3147; options radtran_3D_ch4.py:200-202, 311; integrals curgods.f:2-98; the methods
 * themselves are in the reference's missing spect_base_module - DESIGN.md 6.1 is the specification]
 * =========================================================================================*/

/* Atmosphere on an altitude grid (km, ascending) with n_band latitude boxes (lat_edges[n_band+1]
 * in degrees, ascending; NULL when n_band == 1).  temp linear, pres log-linear, vmr and tvib
 * linear in altitude.  One entry of vmr per LUT of the later LOS call (isotopologues of one gas
 * repeat the profile).  tvib_on[n_gas][n_sets_max]: 1 = the level has its own T_vib profile,
 * 0 = T_vib is the step temperature, -1 = the gas has no such level (padding, 100 K). */
typedef struct {
    int n_band, n_z, n_gas, n_sets_max;
    const double *lat_edges, *z;
    const double *temp, *pres;     /* [n_band][n_z], K and hPa */
    const double *vmr;             /* [n_gas][n_band][n_z] */
    const double *tvib;            /* [n_gas][n_sets_max][n_band][n_sza][n_z] or NULL */
    const int* tvib_on;            /* [n_gas][n_sets_max] (NULL when n_sets_max == 0) */
    double radius_km, top_km;      /* planet radius, atmosphere extension above it */
    /* Solar-zenith-angle axis of the vibrational temperatures (the reference's 3-D T_vib profiles
     * are functions of (lat, SZA, alt): radtran_3D_ch4.py:249-250, add_nLTE_molecs_from_tvibmanuel_3D):
     * n_sza nodes in degrees, ascending; linear between nodes, clamped outside.  n_sza <= 1: no SZA
     * dependence (sza_nodes may be NULL). */
    int n_sza;
    const double* sza_nodes;
} sr_atmosphere;

/* Rays of a LOS batch.  sun: unit vector from the planet centre towards the Sun per LOS (the
 * pixel's sub-solar point, spect_main_module.py:3096): the SZA of every sample point is the angle
 * between its position vector and sun (LineOfSight.calc_SZA_along_los, :3141).  sza_fixed: one SZA
 * (degrees) used at every point of the LOS instead (use_tangent_sza, :3138-3139).  Both may be
 * NULL when the atmosphere has no SZA axis. */
typedef struct {
    int n_los;
    const double *origin, *direction;   /* [n_los][3] km, unit vectors */
    const double *sun;                  /* [n_los][3] or NULL */
    const double *sza_fixed;            /* [n_los] or NULL */
} sr_los_rays;

/* Options of calc_atm_intersections / calc_radtran_steps (radtran_3D_ch4.py:200-202, 311;
 * spect_radtran_test.py:175).  max_opt_depth > 0 adds a third merge limit: a step is closed when
 * sum_gas sigma_peak[gas] * column_gas over it exceeds max_opt_depth (sigma_peak: the caller's
 * estimate of the largest absorption cross-section of each gas, cm2 per molecule).  photon_order:
 * LOS_order='photon' (invert_LOS_direction, spect_main_module.py:3135-3137): samples and steps run
 * from the observer's side of the atmosphere to the far end, i.e. the layer recursion treats the
 * observer-side top of the atmosphere as the photons' entry point. */
typedef struct {
    double delta_x_km, max_T_variation, max_Plog_variation, max_opt_depth;
    const double* sigma_peak;           /* [n_gas] or NULL */
    int photon_order;
} sr_steps_opt;

/* For n_los rays (origin = observer position, direction = unit vector, planetocentric Cartesian
 * km): samples every delta_x_km anchored on the tangent point between the atmosphere
 * intersections (or down to the surface), ordered far end -> observer; consecutive samples are
 * merged into steps while max T - min T <= max_T_variation and max ln P - min ln P <=
 * max_Plog_variation; every step gets air-weighted Curtis-Godson T and P, the gas columns
 * (molecules cm-2) and column-weighted vibrational temperatures.  Outputs are HOST arrays in the
 * sr_los_steps layout with n_steps_max columns (padding: 100 K, 1e-6 hPa, column 0).
 * n_par > 0: masks[n_par][n_band][n_z] are the weights (latitude box x altitude, linear in
 * altitude) of retrieval parameters of gas entry jac_gas; dfrac[n_los][n_steps_max][n_par] = (d column / d parameter) / column for
 * sr_los_rt_lut_jac_*.  n_steps_needed (optional) returns the largest step count; when it exceeds
 * n_steps_max the call returns SR_ERR_LIMIT. */
int sr_los_steps_build(const sr_atmosphere* atm, int n_los, const double* origin,
                       const double* direction, double delta_x_km, double max_T_variation,
                       double max_Plog_variation, int n_par, const double* masks, int jac_gas,
                       int n_steps_max, int* n_steps, double* temp, double* pres, double* column,
                       double* tvib, double* dfrac, int* n_steps_needed);
/* Same with the per-LOS Sun geometry and the full option set. */
int sr_los_steps_build_rays(const sr_atmosphere* atm, const sr_los_rays* rays, const sr_steps_opt* opt,
                            int n_par, const double* masks, int jac_gas, int n_steps_max,
                            int* n_steps, double* temp, double* pres, double* column, double* tvib,
                            double* dfrac, int* n_steps_needed);

/* Per-kernel device timing for bench.py's live roofline: while enabled, the library brackets the
 * launches of the kernels below with CUDA events on the launching stream.  sr_prof_enable clears the
 * records; sr_prof_summary synchronises the device and returns, for one kernel kind, the number of
 * launches, their summed duration (ms) and their summed algorithmic work (FP64 flop for
 * SR_PROF_LOS_MMA and SR_PROF_LOS_FUSED, bytes for SR_PROF_LOS_LAYERS and SR_PROF_CONV,
 * line*gridpoint evaluations for SR_PROF_VOIGT_TILE, centre points for SR_PROF_VOIGT_CORE). */
#define SR_PROF_LOS_MMA     0
#define SR_PROF_LOS_LAYERS  1
#define SR_PROF_CONV        2
#define SR_PROF_VOIGT_TILE  3
#define SR_PROF_VOIGT_CORE  4
#define SR_PROF_LOS_FUSED   5
int sr_prof_enable(int on);
int sr_prof_summary(int kind, long long* launches, double* ms, double* work);

/* FP64 FMA micro-benchmark used by bench.py for the K1/K2 roofline denominator: runs
 * `iters` dependent-chain FMAs per thread on a full grid and returns achieved FLOP/s. */
int sr_fp64_peak(int iters, double* flops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SPECTROBOT_H */
