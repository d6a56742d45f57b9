"""ctypes front-end of oracle/libsr_oracle.so + NumPy restatements of the reference's Python glue.

TEST INFRASTRUCTURE ONLY (see the header of sr_oracle.c): imported by tests/, by
__graft_entry__.smoke() and by the cpu_baseline / --impl reference legs of bench.py.  The product
package spectrobot_b200 never imports this module.

Every function cites the reference file:line it restates (paths relative to /root/reference).
"""
import ctypes as C
import math as mt
import os
import subprocess

import numpy as np
import scipy.constants as const
from scipy.interpolate import lagrange

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

IMXSIG = 13010          # spect_classes.py:27
IMXLINES = 40000        # spect_classes.py:28
IMXSIG_LONG = 2000000   # spect_classes.py:29

# spect_classes.py:44-47 -- constants come from the installed scipy
h_cgs = const.physical_constants['Planck constant'][0] * 1.e7
c_cgs = const.c * 1.e2
k_cgs = const.physical_constants['Boltzmann constant'][0] * 1.e7
c2 = h_cgs * c_cgs / k_cgs
T_ref = 296.0
hpa_to_atm = 0.00098692326671601
CONSTS = np.array([h_cgs, c_cgs, k_cgs, const.Avogadro], dtype=np.float64)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build():
    """Compile libsr_oracle.so with the Makefile beside this file (no-op when up to date)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libsr_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_humliv_bb.restype = C.c_int
        L.orc_humliv_bb.argtypes = [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_double, _dp]
        L.orc_humliv_regions.restype = C.c_int
        L.orc_humliv_regions.argtypes = [_dp, C.c_int, C.c_int, C.c_double, C.c_double,
                                         C.c_double, _ip]
        L.orc_humli_bb.restype = C.c_double
        L.orc_humli_bb.argtypes = [C.c_double, C.c_double]
        L.orc_sum_all_lines.restype = None
        L.orc_sum_all_lines.argtypes = [_dp, _dp, _ip, _ip, C.c_int, C.c_int, C.c_int, _dp]
        for k, na in ((1, 2), (2, 3), (3, 4), (4, 4)):
            f = getattr(L, "orc_curgod_%d" % k)
            f.restype = C.c_double
            f.argtypes = [_dp] * na + [C.c_int]
        L.orc_bd_tips_2003.restype = C.c_int
        L.orc_bd_tips_2003.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp]
        L.orc_partition_sum.restype = C.c_double
        L.orc_partition_sum.argtypes = [C.c_int, C.c_int, C.c_double]
        L.orc_widths.restype = None
        L.orc_widths.argtypes = [C.c_double] * 6 + [_dp, _dp, _dp]
        L.orc_gcoeffs.restype = None
        L.orc_gcoeffs.argtypes = [C.c_double] * 8 + [_dp, _dp]
        L.orc_closest_grid.restype = C.c_long
        L.orc_closest_grid.argtypes = [_dp, C.c_long, C.c_double]
        L.orc_line_shape.restype = C.c_int
        L.orc_line_shape.argtypes = [C.c_double] * 4 + [_dp, _dp, _dp]
        L.orc_gcoeff_cell.restype = C.c_int
        L.orc_gcoeff_cell.argtypes = ([C.c_int] + [_dp] * 9 + [_ip, _ip, _dp, C.c_long, _dp,
                                      C.c_double, C.c_double, C.c_double, _dp, C.c_int, _dp,
                                      C.c_int])
        if hasattr(L, "orc_los_rt"):
            L.orc_los_rt.restype = C.c_int
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_ip)


# ---------------------------------------------------------------------------------------------
# f2py-shaped entry points (lineshape.f / fparts_mod.f / curgods.f)
# ---------------------------------------------------------------------------------------------
def humliv_bb(x, i1, i2, x0, lw, dw):
    """lineshape.f:226-569.  Returns a new array; untouched entries are 0."""
    x, px = _d(x)
    y = np.zeros_like(x)
    rc = lib().orc_humliv_bb(px, len(x), int(i1), int(i2), float(x0), float(lw), float(dw),
                             y.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("humliv_bb: Fortran STOP condition %d" % rc)
    return y


def humliv_regions(x, i1, i2, x0, lw, dw):
    x, px = _d(x)
    out = np.zeros(4, dtype=np.int32)
    rc = lib().orc_humliv_regions(px, int(i1), int(i2), float(x0), float(lw), float(dw),
                                  out.ctypes.data_as(_ip))
    if rc:
        raise RuntimeError("humliv_regions: %d" % rc)
    return out


def humli_bb(rx, ry):
    """lineshape.f:150-205 (scalar, D0 coefficients)."""
    return lib().orc_humli_bb(float(rx), float(ry))


def sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe=None):
    """lineshape.f:2-25; matrix is (ld, n_win) Fortran-ordered, init/fin 1-based inclusive."""
    spe_ini, ps = _d(spe_ini)
    m = np.asfortranarray(matrix, dtype=np.float64)
    init, pi_ = _i(init)
    fin, pf = _i(fin)
    out = np.empty_like(spe_ini)
    lib().orc_sum_all_lines(ps, m.ctypes.data_as(_dp), pi_, pf, int(n_lines), m.shape[0],
                            len(spe_ini), out.ctypes.data_as(_dp))
    return out


def curgod(k, *arrs):
    """curgods.f:2-98, k in 1..4; arrays as in the Fortran signature, n_p = len(x)."""
    keep = [_d(a) for a in arrs]
    n_p = len(keep[-1][0])
    return getattr(lib(), "orc_curgod_%d" % k)(*[p for _, p in keep], n_p)


def bd_tips_2003(mol, iso):
    """fparts_mod.f:33-295 -> (gi, t_grid[119], QT_grid[119])."""
    gi = C.c_double()
    t = np.empty(119)
    q = np.empty(119)
    rc = lib().orc_bd_tips_2003(int(mol), int(iso), C.byref(gi), t.ctypes.data_as(_dp),
                                q.ctypes.data_as(_dp))
    if rc:
        raise KeyError("bd_tips_2003: no table for mol %d iso %d" % (mol, iso))
    return gi.value, t, q


def CalcPartitionSum(mol, iso, temp=296.0):
    """spect_classes.py:1692-1710, literally (scipy lagrange through <=2 + <=2 nodes)."""
    gi, T_grid, Q_grid = bd_tips_2003(mol, iso)
    x1 = T_grid[T_grid <= temp][-2:]
    x2 = T_grid[T_grid > temp][:2]
    x = np.hstack([x1, x2])
    qg1 = Q_grid[T_grid <= temp][-2:]
    qg2 = Q_grid[T_grid > temp][:2]
    qg = np.hstack([qg1, qg2])
    poli = lagrange(x, qg)
    return poli(temp)


def partition_sum_c(mol, iso, temp):
    return lib().orc_partition_sum(int(mol), int(iso), float(temp))


# ---------------------------------------------------------------------------------------------
# spect_classes.py per-line physics, literal NumPy/math restatements
# ---------------------------------------------------------------------------------------------
def Boltz_ratio_nodeg(wavenumber, temp):
    """spect_classes.py:1876-1878"""
    return np.exp(-c2 * wavenumber / temp)


def Lorenz_width(Temp, Pres_atm, T_dep_broad, Air_broad, Self_broad=0.0, Self_pres_atm=0.0):
    """spect_classes.py:1967-1974"""
    return (T_ref / Temp) ** T_dep_broad * (Air_broad * (Pres_atm - Self_pres_atm)
                                            + Self_broad * Self_pres_atm)


def Doppler_width(Temp, MM, wn_0):
    """spect_classes.py:1976-1986"""
    return wn_0 / c_cgs * mt.sqrt(2 * const.Avogadro * k_cgs * Temp * mt.log(2.0) / MM)


def Einstein_A_to_B(A_coeff, wavenumber):
    """spect_classes.py:1736-1754, units 'cm3ergcm2'"""
    fact_2 = 2 * h_cgs * c_cgs ** 2 * wavenumber ** 3
    return A_coeff / fact_2


def Calc_Gcoeffs(Freq, A_coeff, E_lower, g_up, g_lo, E_vib_up, E_vib_lo, Temp):
    """spect_classes.py:312-343 + 1806-1853 -> (sp_emission, ind_emission, absorption)."""
    if not (A_coeff != 0.0 and g_lo != 0.0 and g_up != 0.0):
        return 0., 0., 0.
    B_21 = Einstein_A_to_B(A_coeff, Freq)
    rot_pop = g_up * Boltz_ratio_nodeg(E_lower + Freq - E_vib_up, Temp)
    G_sp = h_cgs * c_cgs * Freq * rot_pop * A_coeff / (4 * np.pi)
    G_in = h_cgs * c_cgs * Freq * rot_pop * B_21 / (4 * np.pi)
    B_12 = B_21 * g_up / g_lo
    rot_pop = g_lo * Boltz_ratio_nodeg(E_lower - E_vib_lo, Temp)
    G_ab = h_cgs * c_cgs * Freq * rot_pop * B_12 / (4 * np.pi)
    return G_sp, G_in, G_ab


def widths_c(freq, air, tdep, T, P, MM):
    lw = C.c_double()
    dw = C.c_double()
    _, pc = _d(CONSTS)
    lib().orc_widths(freq, air, tdep, T, P, MM, pc, C.byref(lw), C.byref(dw))
    return lw.value, dw.value


def gcoeffs_c(freq, A, El, gu, gl, evu, evl, T):
    g = np.zeros(3)
    _, pc = _d(CONSTS)
    lib().orc_gcoeffs(freq, A, El, gu, gl, evu, evl, T, pc, g.ctypes.data_as(_dp))
    return g


def prepare_spe_grid(wn_range, sp_step=5.e-4):
    """spect_main_module.py:1262-1272 (grid only)"""
    return np.arange(wn_range[0], wn_range[1] + sp_step / 2, sp_step, dtype=float)


def line_window_offsets(grid):
    """spect_classes.py:1445-1446: lin_grid built from wn_arr.step() = grid[1]-grid[0]."""
    sp_step = grid[1] - grid[0]
    lin = np.arange(-IMXSIG * sp_step / 2, IMXSIG * sp_step / 2, sp_step, dtype=float)
    if len(lin) != IMXSIG:
        raise ValueError("window has %d points, humliv_bb needs exactly %d" % (len(lin), IMXSIG))
    return lin


def closest_grid(grid, wn_0):
    """spect_classes.py:1937-1943"""
    ind = np.argmin(np.abs(grid - wn_0))
    return ind, grid[ind]


def line_shape(freq, lw, dw, centre, lin_grid):
    """MakeShape, spect_classes.py:1990-2008 on the window lin_grid+centre."""
    lin_grid, pl = _d(lin_grid)
    xbuf = np.empty(IMXSIG)
    shape = np.empty(IMXSIG)
    rc = lib().orc_line_shape(freq, lw, dw, centre, pl, xbuf.ctypes.data_as(_dp),
                              shape.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("line_shape: %d" % rc)
    return shape


class _orc_lut(C.Structure):
    _fields_ = [("g32", C.POINTER(C.c_float)), ("pt", _dp), ("level_energy", _dp),
                ("n_cells", C.c_int), ("n_sets", C.c_int), ("mol", C.c_int), ("iso", C.c_int),
                ("lte_unidentified", C.c_int), ("iso_ratio", C.c_double)]


def weight(v, v1, v2):
    """sbm.weight(v, v1, v2, 'lin') as specified in DESIGN.md 6.2 (the module is missing)."""
    return (v2 - v) / (v2 - v1), (v - v1) / (v2 - v1)


def lut_weights(PTcouples, Pres, Temp):
    """C restatement of LutSet.calculate's node choice -> (cells[4], weights[4])."""
    pt, pp = _d(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
    cell = np.zeros(4, dtype=np.int32)
    w = np.zeros(4)
    L = lib()
    L.orc_lut_weights.restype = C.c_int
    L.orc_lut_weights.argtypes = [_dp, C.c_int, C.c_double, C.c_double, _ip, _dp]
    rc = L.orc_lut_weights(pp, pt.shape[0], float(Pres), float(Temp), cell.ctypes.data_as(_ip),
                           w.ctypes.data_as(_dp))
    if rc == 7:
        raise ValueError('Extrapolating in P')
    if rc:
        raise ValueError('couple not found!')
    return cell, w


def LutSet_calculate(PTcouples, sets, Pres, Temp):
    """LITERAL NumPy restatement of LutSet.calculate + SpectralGcoeff.interpolate
    (spect_main_module.py:997-1066, spect_classes.py:1349-1375) for ONE ctype:
    `sets` is the list of per-cell spectra (or None), PTcouples the list of [P, T]."""
    PTcouples = [list(map(float, pt)) for pt in PTcouples]

    def find(P, T):
        if [P, T] not in PTcouples:
            raise ValueError('{} couple not found!'.format([P, T]))
        return PTcouples.index([P, T])

    def interpolate(s1, v1, s2, v2, v):
        w1, w2 = weight(v, v1, v2)
        return w1 * s1 + w2 * s2

    Ps = np.unique(np.array([PT[0] for PT in PTcouples]))
    Ts = np.unique(np.array([PT[1] for PT in PTcouples]))
    if Pres <= np.min(Ps):
        closest_P1 = np.min(Ps)
        closest_TA = Ts[np.argmin(np.abs(Ts - Temp))]
        closest_TB = Ts[np.argsort(np.abs(Ts - Temp), kind='stable')[1]]
        c1 = sets[find(closest_P1, closest_TA)]
        c2_ = sets[find(closest_P1, closest_TB)]
        if c1 is None or c2_ is None:
            return None
        return interpolate(c1, closest_TA, c2_, closest_TB, Temp)
    elif Pres > np.min(Ps) and Pres <= np.max(Ps):
        closest_P1 = Ps[np.argmin(np.abs(Ps - Pres))]
        closest_P2 = Ps[np.argsort(np.abs(Ps - Pres), kind='stable')[1]]
        closest_T1 = Ts[np.argmin(np.abs(Ts - Temp))]
        closest_T2 = Ts[np.argsort(np.abs(Ts - Temp), kind='stable')[1]]
        c1 = sets[find(closest_P1, closest_T1)]
        c2_ = sets[find(closest_P1, closest_T2)]
        c3 = sets[find(closest_P2, closest_T1)]
        c4 = sets[find(closest_P2, closest_T2)]
        if c1 is None or c2_ is None or c3 is None or c4 is None:
            return None
        c13 = interpolate(c1, closest_P1, c3, closest_P2, Pres)
        c24 = interpolate(c2_, closest_P1, c4, closest_P2, Pres)
        return interpolate(c13, closest_T1, c24, closest_T2, Temp)
    else:
        raise ValueError('Extrapolating in P')


def make_abscoeff_LUTS_fast(lut, Temps, Press, tvib=None):
    """LITERAL NumPy restatement of make_abscoeff_LUTS_fast (spect_main_module.py:2200-2250) for
    one isotopologue.  lut: dict(g32[n_cells,n_sets,3,n_grid] float32, pt, level_energy, mol, iso,
    lte_unidentified).  tvib: [n_sets, n_steps] or None.  Returns (abs[n_steps,n_grid],
    emi[n_steps,n_grid]) WITHOUT isotopic ratio / column (the caller applies them)."""
    g = lut["g32"]
    n_cells, n_sets, _, n_grid = g.shape
    abs_all, emi_all = [], []
    for num, (Pres, Temp) in enumerate(zip(Press, Temps)):
        abs_coeff = np.zeros(n_grid)
        emi_coeff = np.zeros(n_grid)
        Q_part = CalcPartitionSum(lut["mol"], lut["iso"], temp=Temp)
        for s in range(n_sets):
            Gco = {}
            for ct, name in enumerate(('sp_emission', 'ind_emission', 'absorption')):
                sets = [g[c, s, ct].astype(float) for c in range(n_cells)]   # double_precision()
                Gco[name] = LutSet_calculate(lut["pt"], sets, Pres, Temp)
            if lut.get("lte_unidentified"):
                pop = 1 / Q_part
            else:
                vibt = Temp if tvib is None else tvib[s][num]
                pop = Boltz_ratio_nodeg(lut["level_energy"][s], vibt) / Q_part
            abs_coeff = abs_coeff + Gco['absorption'] * pop
            abs_coeff = abs_coeff - Gco['ind_emission'] * pop
            emi_coeff = emi_coeff + Gco['sp_emission'] * pop
        abs_all.append(abs_coeff)
        emi_all.append(emi_coeff)
    return np.array(abs_all), np.array(emi_all)


def _lut_structs(luts):
    keep = []
    arr = (_orc_lut * len(luts))()
    for i, l in enumerate(luts):
        g = np.ascontiguousarray(l["g32"], dtype=np.float32)
        pt = np.ascontiguousarray(np.asarray(l["pt"], dtype=float).reshape(-1, 2))
        le = np.ascontiguousarray(l.get("level_energy") if l.get("level_energy") is not None
                                  else np.zeros(g.shape[1]), dtype=float)
        keep += [g, pt, le]
        arr[i].g32 = g.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].pt = pt.ctypes.data_as(_dp)
        arr[i].level_energy = le.ctypes.data_as(_dp)
        arr[i].n_cells, arr[i].n_sets = g.shape[0], g.shape[1]
        arr[i].mol, arr[i].iso = int(l["mol"]), int(l["iso"])
        arr[i].lte_unidentified = int(bool(l.get("lte_unidentified")))
        arr[i].iso_ratio = float(l["iso_ratio"])
    return arr, keep


def los_rt(luts, n_steps, temp, pres, column, tvib=None, pt0=0, n_pts=None, i0=None,
           solo_absorption=False, materialise=False, n_threads=1):
    """CPU LOS radiances -> rad[n_los, n_pts] (and tau, src when
    materialise=True).  Array conventions as sr_los_steps in include/spectrobot.h."""
    arr, keep = _lut_structs(luts)
    n_grid = luts[0]["g32"].shape[3]
    if n_pts is None:
        n_pts = n_grid - pt0
    n_steps, pn = _i(n_steps)
    temp, ptm = _d(temp)
    pres, ppr = _d(pres)
    column = np.ascontiguousarray(column, dtype=float)
    if column.ndim == 2:
        column = column[None]
    n_los, n_steps_max = temp.shape
    n_sets_max = 0
    ptv = None
    if tvib is not None:
        tvib, ptv = _d(tvib)
        n_sets_max = tvib.shape[1]
    rad = np.empty((n_los, n_pts))
    tau = src = None
    if materialise:
        tau = np.zeros((n_los, n_steps_max, n_pts))
        src = np.zeros((n_los, n_steps_max, n_pts))
    i0p = None
    if i0 is not None:
        i0, i0p = _d(i0)
    _, pc = _d(CONSTS)
    L = lib()
    L.orc_los_rt.restype = C.c_int
    L.orc_los_rt.argtypes = [C.POINTER(_orc_lut), C.c_int, C.c_long, C.c_int, C.c_int, C.c_int,
                             _ip, _dp, _dp, _dp, _dp, _dp, C.c_long, C.c_long, _dp, C.c_int, _dp,
                             _dp, _dp, C.c_int]
    rc = L.orc_los_rt(arr, len(luts), n_grid, n_los, n_steps_max, n_sets_max, pn, ptm, ppr,
                      column.ctypes.data_as(_dp), ptv, pc, int(pt0), int(n_pts), i0p,
                      int(bool(solo_absorption)), rad.ctypes.data_as(_dp),
                      None if tau is None else tau.ctypes.data_as(_dp),
                      None if src is None else src.ctypes.data_as(_dp), int(n_threads))
    if rc == 7:
        raise ValueError('Extrapolating in P')
    if rc:
        raise ValueError('couple not found!')
    return (rad, tau, src) if materialise else rad


def los_layers(tau, src, n_steps, i0=None, solo_absorption=False):
    """CPU recursion over materialised layers (orc_los_layers)."""
    tau, pt_ = _d(tau)
    src, ps = _d(src)
    n_steps, pn = _i(n_steps)
    n_los, n_steps_max, n_pts = tau.shape
    rad = np.empty((n_los, n_pts))
    i0p = None
    if i0 is not None:
        i0, i0p = _d(i0)
    L = lib()
    L.orc_los_layers.restype = None
    L.orc_los_layers.argtypes = [_dp, _dp, _ip, C.c_int, C.c_int, C.c_long, _dp, C.c_int, _dp]
    L.orc_los_layers(pt_, ps, pn, n_los, n_steps_max, n_pts, i0p, int(bool(solo_absorption)),
                     rad.ctypes.data_as(_dp))
    return rad


def los_layers_jac(tau, emi, dfrac, n_steps, tau_g=None, emi_g=None, i0=None,
                   solo_absorption=False):
    """CPU radiances + analytic Jacobians of the layer recursion in closed-sum form
    (orc_los_layers_jac): (rad [n_los, n_pts], jac [n_los, n_par, n_pts])."""
    tau, pt_ = _d(tau)
    emi, pe = _d(emi)
    dfrac, pf = _d(dfrac)
    n_steps, pn = _i(n_steps)
    n_los, n_steps_max, n_pts = tau.shape
    n_par = dfrac.shape[2]
    ptg = peg = None
    if tau_g is not None:
        tau_g, ptg = _d(tau_g)
        emi_g, peg = _d(emi_g)
    i0p = None
    if i0 is not None:
        i0, i0p = _d(i0)
    rad = np.empty((n_los, n_pts))
    jac = np.empty((n_los, n_par, n_pts))
    L = lib()
    L.orc_los_layers_jac.restype = None
    L.orc_los_layers_jac.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_int, _ip, C.c_int, C.c_int,
                                     C.c_long, _dp, C.c_int, _dp, _dp]
    L.orc_los_layers_jac(pt_, pe, ptg, peg, pf, n_par, pn, n_los, n_steps_max, n_pts, i0p,
                         int(bool(solo_absorption)), rad.ctypes.data_as(_dp),
                         jac.ctypes.data_as(_dp))
    return rad, jac


def gcoeff_cell(lines, grid, T, P, MM, n_sets, n_threads=1, lin_grid=None):
    """One LUT cell on the CPU: out[n_sets, 3, n_grid] (see orc_gcoeff_cell).

    `lines` is a dict of equal-length arrays: freq, a_coeff, air_broad, t_dep, e_lower, g_up, g_lo,
    e_vib_up, e_vib_lo (float64) and up_set, lo_set (int32)."""
    grid, pg = _d(grid)
    if lin_grid is None:
        lin_grid = line_window_offsets(grid)
    lin_grid, pl = _d(lin_grid)
    keep = [_d(lines[k]) for k in ("freq", "a_coeff", "air_broad", "t_dep", "e_lower", "g_up",
                                   "g_lo", "e_vib_up", "e_vib_lo")]
    up, pu = _i(lines["up_set"])
    lo, plo = _i(lines["lo_set"])
    _, pc = _d(CONSTS)
    out = np.empty((n_sets, 3, len(grid)))
    rc = lib().orc_gcoeff_cell(len(up), *[p for _, p in keep], pu, plo, pg, len(grid), pl,
                               float(T), float(P), float(MM), pc, int(n_sets),
                               out.ctypes.data_as(_dp), int(n_threads))
    if rc:
        raise RuntimeError("gcoeff_cell: humliv_bb status %d" % rc)
    return out


def convolve_to_grid_from_irregular(grid, spectrum, new_grid, spectral_widths, n_sigma=5.):
    """Literal NumPy restatement of SpectralObject.convolve_to_grid_from_irregular
    (spect_classes.py:883-918) with gaussian (:1926-1934) and conv_single (:1162-1164)."""
    grid = np.asarray(grid, dtype=float)
    spectrum = np.asarray(spectrum, dtype=float)
    out = np.zeros(len(new_grid), dtype=float)
    for num, (freq, wid) in enumerate(zip(new_grid, spectral_widths)):
        ok_po = (grid >= freq - n_sigma * wid) & (grid <= freq + n_sigma * wid)
        lin_grid_ok = grid[ok_po]
        spect_old = spectrum[ok_po]
        if len(spect_old) == 0:
            out[num] = 0.0
            continue
        fac = 1 / (wid * np.sqrt(2. * np.pi))
        gauss = fac * np.exp(-0.5 * ((lin_grid_ok - freq) / wid) ** 2)
        out[num] = np.trapezoid(spect_old * gauss, x=lin_grid_ok)
    return out


def FOV_integr_1D(spectra, grid, pixel_rot=0.0):
    """Literal restatement of FOV_integr_1D (spect_main_module.py:3342-3374): RectBivariateSpline
    (kx=ky=2) through the low / centre / up LOS spectra and scipy quad of spline*trapezoid response
    for every wavenumber.  spectra: [3][n]; returns [n]."""
    from scipy import integrate
    from scipy.interpolate import RectBivariateSpline as spline2D
    pixel_rot = abs(np.pi * pixel_rot / 180.0)
    dmax = np.sqrt(2.) / 2. * np.cos(np.pi / 4 - pixel_rot)
    delta = dmax - np.sin(pixel_rot)
    esse = 1 / np.cos(pixel_rot)
    x_integ = np.array([-dmax, 0, dmax])
    intens_spl = spline2D(x_integ, np.asarray(grid, dtype=float), np.asarray(spectra, dtype=float),
                          kx=2, ky=2)

    def integrand(x, ww):
        if abs(x) <= delta:
            return intens_spl(x, ww)[0, 0] * esse
        return intens_spl(x, ww)[0, 0] * esse * abs(dmax - abs(x)) / (dmax - delta)

    return np.array([integrate.quad(integrand, -dmax, dmax, args=(ww,))[0] for ww in grid])


def los_steps_build(z, temp, pres, vmr, origins, directions, tvib=None, tvib_on=None,
                    lat_edges=None, radius=2575.0, top=1500.0, delta_x=5.0, max_T_variation=5.0,
                    max_Plog_variation=1.0, masks=None, jac_gas=-1, sza_nodes=None, sun=None,
                    sza_fixed=None, photon_order=False, max_opt_depth=None, sigma_peak=None,
                    lat_centres=None):
    """CPU restatement of the LOS geometry + radtran-step specification with
    the oracle's own Curtis-Godson integrals (orc_curgod_1..4, curgods.f:2-98).  Array conventions
    as sr_atmosphere / sr_los_rays / sr_steps_opt in include/spectrobot.h: tvib may carry an SZA
    axis [n_gas][n_sets][n_band][n_sza][n_z] with sza_nodes (linear between nodes, clamped); the SZA
    of a sample is the angle between its position and `sun` [n_los][3], or sza_fixed[n_los];
    photon_order reverses the sample sequence; max_opt_depth closes a step when sum_gas sigma_peak *
    column exceeds it.  lat_centres (instead of lat_edges): the rows are given AT these latitudes
    and every profile value is the linear blend of the two rows that bracket the sample's latitude,
    each interpolated in altitude (and SZA) first, constant outside the first / last centre
    (`['lin', ...]` profiles of radtran_3Dvs2D_radtrans_new.py:82-111; P rows are interpolated
    log-linearly in altitude and blended linearly).  Returns per LOS a dict with n_steps, temp[], pres[], column[n_gas][],
    tvib[n_gas][n_sets][], dfrac[][n_par].  Plain Python loops: small cases only."""
    kb_hpa = 1.38065e-19
    z = np.asarray(z, dtype=float)
    temp = np.asarray(temp, dtype=float).reshape(-1, len(z))
    pres = np.asarray(pres, dtype=float).reshape(-1, len(z))
    n_band = temp.shape[0]
    vmr = np.asarray(vmr, dtype=float).reshape(-1, n_band, len(z))
    n_gas = vmr.shape[0]
    n_sets = 0 if tvib_on is None else np.asarray(tvib_on).reshape(n_gas, -1).shape[1]
    if n_sets:
        tvib_on = np.asarray(tvib_on).reshape(n_gas, n_sets)
        n_sza = 1 if sza_nodes is None else max(len(sza_nodes), 1)
        tvib = None if tvib is None else np.asarray(tvib, dtype=float).reshape(n_gas, n_sets, n_band,
                                                                               n_sza, len(z))
    if masks is not None:   # [n_par][n_z] or [n_par][n_band][n_z]
        masks = np.asarray(masks, dtype=float)
        if masks.ndim == 2:
            masks = np.repeat(masks[:, None, :], n_band, axis=1)
    out = []
    origins = np.asarray(origins, dtype=float).reshape(-1, 3)
    if sun is not None:
        sun = np.asarray(sun, dtype=float)
        sun = np.broadcast_to(sun / np.linalg.norm(sun, axis=-1, keepdims=True), origins.shape)
    if sza_fixed is not None:
        sza_fixed = np.broadcast_to(np.asarray(sza_fixed, dtype=float), (len(origins),))
    for il, (o, d) in enumerate(zip(origins, np.asarray(directions, dtype=float).reshape(-1, 3))):
        st = -float(np.dot(o, d))
        rt = float(np.linalg.norm(o + st * d))
        r_top = radius + top
        res = dict(n_steps=0, temp=[], pres=[], column=[[] for _ in range(n_gas)],
                   tvib=[[[] for _ in range(n_sets)] for _ in range(n_gas)], dfrac=[])
        out.append(res)
        if rt >= r_top:
            continue
        half = mt.sqrt(r_top ** 2 - rt ** 2)
        s_near, s_far = st - half, st + half
        if rt < radius:
            s_far = st - mt.sqrt(radius ** 2 - rt ** 2)
        kmax = int(mt.floor(half / delta_x - 1e-9))
        inner = st + delta_x * np.arange(kmax, -kmax - 1, -1)
        inner = inner[(inner < s_far - 1e-6) & (inner > s_near + 1e-6)]
        s = np.concatenate([[s_far], inner, [s_near]])
        if photon_order:
            s = s[::-1]
        pts = o[None, :] + s[:, None] * d[None, :]
        r = np.sqrt((pts ** 2).sum(axis=1))
        alt = r - radius
        sza = np.zeros(len(s))
        if sza_fixed is not None:
            sza[:] = sza_fixed[il]
        elif sun is not None:
            sza = np.degrees(np.arccos(np.clip(pts @ sun[il] / r, -1.0, 1.0)))
        band = np.zeros(len(s), dtype=int)
        wlat = np.zeros(len(s))
        lat_lin = lat_centres is not None and n_band > 1
        if lat_lin:
            cen = np.asarray(lat_centres, dtype=float)
            lat = np.degrees(np.arcsin(pts[:, 2] / r))
            band = np.clip(np.searchsorted(cen, lat, side='right') - 1, 0, n_band - 2)
            wlat = np.clip((lat - cen[band]) / (cen[band + 1] - cen[band]), 0.0, 1.0)
        elif n_band > 1:
            lat = np.degrees(np.arcsin(pts[:, 2] / r))
            band = np.clip(np.searchsorted(lat_edges, lat, side='right') - 1, 0, n_band - 1)

        def blend(row_value):   # row_value(i, b): value of row b at sample i
            if not lat_lin:
                return np.array([row_value(i, b) for i, b in enumerate(band)])
            return np.array([(1.0 - w) * row_value(i, b) + w * row_value(i, b + 1)
                             for i, (b, w) in enumerate(zip(band, wlat))])

        at = lambda tab: blend(lambda i, b: np.interp(alt[i], z, tab[b]))  # noqa: E731

        def at_sza(tab):   # [n_band][n_sza][n_z]: linear in z, then linear between SZA nodes
            if tab.shape[1] == 1:
                return at(tab[:, 0])
            return blend(lambda i, b: np.interp(
                sza[i], sza_nodes, np.array([np.interp(alt[i], z, tab[b, j]) for j in range(tab.shape[1])])))

        T = at(temp)
        P = blend(lambda i, b: np.exp(np.interp(alt[i], z, np.log(pres[b]))))
        nd = P / (kb_hpa * T)
        x = np.abs(s[0] - s) * 1.e5
        lnP = np.log(P)
        n = len(s)
        vm = [at(vmr[m]) for m in range(n_gas)]
        bounds, i0 = [], 0
        for i in range(1, n):
            sl = slice(i0, i + 1)
            too_thick = False
            if max_opt_depth is not None and max_opt_depth > 0.0:
                tau = sum(sigma_peak[m] * curgod(2, nd[sl], vm[m][sl], x[sl]) for m in range(n_gas))
                too_thick = tau > max_opt_depth
            if (T[sl].max() - T[sl].min() > max_T_variation or
                    lnP[sl].max() - lnP[sl].min() > max_Plog_variation or too_thick) and i - i0 >= 2:
                bounds.append((i0, i - 1))
                i0 = i - 1
        if n >= 2:
            bounds.append((i0, n - 1))
        res["n_steps"] = len(bounds)
        for a, e in bounds:
            sl = slice(a, e + 1)
            ones = np.ones(e - a + 1)
            air = curgod(1, nd[sl], x[sl])
            t_cg = curgod(4, nd[sl], ones, T[sl], x[sl]) / air
            res["temp"].append(t_cg)
            res["pres"].append(curgod(4, nd[sl], ones, P[sl], x[sl]) / air)
            for m in range(n_gas):
                col = curgod(2, nd[sl], vm[m][sl], x[sl])
                res["column"][m].append(col)
                for j in range(n_sets):
                    if tvib_on[m, j] > 0:
                        tv = curgod(3, nd[sl], vm[m][sl], at_sza(tvib[m, j])[sl], x[sl]) / col
                    else:
                        tv = t_cg if tvib_on[m, j] == 0 else 100.0
                    res["tvib"][m][j].append(tv)
            if masks is not None:
                row = []
                col = res["column"][jac_gas][-1] if 0 <= jac_gas < n_gas else 0.0
                for q in range(len(masks)):
                    mk = at(masks[q])[sl]
                    dcol = curgod(2, nd[sl], mk, x[sl]) if np.any(mk != 0.0) else 0.0
                    row.append(dcol / col if col != 0.0 else 0.0)
                res["dfrac"].append(row)
    return out
