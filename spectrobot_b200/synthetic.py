"""Synthetic inputs of the shapes SURVEY.md section 8(d) names (seed 20067).

The reference ships no data (no HITRAN file, no Titan profiles, no vibrational temperatures, no
VIMS pixels), so every benchmark and parity test runs on these generators.  They only produce
INPUTS (line tables, grids, atmospheres, cells, LOS step tables); no hot-path arithmetic lives
here.
"""
import math as mt

import numpy as np
import scipy.constants as const

SEED = 20067

# spect_classes.py:44-47
h_cgs = const.physical_constants['Planck constant'][0] * 1.e7
c_cgs = const.c * 1.e2
k_cgs = const.physical_constants['Boltzmann constant'][0] * 1.e7
c2 = h_cgs * c_cgs / k_cgs

# CH4 iso 1 (molparam.txt:49): molar mass and abundance
CH4_MM = 16.0313
CH4_RATIO = 0.988274

# 12-level CH4 keep-list stand-in (radtran_3D_ch4.py:281): vibrational energies in cm-1
CH4_LEVEL_ENERGIES = np.array([0.0, 1310.76, 1533.33, 2587.04, 2614.26, 2830.32, 2846.08,
                               2916.48, 3019.49, 3063.65, 3064.48, 3100.0])


def spectral_grid(w0, w1, step=5.e-4):
    """prepare_spe_grid (spect_main_module.py:1267): np.arange, NOT w0+i*step."""
    return np.arange(w0, w1 + step / 2, step, dtype=float)


def level_bands(n_levels):
    """(upper, lower) level-index pairs the synthetic lines are drawn from."""
    bands = [(u, 0) for u in range(1, n_levels)]
    for u in range(5, n_levels):
        bands.append((u, 1))
        bands.append((u, 2))
    return bands


def line_table(n_lines, w0, w1, n_levels=12, q296=590.52, iso_ratio=CH4_RATIO, seed=SEED,
               level_energies=None, frac_unlinked=0.0):
    """Synthetic HITRAN-like line list as a dict of arrays.

    A_coeff is derived from Strength with the inverse of CalcStrength_from_Einstein at 296 K
    (spect_classes.py:291-309) so the LTE identity holds.  `frac_unlinked` marks that fraction of
    lines as not linked to a known level (up_set/lo_set = -1), which the non-LTE path must drop
    (spect_classes.py:1384-1388).
    """
    rng = np.random.default_rng(seed)
    if level_energies is None:
        level_energies = CH4_LEVEL_ENERGIES[:n_levels] if n_levels <= 12 else \
            np.concatenate([CH4_LEVEL_ENERGIES, 3100.0 + 40.0 * np.arange(1, n_levels - 11)])
    level_energies = np.asarray(level_energies, dtype=float)
    freq = np.sort(rng.uniform(w0 + 3.3, w1 - 3.3, n_lines))
    strength = 10.0 ** rng.uniform(-28.0, -19.0, n_lines)
    air = rng.uniform(0.04, 0.08, n_lines)
    tdep = rng.uniform(0.55, 0.85, n_lines)
    pshift = rng.uniform(-0.01, 0.0, n_lines)
    e_lower = rng.uniform(0.0, 2500.0, n_lines)
    J = rng.integers(0, 26, n_lines)
    gs = np.array([2, 3, 5])
    g_lo = (gs[rng.integers(0, 3, n_lines)] * (2 * J + 1)).astype(float)
    dJ = rng.integers(-1, 2, n_lines)
    Jup = np.clip(J + dJ, 0, 26)
    g_up = (gs[rng.integers(0, 3, n_lines)] * (2 * Jup + 1)).astype(float)
    if n_levels > 1:
        bands = np.array(level_bands(n_levels))
        pick = rng.integers(0, len(bands), n_lines)
        up = bands[pick, 0].astype(np.int32)
        lo = bands[pick, 1].astype(np.int32)
        e_vib_up = level_energies[up]
        e_vib_lo = level_energies[lo]
    else:  # LTE isotopologue: single set 'all', E_vib = 0 (spect_classes.py:318-321)
        up = np.zeros(n_lines, dtype=np.int32)
        lo = np.zeros(n_lines, dtype=np.int32)
        e_vib_up = np.zeros(n_lines)
        e_vib_lo = np.zeros(n_lines)
    if frac_unlinked > 0.0:
        drop = rng.uniform(size=n_lines) < frac_unlinked
        up = np.where(drop, -1, up).astype(np.int32)
        lo = np.where(drop, -1, lo).astype(np.int32)
    # calc_A_coeff_from_strength, spect_classes.py:303-305
    T = 296.0
    b21 = strength * (4 * np.pi * q296) / ((np.exp(-c2 * e_lower / T)
                                            - np.exp(-c2 * (e_lower + freq) / T))
                                           * h_cgs * c_cgs * freq * g_up * iso_ratio)
    a_coeff = b21 * (2 * h_cgs * c_cgs ** 2 * freq ** 3)
    return dict(freq=freq, strength=strength, a_coeff=a_coeff, air_broad=air, t_dep=tdep,
                p_shift=pshift, e_lower=e_lower, g_up=g_up, g_lo=g_lo,
                e_vib_up=np.ascontiguousarray(e_vib_up), e_vib_lo=np.ascontiguousarray(e_vib_lo),
                up_set=up, lo_set=lo, level_energies=level_energies, n_sets=max(n_levels, 1))


# ---------------------------------------------------------------------------------------------
# Titan-like atmosphere: z = 0..1500 km step 10 km, 7 latitude bands
# ---------------------------------------------------------------------------------------------
R_TITAN_KM = 2575.0          # spect_classes.py:32
M_TITAN_KG = 1.3452e23       # spect_classes.py:33
G_NEWTON = 6.67408e-11       # spect_classes.py:36
R_GAS = 8.31446              # spect_classes.py:35
MEAN_MASS = 0.98 * 28.0 + 0.02 * 16.0   # titanatm.py:27
LAT_EDGES = np.array([-90., -75., -60., -30., 30., 60., 75., 90.])   # radtran_3D_ch4.py:70-72


def titan_temperature(z_km, band=3):
    """Analytic Titan-like T(z): 94 K surface, 70 K tropopause at 45 km, 180 K at 300 km,
    150-175 K thermosphere; +-10 K spread over the 7 latitude bands."""
    z = np.asarray(z_km, dtype=float)
    t_low = 94.0 + (70.0 - 94.0) * np.clip(z / 45.0, 0, 1) ** 1.2
    rise = 70.0 + (180.0 - 70.0) * (1 - np.exp(-np.clip(z - 45.0, 0, None) / 70.0)) / \
        (1 - np.exp(-(300.0 - 45.0) / 70.0))
    t_mid = np.where(z <= 45.0, t_low, np.minimum(rise, 180.0))
    thermo = 162.5 + 12.5 * np.cos((z - 300.0) / 1200.0 * 2 * np.pi)
    w = np.clip((z - 300.0) / 150.0, 0, 1)
    t = (1 - w) * t_mid + w * thermo
    return t + (band - 3) * (10.0 / 3.0) * np.clip(z / 300.0, 0, 1)


def titan_atmosphere(n_bands=7, z_top=1500.0, dz=10.0):
    """Returns dict(z[nz], temp[nb,nz], pres[nb,nz] hPa, ndens[nb,nz] cm-3, vmr_ch4[nz])."""
    z = np.arange(0.0, z_top + dz / 2, dz)
    temp = np.stack([titan_temperature(z, b) for b in range(n_bands)])
    pres = np.empty_like(temp)
    g0 = G_NEWTON * M_TITAN_KG / (R_TITAN_KM * 1e3) ** 2
    for b in range(n_bands):
        g = g0 * (R_TITAN_KM / (R_TITAN_KM + z)) ** 2
        H = R_GAS * temp[b] / (MEAN_MASS * 1e-3 * g) / 1e3           # km
        invH = 1.0 / H
        integ = np.concatenate([[0.0], np.cumsum(0.5 * (invH[1:] + invH[:-1]) * dz)])
        pres[b] = 1467.0 * np.exp(-integ)
    kb_hpa = 1.38065e-19                                             # spect_classes.py:34
    ndens = pres / (kb_hpa * temp)
    vmr = np.full_like(z, 0.015)
    return dict(z=z, temp=temp, pres=pres, ndens=ndens, vmr_ch4=vmr, lat_edges=LAT_EDGES[:n_bands + 1])


def vib_temperatures(z_km, temp, level_energies, sza_deg=60.0):
    """Smooth non-LTE vibrational temperatures T_vib = T + dT_lev(z, SZA), growing above 400 km."""
    z = np.asarray(z_km, dtype=float)
    n_lev = len(level_energies)
    mu = mt.cos(mt.radians(sza_deg))
    grow = np.clip((z - 400.0) / 600.0, 0, None) ** 1.5
    out = np.empty((n_lev,) + z.shape)
    for i, e in enumerate(level_energies):
        out[i] = temp + (0.0 if e == 0.0 else (20.0 + 0.02 * e) * grow * (0.3 + 0.7 * mu))
    return out


def limb_los_steps(tangent_km, bands, szas, atm, level_energies, max_T_variation=5.0,
                   max_Plog_variation=1.0, delta_x=5.0, vmr=0.015, n_steps_max=None):
    """Synthetic step tables for limb lines of sight (far end -> observer), SURVEY 8d.

    For every LOS the ray is sampled every delta_x km between its two top-of-atmosphere
    intersections, consecutive samples are merged into steps while the temperature varies by less
    than max_T_variation and ln P by less than max_Plog_variation (the radtran_opt of
    radtran_3D_ch4.py:200-202), and each step gets column-weighted (Curtis-Godson) T and P, the gas
    column (molecules cm-2) and per-level vibrational temperatures at the weighted altitude.
    Returns dict(n_steps, temp, pres, column[1,...], tvib[1,n_lev,...]) padded to n_steps_max."""
    z = atm["z"]
    z_top = z[-1]
    out = []
    for ht, b, sza in zip(tangent_km, bands, szas):
        rt = R_TITAN_KM + ht
        smax = mt.sqrt((R_TITAN_KM + z_top) ** 2 - rt ** 2)
        s = np.arange(-smax, smax + delta_x / 2, delta_x)
        s[-1] = smax
        zz = np.sqrt(rt ** 2 + s ** 2) - R_TITAN_KM
        T = np.interp(zz, z, atm["temp"][b])
        lnP = np.interp(zz, z, np.log(atm["pres"][b]))
        nd = np.exp(np.interp(zz, z, np.log(atm["ndens"][b]))) * vmr
        steps = []
        i0 = 0
        # running extremes of T and ln P over the samples i0..i of the open step
        tl = th = T[0]
        pl = ph = lnP[0]
        for i in range(1, len(s)):
            ti, pi = T[i], lnP[i]
            tl2, th2 = (ti if ti < tl else tl), (ti if ti > th else th)
            pl2, ph2 = (pi if pi < pl else pl), (pi if pi > ph else ph)
            if (th2 - tl2 > max_T_variation or ph2 - pl2 > max_Plog_variation) and i - i0 >= 2:
                steps.append((i0, i - 1))
                i0 = i - 1
                tl, th = min(T[i0], ti), max(T[i0], ti)
                pl, ph = min(lnP[i0], pi), max(lnP[i0], pi)
            else:
                tl, th, pl, ph = tl2, th2, pl2, ph2
        steps.append((i0, len(s) - 1))
        rows = []
        for a, e in steps:
            sl = slice(a, e + 1)
            ds = np.diff(s[sl]) * 1e5                                   # cm
            w = 0.5 * (nd[sl][1:] + nd[sl][:-1]) * ds
            col = w.sum()
            mid = lambda v: 0.5 * (v[sl][1:] + v[sl][:-1])
            Tcg = (mid(T) * w).sum() / col
            Pcg = mt.exp((mid(lnP) * w).sum() / col)
            zcg = (mid(zz) * w).sum() / col
            rows.append((Tcg, Pcg, col, zcg))
        rows = np.array(rows)
        tv = vib_temperatures(rows[:, 3], rows[:, 0], level_energies, sza)
        out.append((rows, tv))
    nmax = max(len(r) for r, _ in out) if n_steps_max is None else n_steps_max
    n_los, n_lev = len(out), len(level_energies)
    n_steps = np.zeros(n_los, dtype=np.int32)
    temp = np.full((n_los, nmax), 100.0)
    pres = np.full((n_los, nmax), 1e-6)
    column = np.zeros((1, n_los, nmax))
    tvib = np.full((1, n_lev, n_los, nmax), 100.0)
    for l, (rows, tv) in enumerate(out):
        k = min(len(rows), nmax)
        n_steps[l] = k
        temp[l, :k] = rows[:k, 0]
        pres[l, :k] = rows[:k, 1]
        column[0, l, :k] = rows[:k, 2]
        tvib[0, :, l, :k] = tv[:, :k]
    return dict(n_steps=n_steps, temp=temp, pres=pres, column=column, tvib=tvib)


def rect_cells(p_lo, p_hi, t_lo, t_hi, pres_step_log=1.0, temp_step=5.0):
    """Rectangular (P,T) cell set: ln P on multiples of pres_step_log, T on multiples of
    temp_step, covering [p_lo,p_hi] x [t_lo,t_hi] (the shape calc_PT_couples_atmosphere yields for
    an isothermal-range atmosphere)."""
    n0 = mt.floor(mt.log(p_lo) / pres_step_log)
    n1 = mt.ceil(mt.log(p_hi) / pres_step_log)
    ps = [mt.exp(n * pres_step_log) for n in range(n0, n1 + 1)]
    t0 = (mt.floor(t_lo / temp_step) - 1) * temp_step
    t1 = (mt.ceil(t_hi / temp_step) + 1) * temp_step
    ts = np.arange(t0, t1 + 0.5 * temp_step, temp_step)
    return [[p, float(t)] for p in ps for t in ts]


def fixed_cells():
    """The 5-cell check set of spect_main_Titan.py:212-213: T=175 K, P in 1e-3..10 hPa."""
    return [[p, 175.0] for p in (1e-3, 1e-2, 0.1, 1.0, 10.0)]


# ---------------------------------------------------------------------------------------------
# object-level synthetic inputs for the spect_classes / spect_main_module API
# ---------------------------------------------------------------------------------------------
def level_strings(n_levels):
    """HITRAN-like global-quanta strings 'v1 v2 v3 v4 sym', one per synthetic level."""
    return ['%d %d 0 0 1A1' % (i // 4, i % 4) for i in range(n_levels)]


def spect_lines(tab, mol=6, iso=1):
    """List of spect_classes.SpectLine from a line_table() dict; unlinked lines (up_set < 0) get
    level strings that match no level of the synthetic isotopologue."""
    from . import spect_classes as spcl
    strings = level_strings(int(tab["n_sets"]))
    out = []
    for i in range(len(tab["freq"])):
        u, l = int(tab["up_set"][i]), int(tab["lo_set"][i])
        d = dict(Mol=mol, Iso=iso, Freq=float(tab["freq"][i]), Strength=float(tab["strength"][i]),
                 A_coeff=float(tab["a_coeff"][i]), Air_broad=float(tab["air_broad"][i]),
                 Self_broad=0.0, E_lower=float(tab["e_lower"][i]),
                 T_dep_broad=float(tab["t_dep"][i]), P_shift=float(tab["p_shift"][i]),
                 Up_lev_str=strings[u] if u >= 0 else '9 9 9 9 1A1',
                 Lo_lev_str=strings[l] if l >= 0 else '9 9 9 8 1A1', Q_num_up='', Q_num_lo='',
                 others='', g_up=float(tab["g_up"][i]), g_lo=float(tab["g_lo"][i]))
        out.append(spcl.SpectLine(d))
    return out


SZA_NODES = np.array([0.0, 30.0, 50.0, 65.0, 80.0, 95.0])   # nodes of the 3-D T_vib tables (deg)


def vib_temperatures_3d(z_km, temp_bands, level_energies, sza_nodes=SZA_NODES):
    """[n_lev][n_band][n_sza][n_z] vibrational temperatures T_vib(lat band, SZA, z): the shape of
    the reference's 3-D profiles (radtran_3D_ch4.py:249-250)."""
    return np.stack([np.stack([vib_temperatures(z_km, t, level_energies, sz) for sz in sza_nodes],
                              axis=1) for t in np.atleast_2d(temp_bands)], axis=1)


def titan_planet(level_energies=None, n_bands=1, vmr=0.015, nonlte=True, sza_deg=60.0,
                 sza_nodes=None):
    """sbm.Titan with the synthetic atmosphere, CH4 (iso 1) and, when nonlte, one vibrational
    level per entry of level_energies carrying a vibrational-temperature profile - on
    ('alt'), ('lat', 'alt') or, with sza_nodes, on ('lat', 'sza', 'alt') / ('sza', 'alt')."""
    from . import spect_base_module as sbm
    atm = titan_atmosphere(n_bands=max(n_bands, 1))
    planet = sbm.Titan(1500.0)
    if n_bands <= 1:
        grid = sbm.AtmGrid('alt', atm["z"])
        sel = lambda a: a[0]
    else:
        grid = sbm.AtmGrid(['lat', 'alt'], [atm["lat_edges"], atm["z"]])
        sel = lambda a: a
    prof = sbm.AtmProfile(grid, sel(atm["temp"]), 'temp', 'lin')
    prof.add_profile(sel(atm["pres"]), 'pres', 'exp')
    planet.add_atmosphere(prof)
    ch4 = sbm.Molec(6, 'CH4')
    im = ch4.add_iso(1, MM=CH4_MM, ratio=CH4_RATIO, LTE=not nonlte)
    ch4.add_clim(sbm.AtmProfile(grid, np.full_like(sel(atm["temp"]), vmr), 'vmr', 'lin'))
    if level_energies is not None:
        tv = [vib_temperatures(atm["z"], atm["temp"][b], level_energies, sza_deg)
              for b in range(max(n_bands, 1))]
        profs = []
        if sza_nodes is not None:
            tv3 = vib_temperatures_3d(atm["z"], atm["temp"][:max(n_bands, 1)], level_energies, sza_nodes)
            g3 = (sbm.AtmGrid(['sza', 'alt'], [sza_nodes, atm["z"]]) if n_bands <= 1 else
                  sbm.AtmGrid(['lat', 'sza', 'alt'], [atm["lat_edges"], sza_nodes, atm["z"]]))
        for i in range(len(level_energies)):
            if sza_nodes is not None:
                v = tv3[i][0] if n_bands <= 1 else tv3[i]
                profs.append(sbm.AtmProfile(g3, v, 'vibtemp', 'lin') if nonlte else None)
                continue
            v = tv[0][i] if n_bands <= 1 else np.stack([t[i] for t in tv])
            profs.append(sbm.AtmProfile(grid, v, 'vibtemp', 'lin') if nonlte else None)
        im.add_levels(level_strings(len(level_energies)), level_energies, vibtemps=profs)
        im.is_in_LTE = not nonlte
    planet.add_gas(ch4)
    return planet


def vims_pixels(tangent_km, lat=10.0, lon=0.0, dist=1.e5, channels=None, widths=None,
                units='cm_1', sza=60.0, obs_units='ergscm2'):
    """sbm.VIMSPixel list looking at the limb at the given tangent altitudes, with an (empty)
    observation that defines the instrument channels.  sza (scalar or one per pixel): solar zenith
    angle at the tangent point; the sub-solar point is placed on the tangent point's parallel."""
    from . import spect_base_module as sbm
    from . import spect_classes as spcl
    out = []
    szas = np.broadcast_to(np.asarray(sza, dtype=float), (len(tangent_km),))
    for ht, sz in zip(tangent_km, szas):
        obs = None
        if channels is not None:
            obs = spcl.SpectralIntensity(np.zeros(len(channels)), spcl.SpectralGrid(channels, units=units),
                                         units=obs_units)
            obs.bands = spcl.SpectralObject(np.asarray(widths, dtype=float), obs.spectral_grid)
            obs.mask = np.ones(len(channels))
            obs.noise = None
        keys = ['sub_obs_lat', 'sub_obs_lon', 'dist', 'limb_tg_lat', 'limb_tg_lon', 'limb_tg_alt',
                'limb_tg_sza', 'sub_solar_lat', 'sub_solar_lon', 'observation', 'pixel_rot']
        # observer on the equatorial plane 90 deg away in longitude: the ray grazes the limb
        # sub-solar point on the same meridian plane offset in longitude so that the angle between
        # the tangent point and the Sun is `sz`: cos(sz) = sin(lat)sin(ls) + cos(lat)cos(ls)cos(dlon)
        # with ls = 0 -> cos(dlon) = cos(sz)/cos(lat)
        cl = mt.cos(mt.radians(lat))
        dlon = mt.degrees(mt.acos(max(-1.0, min(1.0, mt.cos(mt.radians(float(sz))) / cl))))
        vals = [lat, lon + 90.0, dist, lat, lon, float(ht), float(sz), 0.0, lon - dlon, obs, 0.0]
        out.append(sbm.VIMSPixel(keys, vals))
    return out
