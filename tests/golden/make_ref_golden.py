"""Generates the reference-executed fixtures tests/golden/ref_*: run from the repo root in a
container that has /root/reference (`python tests/golden/make_ref_golden.py`).

Every number written here comes out of the reference's OWN spect_classes.py /
spect_main_module.py executed by ref_exec.py (Fortran replaced by the C restatement, the missing
spect_base_module by ref_sbm_stub.py).  The tests compare the oracle (CPU) and the CUDA library
(GPU) with these files; /root/reference is not needed to run them.
"""
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_exec as R  # noqa: E402

CTYPES = ['sp_emission', 'ind_emission', 'absorption']
LEVELS = ['0 0 0 0 1A1', '0 0 1 0 1F2', '0 1 0 0 1E']          # CH4 ground, nu3, nu2
ENERGIES = [0.0, 3019.4935, 1533.3326]
WN_RANGE = [2998.0, 3006.0]                                      # 16001 points (> imxsig)
CELLS = [[0.05, 150.0], [2.0, 155.0]]
CELL_LTE = [0.3, 170.0]


def case_lines(spcl, seed=20067, n=28):
    """Synthetic CH4 lines as the reference's SpectLine objects.  Strength -> A_coeff through the
    reference's calc_A_coeff_from_strength, so the LTE identity holds."""
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n):
        iso = 1 if i % 4 else 2
        up = int(rng.integers(1, 3))
        up_str, lo_str = LEVELS[up], LEVELS[0]
        if i == 5:
            up_str = '0 2 0 0 1A1'                   # level unknown to the IsoMolec: line dropped
        if i == 9:
            up_str, lo_str = LEVELS[1], LEVELS[2]    # hot band nu3 <- nu2
        d = dict(Mol=6, Iso=iso, Freq=float(np.round(rng.uniform(2995.0, 3009.0), 6)),
                 Strength=float('%.3e' % 10 ** rng.uniform(-24, -20)), A_coeff=0.0,
                 Air_broad=round(rng.uniform(0.04, 0.08), 4), Self_broad=0.07,
                 E_lower=round(rng.uniform(0, 1500), 4), T_dep_broad=round(rng.uniform(.55, .85), 2),
                 P_shift=-0.005, Up_lev_str=up_str.rjust(15), Lo_lev_str=lo_str.rjust(15),
                 Q_num_up=' ' * 15, Q_num_lo=' ' * 15, others=' ' * 19,
                 g_up=float(3 * (2 * rng.integers(1, 20) + 1)),
                 g_lo=float(3 * (2 * rng.integers(1, 20) + 1)))
        lin = spcl.SpectLine([d[k] for k in spcl.cose_hit], nomi=spcl.cose_hit)
        lin.calc_A_coeff_from_strength(set_attr=True)
        lin.A_coeff = float('%.3e' % lin.A_coeff)    # what a HITRAN record can carry
        lines.append(lin)
    lines.sort(key=lambda l: l.Freq)
    return lines


def case_isomolecs(sbm):
    iso1 = sbm.IsoMolec(6, 1)
    iso1.add_levels(LEVELS, ENERGIES)
    iso2 = sbm.IsoMolec(6, 2)                        # no levels: LTE, one set 'all'
    return iso1, iso2


def read_stream(fn):
    with open(fn, 'rb') as f:
        pts = pickle.load(f)
        return pts, [pickle.load(f) for _ in pts]


def main():
    spcl, smm, sbm = R.load()
    out = dict()
    work = tempfile.mkdtemp() + '/'
    cwd = os.getcwd()
    os.chdir(work)                                   # the reference appends to ./control_spectrobot

    # ---- line database: written with the reference's Print_hitran, read with its reader -------
    lines = case_lines(spcl)
    par = os.path.join(HERE, 'ref_lines.par')
    with open(par, 'w') as f:
        for lin in lines:
            lin.Print_hitran(ofile=f)
    rd = spcl.read_line_database(par)
    dec = (lambda v: v.decode() if isinstance(v, bytes) else v)
    for k in spcl.cose_hit:
        v = [dec(getattr(l, k)) for l in rd]
        out['db_' + k] = np.array(v)
    # the fixture lines ARE the re-read lines: what a reference run on this file would use
    for l in rd:
        for k in ('Up_lev_str', 'Lo_lev_str', 'Q_num_up', 'Q_num_lo', 'others'):
            setattr(l, k, dec(getattr(l, k)))
        l.Mol, l.Iso = int(l.Mol), int(l.Iso)
    lines = rd
    iso1, iso2 = case_isomolecs(sbm)

    # ---- A3 / A4 / A9: widths, G coefficients, strengths, partition sums ------------------------
    PT = [(150.0, 0.05), (155.0, 2.0), (92.3, 800.0), (181.7, 1.3e-5), (70.0, 130.0)]
    out['phys_PT'] = np.array(PT)
    out['phys_widths'] = np.array([[lin.CheckWidths(T, P, iso1.MM if lin.Iso == 1 else iso2.MM)
                                    for lin in lines] for T, P in PT])         # (dw, lw, shift)
    with R.quiet():
        out['phys_gcoeff'] = np.array([[[lin.Calc_Gcoeffs(T, isomolec=iso1 if lin.Iso == 1 else None)[c]
                                         for c in CTYPES] for lin in lines] for T, P in PT])
    out['phys_link_ok'] = np.array([bool(lin.LinkToMolec(iso1)) if lin.Iso == 1 else False
                                    for lin in lines])
    out['phys_strength_T'] = np.array([[lin.CalcStrength(T) for lin in lines] for T, P in PT])
    with R.quiet():
        # (the reference raises for a line with an unknown level, spcl:245-251: NaN there)
        out['phys_strength_einstein'] = np.array(
            [[lin.CalcStrength_from_Einstein(T, isomolec=iso1 if lin.Iso == 1 else None)
              if (lin.Iso == 2 or ok) else (np.nan, np.nan)
              for lin, ok in zip(lines, out['phys_link_ok'])] for T, P in PT])
    out['phys_A_from_strength'] = np.array([lin.calc_A_coeff_from_strength() for lin in lines])
    qt = np.array([60.0, 70.0, 84.9, 85.0, 92.3, 150.0, 181.7, 296.0, 500.0, 1000.0, 2985.0])
    out['q_temps'] = qt
    out['q_molisos'] = np.array([[6, 1], [6, 2], [5, 1], [23, 1], [26, 1]])
    out['q_values'] = np.array([[spcl.CalcPartitionSum(m, i, temp=t) for t in qt]
                                for m, i in out['q_molisos']])

    # ---- A1 / A2(through Python) / A5 / A6 / A7: a LUT built by the reference --------------------
    sp = smm.prepare_spe_grid(WN_RANGE).spectral_grid
    out['grid'] = sp.grid
    with R.quiet():
        shp = spcl.calc_shapes_lines(sp, [l for l in lines if l.Iso == 1], CELLS[0][1], CELLS[0][0],
                                     iso1, n_threads=2)
    out['shape_freq'] = np.array([l.Freq for l in shp])
    out['shape_centre'] = np.array([spcl.closest_grid(sp, l.Freq)[0] for l in shp])
    out['shape_first'] = np.array([l.shape.spectral_grid.grid[0] for l in shp])
    pick = [0, len(shp) // 2, len(shp) - 1]
    out['shape_pick'] = np.array(pick)
    out['shape_spectra'] = np.array([shp[i].shape.spectrum for i in pick])
    out['shape_gcoeff'] = np.array([[l.G_coeffs[c] for c in CTYPES] for l in shp])

    lut1 = smm.LookUpTable(iso1, sp.wn_range(), False)
    with R.quiet():
        lut1.make(sp, lines, CELLS, cartLUTs=work, n_threads=2)
    cells = np.zeros((len(CELLS), len(LEVELS), 3, len(sp.grid)))
    for s, lev in enumerate(iso1.levels):
        pts, sets = read_stream(lut1.sets[lev].filename)
        assert pts == CELLS
        for c, set_ in enumerate(sets):
            for k, ct in enumerate(CTYPES):
                cells[c, s, k] = set_[ct].spectrum
    out['cells_nonlte'] = cells[0]
    out['cells_nonlte_b8'] = cells[1][:, :, ::8]
    out['cells_PT'] = np.array(CELLS)
    shutil.copy(lut1.sets['lev_01'].filename, os.path.join(HERE, 'ref_LUT_mol06_iso1_nonLTE_lev_01.pic'))

    lut2 = smm.LookUpTable(iso2, sp.wn_range(), True)
    with R.quiet():
        lut2.make(sp, lines, [CELL_LTE], cartLUTs=work, n_threads=1)
    pts, sets = read_stream(lut2.sets['all'].filename)
    out['cells_lte'] = np.array([sets[0][ct].spectrum for ct in CTYPES])
    out['cells_lte_PT'] = np.array(CELL_LTE)

    # ---- f3: split + compressed LUT files written by the reference -----------------------------
    with R.quiet():
        allL, n_split, sp_grids = smm.split_and_compress_LUTS(sp, {('CH4', 1): lut1}, work, 2, n_split=3)
    L = allL[('CH4', 1)]
    out['split_lens'] = np.array([len(g.grid) for g in sp_grids])
    shutil.copy(L.splitfiles[1], os.path.join(HERE, 'ref_LUT_csplit01_mol06_iso1_nonLTE.pic'))
    with R.quiet():
        L.load_split(1)
    out['split1_lev01_cell0_abs'] = np.zeros(0) if L.sets['lev_01'].sets[0]['absorption'] is None \
        else L.sets['lev_01'].sets[0]['absorption'].spectrum
    out['split1_lev01_cell0_sp'] = L.sets['lev_01'].sets[0]['sp_emission'].spectrum
    out['split1_none'] = np.array([[[L.sets[lev].sets[c][ct] is None for ct in CTYPES]
                                    for lev in iso1.levels] for c in range(len(CELLS))])

    # ---- A10 / A11: LutSet.calculate and make_abscoeff_LUTS_fast on a small synthetic LUT ------
    rng = np.random.default_rng(7)
    n_g = 48
    sg = spcl.SpectralGrid(np.arange(n_g) * 5e-4 + 3000.0, units='cm_1')
    Ps = [1.0e-4, 3.0e-3, 0.12, 2.5]
    Ts = [140.0, 145.0, 150.0, 155.0, 160.0]
    PTs = [[p, t] for p in Ps for t in Ts]
    g32 = (10 ** rng.uniform(-26, -20, size=(len(PTs), len(LEVELS), 3, n_g))).astype(np.float32)
    g32[:, 0, 0] = 0.0                                # ground state: no emission from it (None)
    g32[:, 0, 1] = 0.0
    g32[:, 2, 2] = 0.0
    lutS = smm.LookUpTable(iso1, sg.wn_range(), False)
    for s, lev in enumerate(iso1.levels):
        st = smm.LutSet(6, 1, iso1.MM, level=getattr(iso1, lev))
        st.PTcouples = [list(pt) for pt in PTs]
        st.spectral_grid = sg
        for c, (p, t) in enumerate(PTs):
            d = dict()
            for k, ct in enumerate(CTYPES):
                if not np.any(g32[c, s, k]):
                    d[ct] = None
                    continue
                co = spcl.SpectralGcoeff(ct, sg, 6, 1, iso1.MM, getattr(iso1, lev).minimal_level_string(),
                                         spectrum=g32[c, s, k].copy(), Pres=p, Temp=t)
                co.double_precision()
                d[ct] = co
            st.sets.append(d)
        lutS.sets[lev] = st
    lutS.PTcouples = PTs
    # no exact ties between two nodes: the reference picks the second node with an unstable
    # argsort (smm:1010, 1029, 1037), so a tie is not defined there
    probes = [(5.0e-5, 151.3), (1.0e-4, 147.4), (2.0e-3, 142.6), (0.06, 158.9), (0.0615, 152.6),
              (1.3, 139.2), (2.5, 161.0), (0.5, 150.0)]
    out['interp_g32'] = g32
    out['interp_PT'] = np.array(PTs)
    out['interp_probes'] = np.array(probes)
    res = np.zeros((len(probes), len(LEVELS), 3, n_g))
    for i, (p, t) in enumerate(probes):
        for s, lev in enumerate(iso1.levels):
            got = lutS.sets[lev].calculate(p, t)
            for k, ct in enumerate(CTYPES):
                res[i, s, k] = 0.0 if got[ct] is None else got[ct].spectrum
    out['interp_result'] = res
    try:
        lutS.sets['lev_01'].calculate(3.0, 150.0)
        out['interp_extrap_raises'] = np.array(False)
    except ValueError:
        out['interp_extrap_raises'] = np.array(True)

    temps = np.array([t for p, t in probes])
    press = np.array([p for p, t in probes])
    tv = np.array([[t, t + 7.5 + 3.0 * i, t - 4.0 + i] for i, t in enumerate(temps)]).T   # [lev][step]
    for s, lev in enumerate(iso1.levels):
        getattr(iso1, lev).local_vibtemp = list(tv[s])
    allLUTs = {(iso1.mol_name, iso1.iso): lutS}
    with R.quiet():
        a_n, e_n = smm.make_abscoeff_LUTS_fast(sg, iso1, temps, press, LTE=False, allLUTs=allLUTs,
                                               cartDROP=work)
        a_l, e_l = smm.make_abscoeff_LUTS_fast(sg, iso1, temps, press, LTE=True, allLUTs=allLUTs,
                                               cartDROP=work)
    out['abscoeff_tvib'] = tv
    out['abscoeff_nonlte'] = np.array([[a.spectrum for a in a_n.set], [e.spectrum for e in e_n.set]])
    out['abscoeff_lte'] = np.array([[a.spectrum for a in a_l.set], [e.spectrum for e in e_l.set]])
    out['iso_ratio'] = np.array(iso1.ratio)
    out['iso_MM'] = np.array([iso1.MM, iso2.MM])

    # ---- A8: calc_PT_couples_atmosphere on an analytic Titan-like atmosphere ----------------------
    z = np.arange(0.0, 1501.0, 10.0)
    T1 = np.where(z < 45, 94 - 24 * z / 45, np.where(z < 300, 70 + 110 * (z - 45) / 255,
                                                    180 - 20 * np.tanh((z - 300) / 400)))
    temp2 = np.array([T1 + d for d in (-8.0, 0.0, 6.5)])
    H = 20.0 + 0.05 * z
    pres2 = np.array([1467.0 * np.exp(-np.cumsum(np.r_[0, np.diff(z)] / H) * f) for f in (1.0, 0.97, 1.04)])
    atm = sbm.AtmProfile(sbm.AtmGrid(['lat', 'alt'], [[-90, -30, 30], z]), temp2, 'temp', ['box', 'lin'])
    atm.add_profile(pres2, 'pres', ['box', 'exp'])
    out['atm_z'], out['atm_temp'], out['atm_pres'] = z, temp2, pres2
    for tag, kw in (('a', dict(pres_step_log=1.0, temp_step=5.0, max_pres=2.0)),
                    ('b', dict(pres_step_log=0.4, temp_step=5.0, max_pres=0.1)),
                    ('c', dict(pres_step_log=1.0, temp_step=5.0, max_pres=2.5, add_lowpres=False))):
        out['ptc_' + tag] = np.array(smm.calc_PT_couples_atmosphere(lines, [iso1, iso2], atm, **kw))

    # ---- f1: instrument convolution (cm-1 and nm observation), FOV integration, masks ------------
    hg = smm.prepare_spe_grid([3000.0, 3030.0]).spectral_grid
    x = hg.grid
    spec = 1e-7 * (1 + 0.5 * np.sin(x * 3.1) + np.exp(-0.5 * ((x - 3012.3) / 0.01) ** 2) * 40
                   + (x - 3000) * 0.03)
    hires = spcl.SpectralIntensity(spec, hg, units='ergscm2')
    out['conv_grid'], out['conv_spec'] = x, spec
    ch = np.array([3001.0, 3004.2, 3010.0, 3012.3, 3019.9, 3027.0, 3029.5, 3040.0])
    wd = np.array([0.8, 1.1, 0.9, 1.3, 1.0, 0.7, 1.2, 1.0])
    obs = spcl.SpectralIntensity(np.zeros(len(ch)), spcl.SpectralGrid(ch, units='cm_1'), units='ergscm2')
    out['conv_cm_centres'], out['conv_cm_widths'] = ch, wd
    out['conv_cm_result'] = np.array(hires.hires_to_lowres(obs, spectral_widths=wd).spectrum)
    ch_nm = np.sort(1.e7 / ch[:-1])
    wd_nm = np.array([1.0, 0.9, 1.2, 0.8, 1.1, 1.0, 0.95])
    obs_nm = spcl.SpectralIntensity(np.zeros(len(ch_nm)), spcl.SpectralGrid(ch_nm, units='nm'), units='Wm2')
    out['conv_nm_centres'], out['conv_nm_widths'] = ch_nm, wd_nm
    low_nm = hires.hires_to_lowres(obs_nm, spectral_widths=wd_nm)
    out['conv_nm_result'] = np.array(low_nm.spectrum)
    out['conv_nm_units'] = np.array([low_nm.units, low_nm.spectral_grid.units])

    three = [spcl.SpectralIntensity(np.array(v), spcl.SpectralGrid(ch[:5], units='cm_1'))
             for v in ([1.0, 2.0, 0.5, 3.0, 1.5], [1.4, 2.2, 0.9, 2.0, 1.1], [2.0, 1.9, 1.6, 1.0, 0.7])]
    out['fov_in'] = np.array([t.spectrum for t in three])
    rots = [0.0, 12.0, -30.0, 44.0]
    out['fov_rot'] = np.array(rots)
    out['fov_out'] = np.array([smm.FOV_integr_1D(three, pixel_rot=r).spectrum for r in rots])

    zz = np.arange(0.0, 1501.0, 10.0)
    out['tri_z'] = zz
    out['tri_mid'] = smm.alt_triangle(zz, 550.0, node_lo=450.0, node_up=700.0).mask
    out['tri_first'] = smm.alt_triangle(zz, 350.0, node_up=450.0, first=True).mask
    out['tri_last'] = smm.alt_triangle(zz, 950.0, node_lo=850.0, last=True).mask
    out['tri_step'] = smm.alt_triangle(zz, 600.0, step=100.0).mask
    lims = [-90., -75., -60., -30., 30., 60., 75.]
    out['latbox_limits'] = np.array(lims)
    out['latbox_probe'] = np.array([-90.0, -76.0, -60.0, 0.0, 74.9, 75.0, 80.0])
    out['latbox'] = np.array([smm.lat_box(lims, la).mask for la in out['latbox_probe']])

    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, 'ref_golden.npz'), **out)
    shutil.rmtree(work, ignore_errors=True)
    print('wrote', os.path.join(HERE, 'ref_golden.npz'), 'with', len(out), 'arrays')


if __name__ == '__main__':
    main()
