"""Second set of reference-executed fixtures (tests/golden/ref_golden2.npz, ref_abscoeff_los.pic):
the host-side helpers either side of the hot path - SpectralObject arithmetic / slicing /
re-gridding / degraded grids, the older host convolution, black-body and shape helpers,
AbsSetLOS streams, make_abscoeff_LUTS_fast with tracked levels and the slow line-by-line
make_abscoeff_isomolec.  Run from the repo root in a container that has /root/reference:

    python tests/golden/make_ref_golden2.py

Every number comes out of the reference's OWN Python executed by ref_exec.py (see its header for
the stand-ins); the tests that read these files do not need /root/reference.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_exec as R  # noqa: E402
import make_ref_golden as M  # noqa: E402

CTYPES = M.CTYPES


class Box(object):
    """Stands for a multiprocessing.Queue: keeps what is put."""

    def put(self, item):
        self.item = item


def test_spectrum(spcl):
    """A hi-res spectrum with a few narrow features on a smooth background, regular cm-1 grid."""
    x = np.arange(3000.0, 3010.0 + 0.005, 0.01)
    y = 1e-6 * (1.0 + 0.2 * np.sin(1.7 * x))
    for c, w, a in ((3001.3, 0.03, 5.0), (3004.71, 0.05, 1.0), (3004.9, 0.02, 0.3), (3008.2, 0.08, 2.5)):
        y = y + a * w ** 2 / ((x - c) ** 2 + w ** 2)
    return spcl.SpectralIntensity(y, spcl.SpectralGrid(x, units='cm_1'), units='ergscm2')


def main():
    spcl, smm, sbm = R.load()
    out = dict()
    work = tempfile.mkdtemp() + '/'
    cwd = os.getcwd()
    os.chdir(work)

    # ---- SpectralObject: slicing, arithmetic, re-gridding --------------------------------------
    spe = test_spectrum(spcl)
    out['so_grid'], out['so_spec'] = spe.spectral_grid.grid, spe.spectrum
    cut = spe[3002.0, 3003.5]
    out['so_cut_grid'], out['so_cut_spec'] = cut.spectral_grid.grid, cut.spectrum
    out['so_cut_none'] = np.array(spe[3020.0, 3021.0] is None)
    part = spe[3004.0, 3006.0]
    part.spectrum = part.spectrum * 0.5
    out['so_add_short'] = (spe + part).spectrum
    out['so_sub_short'] = (spe - part).spectrum
    out['so_add_scalar'] = (spe + 0.25).spectrum
    out['so_mul'] = (spe * spe).spectrum
    a = spcl.SpectralObject(spe.spectrum.copy(), spe.spectral_grid)
    a.add_to_spectrum(part, Strength=-2.0)
    out['so_add_to_spectrum'] = a.spectrum
    a = spcl.SpectralObject(spe.spectrum.copy(), spe.spectral_grid)
    a.add_to_spectrum_slow(part, Strength=3.0)
    out['so_add_to_spectrum_slow'] = a.spectrum
    out['so_exp'] = spe.exp_elementwise(-0.7).spectrum
    out['so_mulel'] = spe.multiply_elementwise(spe, save=False).spectrum
    out['so_divel'] = spe.divide_elementwise(spcl.SpectralObject(spe.spectrum + 1.0, spe.spectral_grid),
                                             save=False).spectrum
    a = spcl.SpectralObject(spe.spectrum.copy(), spe.spectral_grid)
    out['so_sum_scalar'] = np.array(a.sum_scalar(1.5))
    nug = spcl.SpectralGrid(np.arange(2999.5, 3010.6, 0.0137), units='cm_1')
    out['so_interp_grid'] = nug.grid
    out['so_interp'] = spe.interp_to_grid(nug).spectrum
    irr = spe.interp_to_grid(spcl.SpectralGrid(np.sort(np.random.default_rng(3).uniform(3000, 3010, 300)),
                                               units='cm_1'))
    out['so_irr_grid'], out['so_irr_spec'] = irr.spectral_grid.grid.copy(), irr.spectrum.copy()
    irr.interp_to_regular_grid()
    out['so_reg_grid'], out['so_reg_spec'] = irr.spectral_grid.grid, irr.spectrum

    with R.quiet():
        d1 = spe.degrade_grid()
        d1b = spe.degrade_grid(thress=[3e-2, 1e-2, 1e-3], factors=[4, 10, 50], consider_derivatives=False)
        d2 = spe.degrade_grid2()
        d2b = spe.degrade_grid2(thres=0.2, num_aside=[3, 2, 2], res_low=[1, 2, 5], consider_derivatives=False)
    for tag, d in (('d1', d1), ('d1b', d1b), ('d2', d2), ('d2b', d2b)):
        out['deg_%s_grid' % tag], out['deg_%s_spec' % tag] = d.spectral_grid.grid, d.spectrum
    weak = spcl.SpectralObject(spe.spectrum * 1e-3, spe.spectral_grid)
    with R.quiet():
        out['best_grid_1'] = smm.best_compressed_grid([{'a': spe, 'b': weak}], alg=1)
        out['best_grid_2'] = smm.best_compressed_grid([{'a': spe}, {'b': weak}], alg=2)

    # ---- the older host convolution (regular grid), hires_to_lowres_old, tolowres ----------------
    ch = np.array([2999.2, 3000.4, 3002.0, 3004.8, 3008.2, 3009.8, 3011.0, 3015.0])
    wd = np.array([0.30, 0.25, 0.40, 0.20, 0.35, 0.30, 0.30, 0.30])
    obs = spcl.SpectralIntensity(np.zeros(len(ch)), spcl.SpectralGrid(ch, units='cm_1'), units='nWcm2')
    out['cv_centres'], out['cv_widths'] = ch, wd
    out['cv_result'] = spe.convolve_to_grid(obs.spectral_grid, spectral_widths=list(wd)).spectrum
    reg = spcl.SpectralGrid(np.arange(3001.0, 3009.01, 0.5), units='cm_1')
    out['cv_reg_grid'] = reg.grid
    out['cv_reg_result'] = spe.convolve_to_grid(reg).spectrum
    old = test_spectrum(spcl)
    low = old.hires_to_lowres_old(obs, spectral_widths=list(wd))
    out['cv_old_result'], out['cv_old_units'] = low.spectrum, np.array([low.units, low.spectral_grid.units])
    ch_nm = np.sort(1.e7 / ch[1:6])
    obs_nm = spcl.SpectralIntensity(np.zeros(len(ch_nm)), spcl.SpectralGrid(ch_nm, units='nm'), units='Wm2')
    obs_nm.add_bands(spcl.SpectralObject(np.array([0.30, 0.22, 0.35, 0.28, 0.31]), obs_nm.spectral_grid))
    out['tl_centres'], out['tl_widths'] = ch_nm, obs_nm.bands.spectrum
    low = smm.tolowres(test_spectrum(spcl), obs_nm)
    out['tl_result'], out['tl_units'] = low.spectrum, np.array([low.units, low.spectral_grid.units])

    # ---- prepare_fortran_sum: clipping and padding of line windows -------------------------------
    big = spcl.SpectralObject(np.zeros(len(spe.spectrum)), spe.spectral_grid)
    wins = []
    for c0, n in ((3000.02, 41), (3005.0, 41), (3009.97, 41), (3004.0, 64)):
        i0 = int(np.argmin(np.abs(big.spectral_grid.grid - c0)))
        g = big.spectral_grid.grid[i0] + 0.01 * (np.arange(n) - n // 2)
        wins.append(spcl.SpectralObject(np.linspace(1.0, 2.0, n) * (1 + len(wins)),
                                        spcl.SpectralGrid(g, units='cm_1')))
    box = Box()
    big.prepare_fortran_sum(wins, 0, box, fix_length=64)
    out['pfs_win_grid0'] = np.array([w.spectral_grid.grid[0] for w in wins])
    out['pfs_win_len'] = np.array([len(w.spectrum) for w in wins])
    out['pfs_matrix'], out['pfs_init'], out['pfs_fin'] = [np.array(v) for v in box.item]

    # ---- scalar / shape helpers -------------------------------------------------------------------
    sg = spcl.SpectralGrid(np.arange(3000.0, 3001.0, 0.01), units='cm_1')
    probes = np.array([3000.0, 3000.504, 3000.99, 2999.93, 2998.71, 3001.3, 3003.456])
    out['cge_probes'] = probes
    out['cge_result'] = np.array([spcl.closest_grid_ext(sg, w) for w in probes], dtype=float)
    xs = np.linspace(2999.0, 3001.0, 41)
    out['shape_x'] = xs
    out['shape_lorentz'] = spcl.Lorentz_shape(xs, 3000.1, 0.07)
    out['shape_doppler'] = spcl.Doppler_shape(xs, 3000.1, 0.004)
    lg = spcl.SpectralGrid(np.arange(-200, 201) * 5e-4 + 3000.0, units='cm_1')
    out['shape_py_grid'] = lg.grid
    out['shape_py'] = spcl.MakeShape_py(lg, 3000.0, 0.004, 0.0045, Strength=2.0).spectrum
    out['hit_strength'] = np.array(spcl.Einstein_A_to_LineStrength_hitran(12.3, 3012.5, 180.0, 420.0, 15.0,
                                                                          312.7, iso_ab=0.988))
    out['boltz_pop'] = np.array(spcl.Boltz_pop_at_T(1533.3, 170.0, 3.0, 240.0))
    out['alpha_nlte'] = np.array(spcl.alpha_nlte(3019.5, 160.0, 1.3, 25.0))
    bg = spcl.SpectralGrid(np.linspace(2900.0, 3100.0, 11), units='cm_1')
    out['bb_grid'] = bg.grid
    out['bb_erg'] = spcl.Calc_BB(bg, 180.0).spectrum
    out['bb_wm2'] = spcl.Calc_BB(bg, 180.0, units='Wm2').spectrum
    out['bb_single'] = np.array(spcl.Calc_BB_single(3019.5, 94.0))
    out['bb_fun'] = np.array([spcl.BB(180.0, 3019.5), spcl.BB_erg(180.0, 3019.5), spcl.BB_nm(180.0, 3311.0),
                              spcl.BB_nm(5800.0, 500.0)])
    out['cm1_to_J'] = np.array(spcl.convert_cm_1_to_J(3019.5))

    lines = spcl.read_line_database(os.path.join(HERE, 'ref_lines.par'))
    dec = (lambda v: v.decode() if isinstance(v, bytes) else v)
    for l in lines:
        for k in ('Up_lev_str', 'Lo_lev_str', 'Q_num_up', 'Q_num_lo', 'others'):
            setattr(l, k, dec(getattr(l, k)))
        l.Mol, l.Iso = int(l.Mol), int(l.Iso)
    iso1, iso2 = M.case_isomolecs(sbm)
    out['print_hitran'] = np.array([lines[0].Print_hitran(ofile=open(os.devnull, 'w')),
                                    lines[7].Print_hitran(ofile=open(os.devnull, 'w'))])
    sfs = []
    for lin in lines:
        if lin.Iso == 1 and lin.LinkToMolec(iso1):
            sfs.append(lin.CalcStrength_from_Strength(165.0, T_vib_lower=165.0, T_vib_upper=190.0))
    out['strength_from_strength'] = np.array(sfs)

    out['equiv'] = np.array([smm.equiv(0, 0), smm.equiv(0, 1e-20), smm.equiv(1.0, 1.0 + 5e-9),
                             smm.equiv(1.0, 1.0 + 5e-8), smm.equiv(-2.0, -2.0), smm.equiv(3.0, 3.1, thres=0.1)])
    open(work + 'name.pic', 'w').close()
    open(work + 'name_001.pic', 'w').close()
    out['free_name'] = np.array([os.path.basename(smm.find_free_name(work + 'name.pic')),
                                 os.path.basename(smm.find_free_name(work + 'other.pic')),
                                 os.path.basename(smm.find_free_name(work + 'name.pic', maxnum=50))])

    # ---- AbsSetLOS stream written by the reference ------------------------------------------------
    asg = spcl.SpectralGrid(np.arange(16) * 5e-4 + 3000.0, units='cm_1')
    rows = np.random.default_rng(11).uniform(0, 1, (3, 16))
    st = smm.AbsSetLOS(work + 'abs.pic', spectral_grid=asg)
    st.prepare_export()
    for r in rows:
        st.add_dump(spcl.SpectralObject(r, asg, link_grid=True))
    st.finalize_IO()
    shutil.copy(work + 'abs.pic', os.path.join(HERE, 'ref_abscoeff_los.pic'))
    out['absset_grid'], out['absset_rows'] = asg.grid, rows

    # ---- A11 with tracked levels: make_abscoeff_LUTS_fast(track_levels=...) on the synthetic LUT ---
    ref1 = np.load(os.path.join(HERE, 'ref_golden.npz'))
    g32, PTs, probes = ref1['interp_g32'], ref1['interp_PT'], ref1['interp_probes']
    tsg = spcl.SpectralGrid(np.arange(g32.shape[3]) * 5e-4 + 3000.0, units='cm_1')
    lutS = smm.LookUpTable(iso1, tsg.wn_range(), False)
    for s, lev in enumerate(iso1.levels):
        ls = smm.LutSet(6, 1, iso1.MM, level=getattr(iso1, lev))
        ls.PTcouples = [list(pt) for pt in PTs]
        ls.spectral_grid = tsg
        for c, (p, t) in enumerate(PTs):
            d = dict()
            for k, ct in enumerate(CTYPES):
                if not np.any(g32[c, s, k]):
                    d[ct] = None
                    continue
                co = spcl.SpectralGcoeff(ct, tsg, 6, 1, iso1.MM, getattr(iso1, lev).minimal_level_string(),
                                         spectrum=g32[c, s, k].copy(), Pres=p, Temp=t)
                co.double_precision()
                d[ct] = co
            ls.sets.append(d)
        lutS.sets[lev] = ls
    lutS.PTcouples = [list(pt) for pt in PTs]
    temps, press, tv = probes[:, 1], probes[:, 0], ref1['abscoeff_tvib']
    for s, lev in enumerate(iso1.levels):
        getattr(iso1, lev).local_vibtemp = list(tv[s])
    track = ['lev_01', 'lev_02']
    with R.quiet():
        a, e, et, at = smm.make_abscoeff_LUTS_fast(tsg, iso1, temps, press, LTE=False,
                                                   allLUTs={(iso1.mol_name, iso1.iso): lutS},
                                                   cartDROP=work, track_levels=track)
    assert np.array_equal(np.array([x.spectrum for x in a.set]), ref1['abscoeff_nonlte'][0])
    out['track_levels'] = np.array(track)
    out['track_emi'] = np.array([[x.spectrum for x in et[lev].set] for lev in track])
    out['track_abs'] = np.array([[x.spectrum for x in at[lev].set] for lev in track])
    out['track_names'] = np.array([os.path.basename(a.filename), os.path.basename(e.filename),
                                   os.path.basename(et['lev_01'].filename),
                                   os.path.basename(at['lev_02'].filename)])

    # ---- slow line-by-line twin: make_abscoeff_isomolec(useLUTs=False) ----------------------------
    steps_T, steps_P = [152.0, 171.5, 160.0], [0.02, 1.1, 0.02]
    tv3 = np.array([[152.0, 171.5, 160.0], [185.0, 176.0, 199.0], [150.0, 172.5, 164.0]])
    for s, lev in enumerate(iso1.levels):
        getattr(iso1, lev).local_vibtemp = list(tv3[s])
    with R.quiet():
        a, e, et, at = smm.make_abscoeff_isomolec(M.WN_RANGE, iso1, steps_T, steps_P, LTE=False,
                                                  useLUTs=False, lines=lines, cartDROP=work,
                                                  track_levels=['lev_01'], n_threads=2)
        a2, e2 = smm.make_abscoeff_isomolec(M.WN_RANGE, iso2, steps_T[:2], steps_P[:2], LTE=True,
                                            useLUTs=False, lines=lines, cartDROP=work, n_threads=2)
    out['slow_T'], out['slow_P'], out['slow_tvib'] = np.array(steps_T), np.array(steps_P), tv3
    out['slow_abs'] = np.array([x.spectrum for x in a.set])
    out['slow_emi'] = np.array([x.spectrum for x in e.set])
    out['slow_track_emi'] = np.array([x.spectrum for x in et['lev_01'].set])
    out['slow_abs_iso2'] = np.array([x.spectrum for x in a2.set])
    out['slow_emi_iso2'] = np.array([x.spectrum for x in e2.set])

    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, 'ref_golden2.npz'), **out)
    shutil.rmtree(work, ignore_errors=True)
    print('wrote', os.path.join(HERE, 'ref_golden2.npz'), 'with', len(out), 'arrays')


if __name__ == '__main__':
    main()
