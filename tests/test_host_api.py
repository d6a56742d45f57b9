"""CPU tests of the host-side mirror of the reference API (spect_classes / spect_main_module /
spect_base_module): everything here is index logic and scalar physics, no GPU."""
import math as mt
import os

import numpy as np
import pytest

from spectrobot_b200 import spect_base_module as sbm
from spectrobot_b200 import spect_classes as spcl
from spectrobot_b200 import spect_main_module as smm
from spectrobot_b200 import synthetic as S


@pytest.fixture(scope="module")
def world():
    tab = S.line_table(120, 2995.0, 3005.0, n_levels=5, seed=11, frac_unlinked=0.1)
    return dict(tab=tab, lines=S.spect_lines(tab), planet=S.titan_planet(tab["level_energies"]))


def test_level_linking_matches_integer_ids(world):
    """LinkToMolec (spect_classes.py:122-150) by strings == the integer set ids of line_table."""
    im = world["planet"].gases['CH4'].iso_1
    tab = world["tab"]
    t2 = spcl.line_table(world["lines"], im)
    assert np.array_equal(t2["up_set"], tab["up_set"]) and np.array_equal(t2["lo_set"], tab["lo_set"])
    for i, lin in enumerate(world["lines"]):
        ok = lin.LinkToMolec(im)
        assert ok == (tab["up_set"][i] >= 0)
        if ok:
            assert lin.Up_lev_id == 'lev_%02d' % tab["up_set"][i]
            assert lin.E_vib_lo == tab["level_energies"][tab["lo_set"][i]]


def test_line_physics_against_oracle(world, oracle):
    im = world["planet"].gases['CH4'].iso_1
    for lin in world["lines"][:25]:
        dw, lw, sh = lin.CheckWidths(163.0, 0.3, im.MM)
        assert dw == pytest.approx(oracle.Doppler_width(163.0, im.MM, lin.Freq), rel=1e-15)
        assert lw == pytest.approx(oracle.Lorenz_width(163.0, 0.3 * spcl.hpa_to_atm,
                                                       lin.T_dep_broad, lin.Air_broad), rel=1e-15)
        G = lin.Calc_Gcoeffs(163.0, isomolec=im)
        if lin.Up_lev_id is None:
            continue
        ref = oracle.Calc_Gcoeffs(lin.Freq, lin.A_coeff, lin.E_lower, lin.g_up, lin.g_lo,
                                  lin.E_vib_up, lin.E_vib_lo, 163.0)
        for j, k in enumerate(spcl.CTYPES):
            assert G[k] == pytest.approx(ref[j], rel=1e-14)


def test_A_from_strength_inverts_strength_from_Einstein(world):
    """calc_A_coeff_from_strength is the inverse of the LTE strength at
    296 K (spect_classes.py:291-309 vs :219-254) -- with the partition sum passed explicitly so
    that no table is needed on the CPU."""
    lin = world["lines"][0]
    q = 590.52
    a = lin.calc_A_coeff_from_strength(iso_ab=S.CH4_RATIO, Q_part=q)
    assert a == pytest.approx(lin.A_coeff, rel=1e-12)
    lin.LinkToMolec(None)
    S_ab, _ = lin.CalcStrength_from_Einstein(296.0, Q_part=q, iso_ab=S.CH4_RATIO)
    assert S_ab == pytest.approx(lin.Strength, rel=1e-10)


def test_PT_couples_ladder(world):
    """calc_PT_couples_atmosphere (spect_main_module.py:1746-1844)."""
    planet, lines = world["planet"], world["lines"]
    gases = list(planet.gases.values())
    PT = smm.calc_PT_couples_atmosphere(lines, gases, planet.atmosphere, pres_step_log=1.0,
                                        temp_step=5.0, max_pres=2.5)
    PT = np.array(PT)
    assert 30 < len(PT) < 600
    # temperatures on the 5 K ladder, pressures on multiples of the log step
    assert np.allclose(PT[:, 1] / 5.0, np.round(PT[:, 1] / 5.0))
    n = np.log(PT[:, 0])
    assert np.allclose(n, np.round(n), atol=1e-9)
    assert PT[:, 0].max() == pytest.approx(mt.exp(mt.ceil(mt.log(2.5))))
    assert PT[:, 0].min() == pytest.approx(mt.exp(mt.floor(mt.log(planet.atmosphere.pres.min()))))
    # Doppler-dominated cells collapse onto exactly two pressures with identical temperature sets
    broad = lines[int(np.argmax([l.Air_broad for l in lines]))]
    dop = [pt for pt in PT if broad.CheckWidths(pt[1], pt[0], S.CH4_MM)[1]
           < 0.01 * broad.CheckWidths(pt[1], pt[0], S.CH4_MM)[0]]
    dop_p = np.unique([p for p, _ in dop])
    assert len(dop_p) == 2
    t_a = sorted(t for p, t in dop if p == dop_p[0])
    t_b = sorted(t for p, t in dop if p == dop_p[1])
    assert t_a == t_b
    # every atmospheric level below max_pres is bracketed in T at its pressure level
    atm_p, atm_t = planet.atmosphere.pres, planet.atmosphere.temp
    for p, t in zip(atm_p, atm_t):
        if p > 2.5:
            continue
        near = PT[np.argmin(np.abs(np.log(PT[:, 0]) - np.log(max(p, PT[:, 0].min()))))][0]
        ts = PT[PT[:, 0] == near][:, 1]
        assert ts.min() <= t <= ts.max()
    # no collapse when thres = 0
    PT0 = smm.calc_PT_couples_atmosphere(lines, gases, planet.atmosphere, pres_step_log=1.0,
                                         temp_step=5.0, max_pres=2.5, thres=0.0)
    assert len(PT0) > len(PT)


def test_lutset_calculate_follows_reference_rule(oracle):
    """LutSet.calculate (spect_main_module.py:997-1066) on a host-resident table against the
    literal restatement in the oracle, incl. the lowest-pressure branch, None propagation and
    'Extrapolating in P'."""
    import torch
    rng = np.random.default_rng(5)
    im = sbm.IsoMolec(6, 1, LTE=False)
    im.add_levels(S.level_strings(2), [0.0, 1310.76])
    grid = spcl.SpectralGrid(np.linspace(3000.0, 3000.1, 41), units='cm_1')
    PT = [[p, float(t)] for p in (0.01, 0.1, 1.0) for t in (150., 155., 160., 165.)]
    lut = smm.LookUpTable(im, [3000.0, 3000.1], LTE=False)
    lut.PTcouples, lut.spectral_grid = PT, grid
    g = rng.uniform(0.5, 1.5, (len(PT), 2, 3, 41)).astype(np.float32)
    g[:, 1, 0] = 0.0                                   # an all-zero spectrum: None in the reference
    lut.g32 = torch.as_tensor(g)
    for s, nam in enumerate(im.levels):
        st = smm.LutSet(6, 1, im.MM, level=getattr(im, nam))
        st.PTcouples, st.spectral_grid, st._table = PT, grid, (lut, s)
        lut.sets[nam] = st
    for P, T in [(0.05, 157.0), (0.3, 163.9), (0.004, 152.0), (0.01, 161.0), (1.0, 150.0)]:
        for s, nam in enumerate(im.levels):
            got = lut.sets[nam].calculate(P, T)
            for k, ct in enumerate(spcl.CTYPES):
                sets = [None if not np.any(g[c, s, k]) else g[c, s, k].astype(float)
                        for c in range(len(PT))]
                ref = oracle.LutSet_calculate(PT, sets, P, T)
                if ref is None:
                    assert got[ct] is None
                else:
                    assert np.array_equal(got[ct].spectrum, ref)
                    assert got[ct].temp == T
    with pytest.raises(ValueError, match='Extrapolating in P'):
        lut.sets['lev_00'].calculate(2.0, 155.0)
    ok, lev = lut.find_lev(S.level_strings(2)[1])
    assert ok and lev == 'lev_01'


def test_gcoeff_interpolate_checks(world):
    grid = spcl.SpectralGrid(np.linspace(1.0, 2.0, 5), units='cm_1')
    a = spcl.SpectralGcoeff('absorption', grid, 6, 1, 16.0, '0 0 0 0', spectrum=np.ones(5),
                            Pres=0.1, Temp=150.0)
    b = spcl.SpectralGcoeff('absorption', grid, 6, 1, 16.0, '0 0 0 0', spectrum=3 * np.ones(5),
                            Pres=0.1, Temp=160.0)
    c = a.interpolate(b, Temp=152.5)
    assert np.allclose(c.spectrum, 1.5) and c.temp == 152.5 and c.pres == 0.1
    with pytest.raises(ValueError):
        a.interpolate(b, Pres=0.2)                     # same P, different T: wrong keyword
    assert a.interpolate(None, Temp=151.0) is None
    assert sbm.weight(152.5, 150.0, 160.0) == (0.75, 0.25)


def test_los_geometry():
    planet = S.titan_planet(None, nonlte=False)
    pix = S.vims_pixels([420.0, 900.0])
    for p, ht in zip(pix, (420.0, 900.0)):
        assert p.LOS().get_tangent_altitude() == pytest.approx(ht, abs=0.05)
        assert p.low_LOS().get_tangent_altitude() == pytest.approx(ht - 12.0, abs=0.05)
        los = p.LOS()
        pts = los.calc_atm_intersections(planet, delta_x=5.0)
        alts = np.array([q.Spherical()[2] for q in pts])
        assert alts[0] == pytest.approx(1500.0, abs=1e-6) and alts[-1] == pytest.approx(1500.0, abs=1e-6)
        assert alts.min() == pytest.approx(los.get_tangent_altitude(), abs=1e-6)
        assert np.all(np.diff(los._s) < 0)             # far end -> observer
        assert np.all(np.abs(np.diff(alts)) > 0)       # no two neighbours at the same altitude
    high = S.vims_pixels([1600.0])[0].LOS()
    assert high.calc_atm_intersections(planet) == []


def test_prepare_spe_grid_is_numpy_arange():
    g = smm.prepare_spe_grid([2850.0, 3450.0]).spectral_grid.grid
    assert len(g) == 1200001
    assert np.array_equal(g, np.arange(2850.0, 3450.0 + 2.5e-4, 5e-4))
    assert g[-1] != 2850.0 + 1200000 * 5e-4            # the grid drifts


def test_fov_integration_closed_form_matches_spline_quad(oracle):
    """fov_integrate (closed form) against the literal RectBivariateSpline + quad restatement of
    FOV_integr_1D (spect_main_module.py:3342-3374), incl. a rotated pixel."""
    rng = np.random.default_rng(3)
    grid = np.linspace(2900.0, 3400.0, 9)
    spectra = rng.uniform(0.5, 2.0, (3, 9)) * np.array([[1.0], [1.3], [0.8]])
    for rot in (0.0, 12.5, -30.0, 44.0):
        got = smm.fov_integrate(spectra, rot)
        ref = oracle.FOV_integr_1D(spectra, grid, rot)
        assert np.allclose(got, ref, rtol=2e-7, atol=0), rot
    # a flat field integrates to itself (unit-area response), any rotation
    flat = np.ones((3, 4)) * 2.5
    assert np.allclose(smm.fov_integrate(flat, 17.0), 2.5, rtol=1e-14)
    sg = spcl.SpectralGrid(grid, units='cm_1')
    rads = [spcl.SpectralIntensity(spectra[i], sg) for i in range(3)]
    out = smm.FOV_integr_1D(rads, 12.5)
    assert np.array_equal(out.spectrum, smm.fov_integrate(spectra, 12.5))


def test_read_line_database_hitran_and_gbb(tmp_path, world):
    """read_line_database (spect_classes.py:1532-1601) on files written in the HITRAN2012
    160-character and the 'gbb' layouts: field round trip, selection rules, default widths."""
    lines = world["lines"][:40]
    lines[3].Air_broad = 0.0                       # -> 0.05 on reading (:1578-1581)
    for fmt in ('HITRAN', 'gbb'):
        fn = tmp_path / ('db_%s.par' % fmt)
        with open(fn, 'w') as f:
            f.write("header to skip\n")
            for lin in lines:
                f.write(spcl.format_line_record(lin, fmt) + "\n")
        if fmt == 'HITRAN':
            assert all(len(l.rstrip("\n")) == 160 for l in open(fn).readlines()[1:])
        got = spcl.read_line_database(str(fn), db_format=fmt, n_skip=1)
        assert len(got) == len(lines)
        for a, b in zip(got, lines):
            assert a.Mol == b.Mol and a.Iso == b.Iso
            assert a.Freq == pytest.approx(b.Freq, abs=6e-7)
            assert a.Strength == pytest.approx(b.Strength, rel=6e-4)
            assert a.A_coeff == pytest.approx(b.A_coeff, rel=6e-4)
            assert a.E_lower == pytest.approx(b.E_lower, abs=6e-5)
            assert a.T_dep_broad == pytest.approx(b.T_dep_broad, abs=6e-3)
            assert a.Up_lev_str.strip() == b.Up_lev_str.strip()
            assert a.Self_broad == 0.07
            if fmt == 'HITRAN':
                assert a.g_up == b.g_up and a.g_lo == b.g_lo
        assert got[3].Air_broad == 0.05
        # frequency window on the sorted file, molecule filter, strength cut
        f0, f1 = lines[10].Freq - 1e-4, lines[20].Freq + 1e-4
        sub = spcl.read_line_database(str(fn), db_format=fmt, n_skip=1, freq_range=[f0, f1])
        assert len(sub) == 11
        assert spcl.read_line_database(str(fn), mol=5, db_format=fmt, n_skip=1) == []
        kept = spcl.read_line_database(str(fn), db_format=fmt, n_skip=1, fraction_to_keep=0.5)
        assert 0 < len(kept) < len(lines)
        thr = np.sort([g.Strength for g in got])[int(0.5 * (len(got) - 1))]     # (:1592-1594)
        assert min(k.Strength for k in kept) == thr and len(kept) == sum(g.Strength >= thr for g in got)
    # level linking while reading
    im = world["planet"].gases['CH4'].iso_1
    fn = tmp_path / 'db_HITRAN.par'
    linked = spcl.read_line_database(str(fn), db_format='HITRAN', n_skip=1, link_to_isomolecs=[im])
    assert any(l.Up_lev_id is not None for l in linked)
    with pytest.raises(ValueError):
        spcl.read_line_database(str(fn), db_format='GEISA')


def test_stat_weights():
    assert spcl.calc_stat_weights_CH4('    7F2 21', '    6F1  3') == (3 * 15, 3 * 13)
    assert spcl.calc_stat_weights_CH4('  12 A1', '  11 A2', formato='hot_bands') == (5 * 25, 5 * 23)
    assert spcl.calc_stat_weights_linear_molec(6, 1, '', ' P 10') == (6 * 19, 6 * 21)
    assert spcl.calc_stat_weights_linear_molec(6, 1, '', ' R 10e') == (6 * 23, 6 * 21)
    assert spcl.calc_stat_weights_linear_molec(1, 1, '', ' Q  4') == (9, 9)
    assert spcl.calc_stat_weights_linear_molec(6, 1, ' 11', ' 10e', formato='GEISA') == (6 * 23, 6 * 21)


def test_lut_level_streams_round_trip_and_reference_module_names(tmp_path):
    """export_levels writes the reference's per-level pickle stream (PTcouples header + one
    {ctype: SpectralGcoeff} per cell, smm:880-892, 1122-1161); read_lutset_stream reads it back,
    also when the classes were pickled under the reference's top-level module names."""
    import pickle
    import sys
    import types
    import torch
    rng = np.random.default_rng(9)
    im = sbm.IsoMolec(6, 1, LTE=False)
    im.add_levels(S.level_strings(2), [0.0, 1310.76])
    grid = spcl.SpectralGrid(np.linspace(3000.0, 3000.1, 33), units='cm_1')
    PT = [[0.01, 150.0], [0.01, 155.0], [0.1, 150.0]]
    lut = smm.LookUpTable(im, [3000.0, 3000.1], LTE=False)
    lut.PTcouples, lut.spectral_grid = PT, grid
    g = rng.uniform(0.5, 1.5, (3, 2, 3, 33)).astype(np.float32)
    lut.g32 = torch.as_tensor(g)
    for s, nam in enumerate(im.levels):
        st = smm.LutSet(6, 1, im.MM, level=getattr(im, nam))
        st.PTcouples, st.spectral_grid, st._table = PT, grid, (lut, s)
        lut.sets[nam] = st
    files = lut.export_levels(str(tmp_path), stamp='_test')
    assert sorted(files) == sorted(im.levels)
    for s, nam in enumerate(im.levels):
        assert files[nam].endswith('LUT_mol06_iso1_nonLTE_%s_test.pic' % nam)
        pts, sets = smm.read_lutset_stream(files[nam])
        assert pts == PT and len(sets) == 3
        for c in range(3):
            for k, ct in enumerate(spcl.CTYPES):
                gc = sets[c][ct]
                assert gc.spectral_grid is None and gc.ctype == ct           # erase_grid (:1143)
                assert gc.temp == PT[c][1] and gc.pres == PT[c][0]
                assert np.array_equal(np.asarray(gc.spectrum, dtype=np.float32), g[c, s, k])
    # a truncated stream (interrupted build) yields the complete cells only
    raw = open(files[im.levels[0]], 'rb').read()
    cut = tmp_path / 'cut.pic'
    cut.write_bytes(raw[:int(len(raw) * 0.6)])
    pts, sets = smm.read_lutset_stream(str(cut))
    assert 0 < len(sets) < 3 and pts == PT[:len(sets)]
    # classes pickled under the reference's module name `spect_classes`
    fake = types.ModuleType('spect_classes')

    class SpectralGcoeff(object):
        pass
    SpectralGcoeff.__module__ = 'spect_classes'
    SpectralGcoeff.__qualname__ = 'SpectralGcoeff'
    fake.SpectralGcoeff = SpectralGcoeff
    sys.modules['spect_classes'] = fake
    try:
        obj = SpectralGcoeff()
        obj.spectrum, obj.ctype, obj.temp, obj.pres = np.arange(33.0), 'absorption', 150.0, 0.01
        ref_file = tmp_path / 'ref.pic'
        with open(ref_file, 'wb') as f:
            pickle.dump([[0.01, 150.0]], f, protocol=2)
            pickle.dump({'absorption': obj, 'sp_emission': obj, 'ind_emission': obj}, f, protocol=2)
    finally:
        del sys.modules['spect_classes']
    pts, sets = smm.read_lutset_stream(str(ref_file))
    assert pts == [[0.01, 150.0]]
    assert isinstance(sets[0]['absorption'], spcl.SpectralGcoeff)
    assert np.array_equal(sets[0]['absorption'].spectrum, np.arange(33.0))


def test_alt_triangle_matches_reference_loop():
    """Vectorised alt_triangle against the literal per-point loop of smm:319-352."""
    from spectrobot_b200 import spect_main_module as smm
    z = np.arange(0.0, 1501.0, 10.0)

    def ref(alt_grid, node_alt, node_lo=None, node_up=None, first=False, last=False):
        cos = np.zeros(len(alt_grid))
        for ii, alt in enumerate(alt_grid):
            if first:
                cos[ii] = 1.0 if alt < node_alt else (1.0 - (alt - node_alt) / (node_up - node_alt) if alt < node_up else 0.0)
            elif last:
                cos[ii] = 1.0 if alt > node_alt else (1.0 - (node_alt - alt) / (node_alt - node_lo) if alt > node_lo else 0.0)
            elif alt < node_lo or alt > node_up:
                cos[ii] = 0.0
            elif alt >= node_alt:
                cos[ii] = 1.0 - (alt - node_alt) / (node_up - node_alt)
            else:
                cos[ii] = 1.0 - (node_alt - alt) / (node_alt - node_lo)
        return cos
    assert np.array_equal(smm.alt_triangle(z, 300., node_up=500., first=True).mask, ref(z, 300., node_up=500., first=True))
    assert np.array_equal(smm.alt_triangle(z, 900., node_lo=700., last=True).mask, ref(z, 900., node_lo=700., last=True))
    assert np.array_equal(smm.alt_triangle(z, 505., node_lo=300., node_up=700.).mask, ref(z, 505., 300., 700.))
    assert np.array_equal(smm.alt_triangle(z, 500., step=200.).mask, ref(z, 500., 300., 700.))


def test_retrieval_parameter_space_host_logic():
    """RetParam / LinearProfile_1D_new / BayesSet (smm:161-296, 442-489, 600-656): the triangle
    masks form a partition of unity between the first and last node, profile() reproduces a
    profile that is linear between the nodes, update_par keeps positive parameters positive and
    build_jacobian stacks the stored per-pixel derivative spectra column by column."""
    z = np.arange(0.0, 1501.0, 10.0)
    nodes = [200., 450., 700., 1100.]
    vals = [1.0e-2, 1.4e-2, 0.9e-2, 2.0e-2]
    prof = smm.LinearProfile_1D_new('CH4', z, nodes, vals, [1e-3] * 4)
    tot = sum(p.maskgrid.mask for p in prof.set)
    assert np.allclose(tot, 1.0, atol=1e-15)
    got = prof.profile().values['vmr']
    assert np.allclose(got, np.interp(z, nodes, vals), rtol=1e-14)
    assert prof.profile().values['CH4'] is not None and prof.keys() == nodes
    assert prof.check_involved(200., dict(alt=[500., 900.])) is False      # LOS entirely above node 2
    assert prof.check_involved(450., dict(alt=[500., 900.])) is True
    bs = smm.BayesSet('t')
    bs.add_set(prof)
    assert bs.n_tot == 4 and [p.key for p in bs.params()] == nodes
    assert np.array_equal(bs.VCM_apriori(), np.diag([1e-6] * 4))
    p0 = bs.params()[0]
    p0.update_par(-5.0e-2)                                                # would go negative: halved
    assert 0.0 < p0.value < 1.0e-2 and p0.old_values == [1.0e-2]
    grid = spcl.SpectralGrid(np.linspace(3000., 3003., 4), units='cm_1')
    for q, par in enumerate(bs.params()):
        for k in range(3):                                                # three pixels
            par.store_deriv(spcl.SpectralIntensity(np.full(4, 10.0 * q + k), grid), num=k)
    J = bs.build_jacobian()
    assert J.shape == (12, 4) and np.array_equal(J[:, 2], np.repeat([20., 21., 22.], 4))
    masks = [np.array([1, 0, 1, 1]), np.ones(4), np.zeros(4)]
    assert bs.build_jacobian(masks).shape == (7, 4)
    assert bs.n_used_par() == 0
    bs.params()[1].set_used()
    assert bs.n_used_par() == 1


def test_group_observations_ladder_and_spline():
    """make_group_observations (smm:3290-3338) and make_radtran_spline (smm:3377-3396)."""
    pixels = S.vims_pixels([700.0, 450.0, 575.0], channels=np.linspace(2997., 3003., 7),
                           widths=np.full(7, 0.6))
    loss, alts, ssps, fszas = smm.make_group_observations(pixels, alt_step=50., alt_first_los=300.)
    assert [p.limb_tg_alt for p in pixels] == [450.0, 575.0, 700.0]           # sorted in place
    lo = pixels[0].low_LOS().get_tangent_altitude()
    hi = pixels[-1].up_LOS().get_tangent_altitude()
    assert alts[0] == 300.0 and alts[0] <= lo and alts[-1] >= hi and np.allclose(np.diff(alts), 50.)
    assert len(loss) == len(alts) == len(ssps) == len(fszas)
    for los, a in zip(loss, alts):
        assert abs(los.get_tangent_altitude() - a) < 0.05      # aimed at the ladder point (as the reference does)
    loss2, alts2, _, _ = smm.make_group_observations(pixels, alt_step=50., alt_first_los=900.)
    assert abs(alts2[0] - lo) < 1e-9                                           # capped at the lowest LOS
    grid = spcl.SpectralGrid(np.linspace(2997., 3003., 7), units='cm_1')
    za = np.array([300., 350., 400., 450., 500.])
    rads = [spcl.SpectralIntensity((1.0 + 0.01 * z) * np.arange(1., 8.), grid) for z in za]
    spl = smm.make_radtran_spline(za, rads)
    for z, r in zip(za, rads):
        assert np.allclose(spl(z).spectrum, r.spectrum, rtol=1e-12)
    assert np.allclose(spl(437.0).spectrum, (1.0 + 4.37) * np.arange(1., 8.), rtol=1e-12)  # linear in z


def test_latitude_boxes_and_2d_profile():
    """lat_box / centre_boxes / LinearProfile_2D (smm:355-439) and AtmGridMask.merge: literal box
    semantics (incl. the strict comparison that empties the last box), parameters keyed
    (box start, altitude node), profile() = sum of value x mask on (latitude band, altitude)."""
    lims = [-90., -30., 30., 60.]
    assert list(smm.lat_box(lims, -90.).mask) == [1., 0., 0., 0.]
    assert list(smm.lat_box(lims, 45.).mask) == [0., 0., 1., 0.]
    assert list(smm.lat_box(lims, 60.).mask) == [0., 0., 0., 0.]      # smm:368 uses '>'
    assert list(smm.lat_box(lims, 61.).mask) == [0., 0., 0., 1.]
    assert smm.centre_boxes(lims) == [-60., 0., 45.]
    planet = S.titan_planet(None, n_bands=7)
    atm = planet.atmosphere
    starts = list(atm.grid.coords['lat'][:-1])
    nodes = [300., 600., 900.]
    vals = [[0.01 * (1 + b), 0.02 * (1 + b), 0.015 * (1 + b)] for b in range(7)]
    prof = smm.LinearProfile_2D('CH4', atm, nodes, starts, vals, [[1e-3] * 3] * 7)
    assert prof.n_par == 21 and prof.set[4].key == (-75.0, 600.0)
    assert prof.set[4].maskgrid.mask.shape == (7, len(atm.grid.coords['alt']))
    p = prof.profile()
    z = atm.grid.coords['alt']
    for b in range(6):
        assert np.allclose(p.values['vmr'][b], np.interp(z, nodes, vals[b]), rtol=1e-14)
    assert np.all(p.values['vmr'][6] == 0.0)                           # the literal last-box quirk
    pt = sbm.Coords([40.0, 0.0, 450.0], s_ref='Spherical')
    assert abs(p.calc(pt, 'vmr') - sum(q.value * q.maskgrid.calc(pt) for q in prof.set)) < 1e-15
    assert prof.check_involved((30.0, 600.), dict(alt=[650., 1000.], lat=[35., 50.]))
    assert not prof.check_involved((60.0, 600.), dict(alt=[650., 1000.], lat=[35., 50.]))
    assert not prof.check_involved((30.0, 300.), dict(alt=[650., 1000.], lat=[35., 50.]))


def test_levels_fundamental_symmetries_and_default_atmosphere():
    """IsoMolec.add_levels(add_fundamental=True), add_simmetries_levels (run_0607_lut.py:94-100),
    Titan.add_default_atm (spect_robot.py:22), AtmProfile arithmetic."""
    im = sbm.IsoMolec(6, 1, LTE=False)
    im.add_levels(['0 0 1 0 1F2', '0 1 0 0 1E'], [3019.4935, 1533.3326], degeneracies=[3, 2],
                  add_fundamental=True)
    assert im.levels == ['lev_00', 'lev_01', 'lev_02']
    assert im.lev_00.minimal_level_string() == '0 0 0 0' and im.lev_00.energy == 0.0
    assert im.lev_01.degen == 3 and im.lev_02.energy == 1533.3326
    im.add_levels(['0 0 0 0 1A1'], [0.0], add_fundamental=True)     # already there by quanta: only the new entry
    assert im.n_lev == 4
    lines = [spcl.SpectLine(dict(Mol=6, Iso=1, Up_lev_str='    0 0 1 0 1F2', Lo_lev_str='    0 0 0 0 1A1')),
             spcl.SpectLine(dict(Mol=6, Iso=1, Up_lev_str='    0 0 1 0 1F1', Lo_lev_str='    0 1 0 0 1E ')),
             spcl.SpectLine(dict(Mol=6, Iso=2, Up_lev_str='    0 0 1 0 1A2', Lo_lev_str='    0 0 0 0 1A1'))]
    found = im.add_simmetries_levels(lines)
    assert found['lev_01'] == ['1F2', '1F1'] and im.lev_01.simmetry == ['1F1', '1F2']
    assert im.lev_02.simmetry == ['1E'] and im.lev_00.simmetry == ['1A1']
    planet = sbm.Titan(1200.)
    atm = planet.add_default_atm()
    assert planet.atmosphere is atm and atm.grid.coords['alt'][-1] == 1200.0
    p0 = atm.calc([0.0, 0.0, 0.0], 'pres')
    assert p0 == pytest.approx(1467.0) and atm.calc([0.0, 0.0, 300.0], 'pres') < 1.0
    two = atm * 2.0 + 1.0
    assert two.calc([0.0, 0.0, 100.0], 'temp') == pytest.approx(2 * atm.calc([0.0, 0.0, 100.0], 'temp') + 1)
    assert atm.calc([0.0, 0.0, 0.0], 'pres') == p0                     # operands untouched
    both = atm + atm
    assert both.temp[3] == 2 * atm.temp[3] and both.pres[3] == 2 * atm.pres[3]


def test_read_orbits(tmp_path):
    """smm.read_orbits (smm:2328-2347): '#' header, 14 numbers per line -> VIMSPixel."""
    fn = str(tmp_path / 'orbit_T34.dat')
    with open(fn, 'w') as f:
        f.write('geometry of the selected pixels\n#\n')
        f.write('3 12000. -10. 45. 512.5 -33. 120. 61.5 12. 90. -20. 200. 9.5 2007.55\n')
        f.write('4 12500. -11. 46. 612.5 -30. 121. 58.0 -5. 91. -20. 200. 9.5 2007.56\n\n')
    orbs = smm.read_orbits(fn, tag='T34')
    assert len(orbs) == 2 and orbs[0]['num'] == 3 and isinstance(orbs[0]['num'], int)
    assert orbs[1]['limb_tg_alt'] == 612.5 and orbs[1]['pixel_rot'] == -5.0 and orbs[0]['tag'] == 'T34'
    pix = sbm.VIMSPixel(orbs[0].keys(), orbs[0].values())
    assert pix.limb_tg_sza == 61.5 and pix.sub_solar_point().Spherical()[0] == pytest.approx(-20.0)
    assert pix.LOS().starting_point.Spherical()[2] == pytest.approx(12000.0)
    with pytest.raises(ValueError):
        smm.read_orbits(fn, formato='other')


def test_bayes_set_parerror_and_cpu_time_estimate(world):
    """BayesSet.update_parerror (smm:191-194), LookUpTable.CPU_time_estimate (smm:791-801)."""
    z = np.arange(0.0, 1001.0, 50.0)
    prof = smm.LinearProfile_1D_new('CH4', z, [200.0, 500.0, 800.0], [1e-2, 2e-2, 3e-2], [1e-3, 1e-3, 1e-3])
    bs = smm.BayesSet('t')
    bs.add_set(prof)
    bs.store_VCM(np.diag([4.0, 9.0, 16.0]))
    bs.update_parerror()
    assert [p.ret_error for p in bs.params()] == [2.0, 3.0, 4.0]
    im = world["planet"].gases['CH4'].iso_1
    lut = smm.LookUpTable(im, [2995.0, 3005.0], False)
    n = len([l for l in world["lines"] if l.Mol == im.mol and l.Iso == im.iso])
    assert lut.CPU_time_estimate(world["lines"], [[1, 2]] * 7) == pytest.approx(n * 3. / 30000. * 7)


def test_atmprofile_latitude_conventions():
    """Latitude coordinates as the reference's drivers pass them: band STARTS with ['box', ...]
    (radtran_3D_ch4.py:83-88) equal the band-edge form; band centres with ['lin', ...]
    (radtran_3Dvs2D_radtrans_new.py:82-87) interpolate linearly, constant outside; the device
    tables refuse the latter instead of silently treating centres as edges."""
    z = np.arange(0.0, 1001.0, 100.0)
    lat_ext = [-90., -75., -60., -30., 30., 60., 75., 90.]
    TT = np.array([150.0 + 5.0 * b + 0.01 * z for b in range(7)])
    starts = sbm.AtmProfile(sbm.AtmGrid(['lat', 'alt'], [lat_ext[:-1], z]), TT, 'temp', ['box', 'lin'])
    edges = sbm.AtmProfile(sbm.AtmGrid(['lat', 'alt'], [lat_ext, z]), TT, 'temp', ['box', 'lin'])
    assert np.array_equal(starts.lat_edges(), np.array(lat_ext)) and starts.n_band() == 7
    for lat in (-90.0, -80.0, -75.0, 0.0, 74.9, 75.0, 89.0, 90.0):
        assert starts.calc([lat, 250.0], 'temp') == edges.calc([lat, 250.0], 'temp')
    assert starts.calc([80.0, 0.0], 'temp') == 180.0                 # the last band is reachable
    with pytest.raises(ValueError):
        sbm.AtmProfile(sbm.AtmGrid(['lat', 'alt'], [lat_ext[:-2], z]), TT, 'temp', ['box', 'lin'])
    lat_c = [(a + b) / 2.0 for a, b in zip(lat_ext[:-1], lat_ext[1:])]
    lin = sbm.AtmProfile(sbm.AtmGrid(['lat', 'alt'], [lat_c, z]), TT, 'temp', ['lin', 'lin'])
    lin.add_profile(np.exp(-TT / 50.0), 'pres', ['lin', 'exp'])
    assert lin.calc([lat_c[2], 300.0], 'temp') == pytest.approx(TT[2, 3])
    mid = 0.5 * (lat_c[2] + lat_c[3])
    assert lin.calc([mid, 300.0], 'temp') == pytest.approx(0.5 * (TT[2, 3] + TT[3, 3]))
    assert lin.calc([-89.0, 300.0], 'temp') == TT[0, 3] and lin.calc([88.0, 300.0], 'temp') == TT[6, 3]
    assert lin.calc([mid, 350.0], 'pres') == pytest.approx(
        0.5 * (np.sqrt(lin.pres[2, 3] * lin.pres[2, 4]) + np.sqrt(lin.pres[3, 3] * lin.pres[3, 4])))
    one = sbm.AtmProfile(sbm.AtmGrid('alt', z), TT[0], 'vmr', 'lin')
    assert one.calc(250.0) == one.calc([0.0, 0.0, 250.0], 'vmr') == pytest.approx(152.5)
    planet = sbm.Titan(1000.)
    planet.add_atmosphere(lin)
    ch4 = sbm.Molec(6, 'CH4')
    ch4.add_iso(1)
    ch4.add_clim(sbm.AtmProfile(lin.grid, np.full(TT.shape, 0.015), 'vmr', ['lin', 'lin']))
    planet.add_gas(ch4)
    with pytest.raises(NotImplementedError):
        smm.planet_atmosphere_tables(planet, [('CH4', 'iso_1')])



def _small_lut(tmp_path):
    import torch
    rng = np.random.default_rng(9)
    im = sbm.IsoMolec(6, 1, LTE=False)
    im.add_levels(S.level_strings(2), [0.0, 1310.76])
    grid = spcl.SpectralGrid(np.linspace(3000.0, 3000.1, 33), units='cm_1')
    PT = [[0.01, 150.0], [0.01, 155.0], [0.1, 150.0]]
    lut = smm.LookUpTable(im, [3000.0, 3000.1], LTE=False)
    lut.PTcouples, lut.spectral_grid = PT, grid
    g = rng.uniform(0.5, 1.5, (3, 2, 3, 33)).astype(np.float32)
    lut.g32 = torch.as_tensor(g)
    for s, nam in enumerate(im.levels):
        st = smm.LutSet(6, 1, im.MM, level=getattr(im, nam))
        st.PTcouples, st.spectral_grid, st._table = PT, grid, (lut, s)
        lut.sets[nam] = st
    return lut, im, PT, g


def test_reference_readable_lut_files_and_check_lut_exists(tmp_path):
    """export_levels(for_reference=True): protocol-2 pickles under the reference's module names
    (and NumPy's pre-2.0 reconstructor path) + the skeleton file; check_LUT_exists with the
    reference's return values (smm:1390-1456) on that directory and on a LookUpTable.export file."""
    import pickle
    lut, im, PT, g = _small_lut(tmp_path)
    cart = str(tmp_path) + '/'
    assert smm.check_LUT_exists(PT, cart, 6, 1, False) == (False, PT, None, None)
    files = lut.export_levels(cart, stamp='_test', for_reference=True)
    assert sorted(files) == sorted(im.levels) + ['skeleton']
    raw = open(files[im.levels[1]], 'rb').read()
    assert raw[:2] == b'\x80\x02' and b'cspect_classes\nSpectralGcoeff\n' in raw
    assert b'spectrobot_b200' not in raw and b'numpy._core' not in raw and b'cnumpy.core.multiarray\n' in raw
    raw = open(files['skeleton'], 'rb').read()
    assert b'cspect_main_module\nLookUpTable\n' in raw and b'cspect_base_module\nIsoMolec\n' in raw
    assert b'spectrobot_b200' not in raw
    pts, sets = smm.read_lutset_stream(files[im.levels[1]])             # ... and reads back here
    assert pts == PT and np.array_equal(np.asarray(sets[2]['absorption'].spectrum, dtype=np.float32), g[2, 1, 2])
    sk = smm.read_Gcoeffs_from_LUTs(cart, os.path.basename(files['skeleton']))
    assert isinstance(sk, smm.LookUpTable) and sk.g32 is None and sk.PTcouples == PT
    assert sk.sets['lev_01'].filename == files['lev_01'] and sk.sets['lev_01'].sets == []
    assert lut.g32 is not None and lut.sets['lev_00']._table is not None      # the table itself is untouched
    st = sk.sets['lev_01']
    st.load_from_file(spectral_grid=lut.spectral_grid)
    assert np.array_equal(st.sets[0]['sp_emission'].spectrum, g[0, 1, 0].astype(float))
    want = PT + [[1.0, 160.0]]
    exists, todo, pt_map, wn_ranges = smm.check_LUT_exists(want, cart, 6, 1, False)
    assert exists and todo == [[1.0, 160.0]] and pt_map == [[files['skeleton'], PT]]
    assert wn_ranges == [[3000.0, 3000.1]]
    assert smm.check_LUT_exists(PT, cart, 6, 2, False)[0] is False              # another isotopologue
    one = lut.export(os.path.join(cart, 'LUT_mol06_iso1_nonLTE_single.pic'))    # the one-file form
    exists, todo, pt_map, wn_ranges = smm.check_LUT_exists(want[2:], cart, 6, 1, False)
    assert exists and todo == [[1.0, 160.0]] and len(pt_map) == 2 and pt_map[0][0] == one
    assert wn_ranges[0] == wn_ranges[1]


@pytest.mark.skipif(not os.path.exists('/root/reference/spect_main_module.py'),
                    reason="reference tree not present (GPU box)")
def test_reference_loads_the_files_written_for_it(tmp_path):
    """In a fresh interpreter that holds ONLY the reference's own modules (tests/golden/ref_exec.py),
    the for_reference files unpickle into the reference's classes, its LutSet.load_from_file reads
    the per-level stream and its check_LUT_exists finds the skeleton."""
    import subprocess
    import sys
    lut, im, PT, g = _small_lut(tmp_path)
    cart = str(tmp_path) + '/'
    files = lut.export_levels(cart, stamp='_test', for_reference=True)
    np.save(os.path.join(cart, 'g.npy'), g)
    here = os.path.dirname(os.path.abspath(__file__))
    code = r"""
import sys, pickle, numpy as np, warnings
warnings.simplefilter('ignore')
sys.path.insert(0, %r)
import ref_exec as R
spcl, smm, sbm = R.load()
assert 'spectrobot_b200.spect_classes' not in sys.modules
cart, lev = %r, %r
g = np.load(cart + 'g.npy')
with open(lev, 'rb') as f:
    pts = pickle.load(f)
    first = pickle.load(f)
assert type(first['absorption']) is spcl.SpectralGcoeff, type(first['absorption'])
st = smm.LutSet(6, 1, 16.0313, level=None, filename=lev)
grid = spcl.SpectralGrid(np.linspace(3000.0, 3000.1, 33), units='cm_1')
st.spectral_grid = grid
st.load_from_file()
assert st.PTcouples == %r and len(st.sets) == 3
assert np.array_equal(st.sets[2]['ind_emission'].spectrum, g[2, 1, 1].astype(float))
assert st.sets[2]['ind_emission'].integrate() > 0
with R.quiet():
    exists, todo, pt_map, wn = smm.check_LUT_exists(%r + [[1.0, 160.0]], cart, 6, 1, False)
assert exists and todo == [[1.0, 160.0]] and len(pt_map) == 1, (exists, todo)
sk = pickle.load(open(pt_map[0][0], 'rb'))
assert type(sk) is smm.LookUpTable and type(sk.sets['lev_01']) is smm.LutSet
assert type(sk.isomolec) is sbm.IsoMolec and sk.find_lev(sk.isomolec.lev_01.lev_string) == (True, 'lev_01')
print('reference loaded the files')
""" % (os.path.join(here, 'golden'), cart, files['lev_01'], PT, PT)
    env = dict(os.environ, PYTHONPATH='')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert out.returncode == 0 and 'reference loaded the files' in out.stdout, out.stderr[-2000:]


def test_split_files_default_and_reference_readable(tmp_path):
    """split_and_compress_LUTS (smm:1614-1728) on a table held in host memory: chunk files in this
    interpreter's pickle protocol and, with for_reference, under the reference's module names;
    both read back through read_split_file / LookUpTable.load_split's reader."""
    lut, im, PT, g = _small_lut(tmp_path)
    for k, for_ref in enumerate((False, True)):
        cart = str(tmp_path / ('split%d' % k)) + '/'
        os.makedirs(cart)
        allL, n_split, sp_grids = smm.split_and_compress_LUTS(lut.spectral_grid, {('CH4', 1): lut}, cart,
                                                              n_split=3, for_reference=for_ref)
        assert n_split == 3 and [len(s.grid) for s in sp_grids] == [11, 11, 11]
        assert len(lut.splitfiles) == 3 and os.path.basename(lut.splitfiles[1]).startswith('LUT_csplit01_mol06_iso1_nonLTE')
        raw = open(lut.splitfiles[1], 'rb').read()
        assert (b'cspect_main_module\nLutSet\n' in raw) == for_ref and (b'spectrobot_b200' in raw) != for_ref
        split = smm.read_split_file(lut.splitfiles[1])
        assert sorted(split) == ['lev_00', 'lev_01']
        st = split['lev_01']
        assert np.array_equal(st.spectral_grid.grid, lut.spectral_grid.grid[11:22]) and st.PTcouples == PT
        co = st.sets[2]['absorption']
        assert co.spectrum.dtype == np.float32 and co.spectral_grid is None
        assert np.array_equal(co.spectrum, g[2, 1, 2, 11:22])
