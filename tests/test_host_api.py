"""CPU tests of the host-side mirror of the reference API (spect_classes / spect_main_module /
spect_base_module): everything here is index logic and scalar physics, no GPU."""
import math as mt

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
