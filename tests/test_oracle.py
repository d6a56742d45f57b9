"""CPU tests of the oracle itself (no GPU): the reference ships no golden vectors, so the oracle is
pinned by the identities SURVEY.md section 4 lists (i)-(vii), by literal NumPy restatements of the
reference's Python statements, and by the committed fixtures in tests/golden/ (regression)."""
import os

import numpy as np
import pytest
from scipy.special import wofz

from conftest import rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def S():
    from spectrobot_b200 import synthetic
    return synthetic


def test_grid_is_numpy_arange_not_ideal(oracle, S):
    """the grid step is fl(fl(w0+5e-4)-w0), and the window has exactly 13010 points."""
    g = S.spectral_grid(2825.0, 3225.0)
    assert len(g) == 800001
    d = g[1] - g[0]
    assert d == 5.000000001018634e-4
    assert np.array_equal(g, 2825.0 + np.arange(800001) * d)
    L = oracle.line_window_offsets(g)
    assert len(L) == 13010 and L[6505] == 0.0
    for w0, w1, n in ((2050, 2250, 400001), (2900, 3200, 600001), (2850, 3450, 1200001)):
        assert len(S.spectral_grid(w0, w1)) == n


@pytest.mark.parametrize("P", [1e-6, 1e-2, 1.0, 300.0])
def test_voigt_normalisation_and_wofz(oracle, S, P):
    """(i) integral(shape)=1; humliv_bb is Humlicek w4: ~1e-4 from the true Voigt, not better."""
    g = S.spectral_grid(2990.0, 3010.0)
    L = oracle.line_window_offsets(g)
    nu0, T = 3000.1234567, 150.0
    ind, c = oracle.closest_grid(g, nu0)
    lw, dw = oracle.widths_c(nu0, 0.06, 0.7, T, P, 16.0313)
    sh = oracle.line_shape(nu0, lw, dw, c, L)
    dwp = dw / np.sqrt(np.log(2.0))
    ref = wofz((L + c - nu0) / dwp + 1j * lw / dwp).real / (dw * np.sqrt(np.pi / np.log(2.0)))
    assert 1e-6 < rel_err(sh, ref) < 5e-4
    tail = 2 * lw / np.pi / 3.2525            # Lorentz wings cut at +-3.2525 cm-1
    assert abs(sh.sum() * (g[1] - g[0]) - (1 - tail)) < 2e-4


def test_float32_artefacts_are_reproduced(oracle):
    """humliv_bb (float32 cmplx + float32 literals) differs from the D0-coefficient
    scalar humli_bb by ~1e-7..1e-6 in the core, and the two agree elsewhere (identity v)."""
    dw, lw = 3.0e-3, 3.0e-9
    x = 3000.0 + (np.arange(13010) - 6505) * 5e-4
    x0 = 3000.0 + 1.3e-4
    y = oracle.humliv_bb(x, 1, 13010, x0, lw, dw)
    reg = oracle.humliv_regions(x, 1, 13010, x0, lw, dw)
    ys = np.array([oracle.humli_bb((xx - x0) / dw, lw / dw) for xx in x])
    core = np.arange(reg[2] + 2, reg[3] - 2)
    d = np.abs(y[core] - ys[core]) / ys[core]
    assert 1e-9 < d.max() < 1e-5
    # region 1 of the vector routine == region 1 of the scalar routine (same rational)
    wing = np.r_[10:reg[0] - 2, reg[1] + 2:13000]
    assert rel_err(y[wing], ys[wing]) < 1e-6


def test_region_indices_are_index_based(oracle):
    """Appendix A: boundaries come from nint() of distances, asymmetric for an off-grid centre."""
    dw, lw = 3.3e-3 / np.sqrt(np.log(2)), 1e-5
    x = 3000.0 + (np.arange(13010) - 6505) * 5e-4
    il, ir, il2, ir2 = oracle.humliv_regions(x, 1, 13010, 3000.0 + 2.4e-4, lw, dw)
    assert 1 < il < il2 < 6506 < ir2 < ir < 13010
    assert (6506 - il) != (ir - 6506) or (6506 - il2) != (ir2 - 6506)


def test_sum_all_lines_semantics(oracle):
    spe = np.arange(20, dtype=float)
    m = np.zeros((5, 4), order="F")
    m[0] = [1, 2, 3, 4]
    m[1] = [10, 20, 30, 40]
    out = oracle.sum_all_lines(spe, m, [3, 18], [6, 19], 2)
    exp = spe.copy()
    exp[2:6] += [1, 2, 3, 4]
    exp[17:19] += [10, 20]
    assert np.array_equal(out, exp)


def test_curgod_identities(oracle):
    """(vi) curgod_2 == vmr*curgod_1 for constant vmr; curgod_3 == curgod_4 for constant f."""
    x = np.linspace(0.0, 400.0, 60)
    nd = 1e13 * np.exp(-x / 55.0)
    vmr = np.full_like(x, 0.015)
    f = np.full_like(x, 170.0)
    c1 = oracle.curgod(1, nd, x)
    assert abs(c1 - 1e13 * 55.0 * (1 - np.exp(-400.0 / 55.0))) < 1e-9 * c1   # analytic column
    assert abs(oracle.curgod(2, nd, vmr, x) - 0.015 * c1) < 1e-9 * c1
    vmr2 = 0.015 * (1 + x / 1000.0)
    c3 = oracle.curgod(3, nd, vmr2, f, x)
    c4 = oracle.curgod(4, nd, vmr2, f, x)
    assert abs(c3 - c4) < 1e-7 * abs(c3)
    assert abs(c4 - 170.0 * oracle.curgod(2, nd, vmr2, x)) < 1e-7 * abs(c4)


def test_tips_tables_vs_molparam(oracle):
    """(iv) two shipped tables cross-check: Lagrange-interpolated TIPS-2003 Q(296) vs molparam.txt
    Q(296 K) (different TIPS vintages -> agreement ~1e-2, a sanity check, not parity)."""
    import json
    mp = json.load(open(os.path.join(os.path.dirname(GOLDEN), "..", "spectrobot_b200", "data",
                                     "molparam.json")))
    for mol, iso in ((6, 1), (5, 1), (23, 1), (26, 1), (6, 2)):
        q = oracle.CalcPartitionSum(mol, iso, 296.0)
        q_mp = mp[str(mol)]["isos"][iso - 1]["Q296"]
        assert abs(q / q_mp - 1) < 2e-2, (mol, iso, q, q_mp)
    gi, t, qt = oracle.bd_tips_2003(6, 1)
    assert gi == 1.0 and t[0] == 60.0 and t[-1] == 3010.0 and len(t) == 119
    assert qt[0] == float(np.float32(0.54791E+02))       # real*4 literal widened
    with pytest.raises(KeyError):
        oracle.bd_tips_2003(40, 1)                        # fparts_mod.f:277-281 falls through


def test_partition_sum_scipy_vs_c(oracle):
    for T in (61.0, 70.0, 84.9, 85.0, 150.3, 296.0, 1000.0, 3005.0):
        a = oracle.CalcPartitionSum(6, 1, T)               # literal: scipy lagrange
        b = oracle.partition_sum_c(6, 1, T)
        assert abs(a - b) < 1e-10 * abs(a), T


def test_python_and_c_line_physics_agree(oracle, S):
    """The C restatement of widths / G coefficients against the literal NumPy statements."""
    lines = S.line_table(200, 2900.0, 3100.0, seed=2)
    for i in range(0, 200, 9):
        T, P = 70.0 + i, 10.0 ** (-6 + i / 30.0)
        lw, dw = oracle.widths_c(lines["freq"][i], lines["air_broad"][i], lines["t_dep"][i], T, P,
                                 S.CH4_MM)
        assert lw == oracle.Lorenz_width(T, P * oracle.hpa_to_atm, lines["t_dep"][i],
                                         lines["air_broad"][i])
        assert dw == oracle.Doppler_width(T, S.CH4_MM, lines["freq"][i])
        gc = oracle.gcoeffs_c(lines["freq"][i], lines["a_coeff"][i], lines["e_lower"][i],
                              lines["g_up"][i], lines["g_lo"][i], lines["e_vib_up"][i],
                              lines["e_vib_lo"][i], T)
        gp = oracle.Calc_Gcoeffs(lines["freq"][i], lines["a_coeff"][i], lines["e_lower"][i],
                                 lines["g_up"][i], lines["g_lo"][i], lines["e_vib_up"][i],
                                 lines["e_vib_lo"][i], T)
        assert np.allclose(gc, gp, rtol=1e-15, atol=0)


def test_lte_identity(oracle, S):
    """(ii)+(iii) sum_lev pop*(G_abs-G_ind)/Q at LTE == S(T)*shape with CalcStrength_at_T, for
    A derived with calc_A_coeff_from_strength (spect_classes.py:291, 1713)."""
    T = 180.0
    q296 = oracle.CalcPartitionSum(6, 1, 296.0)
    qT = oracle.CalcPartitionSum(6, 1, T)
    lines = S.line_table(50, 2900.0, 3100.0, q296=q296, seed=7)
    for i in range(50):
        g = oracle.Calc_Gcoeffs(lines["freq"][i], lines["a_coeff"][i], lines["e_lower"][i],
                                lines["g_up"][i], lines["g_lo"][i], lines["e_vib_up"][i],
                                lines["e_vib_lo"][i], T)
        pop_lo = oracle.Boltz_ratio_nodeg(lines["e_vib_lo"][i], T)
        pop_up = oracle.Boltz_ratio_nodeg(lines["e_vib_up"][i], T)
        s_ein = S.CH4_RATIO * (g[2] * pop_lo - g[1] * pop_up) / qT
        f = lambda t, q: (oracle.Boltz_ratio_nodeg(lines["e_lower"][i], t) *
                          (1 - oracle.Boltz_ratio_nodeg(lines["freq"][i], t)) / q)
        s_hit = lines["strength"][i] * f(T, qT) / f(296.0, q296)
        assert abs(s_ein / s_hit - 1) < 1e-10


def test_cell_threads_and_golden(oracle, S):
    g = S.spectral_grid(2999.0, 3001.0)
    lines = S.line_table(80, 2996.0, 3004.0, n_levels=4, seed=12)
    a = oracle.gcoeff_cell(lines, g, 150.0, 0.1, S.CH4_MM, 4, n_threads=1)
    b = oracle.gcoeff_cell(lines, g, 150.0, 0.1, S.CH4_MM, 4, n_threads=3)
    assert np.array_equal(a, b)
    gold = np.load(os.path.join(GOLDEN, "cell_small.npz"))
    assert rel_err(a[:, :, ::40], gold["cell"]) < 1e-12


def test_lut_rule_literal_vs_c(oracle):
    """(vii) + index logic: C orc_lut_weights against the literal NumPy LutSet.calculate."""
    cells = [[p, t] for p in (1e-3, 2.7e-3, 7.4e-3, 2e-2) for t in (140., 145., 150., 155., 160.)]
    rng = np.random.default_rng(5)
    specs = [rng.uniform(size=7) for _ in cells]
    for _ in range(300):
        P = np.exp(rng.uniform(np.log(2e-4), np.log(2e-2)))
        T = rng.uniform(140.0, 160.0)
        ref = oracle.LutSet_calculate(cells, specs, P, T)
        c, w = oracle.lut_weights(cells, P, T)
        got = sum(w[i] * specs[c[i]] for i in range(4) if c[i] >= 0)
        assert np.allclose(got, ref, rtol=1e-13, atol=0)
    # at a node the cell is reproduced exactly
    c, w = oracle.lut_weights(cells, 7.4e-3, 150.0)
    got = sum(w[i] * specs[c[i]] for i in range(4) if c[i] >= 0)
    assert np.allclose(got, specs[cells.index([7.4e-3, 150.0])], rtol=1e-15)
    with pytest.raises(ValueError):
        oracle.lut_weights(cells, 0.5, 150.0)


def test_los_oracle_limits(oracle, S):
    """LOS spec sanity: optically thick isothermal LTE column -> Planck-like source J/tau, and the
    materialised recursion equals the fused one."""
    g = S.spectral_grid(2999.5, 3000.5)
    lines = S.line_table(30, 2996.0, 3004.0, n_levels=3, seed=3)
    cells = [[p, t] for p in (1e-3, 2.7e-3) for t in (145., 150., 155.)]
    g32 = np.stack([oracle.gcoeff_cell(lines, g, T, P, S.CH4_MM, 3) for P, T in cells]).astype(np.float32)
    lut = dict(g32=g32, pt=np.array(cells), level_energy=lines["level_energies"], mol=6, iso=1,
               iso_ratio=1.0, lte_unidentified=False)
    n_steps = np.array([3, 2], dtype=np.int32)
    temp = np.array([[150., 151., 149.], [150., 150., 0.]])
    pres = np.array([[1.5e-3, 2e-3, 1.2e-3], [2e-3, 2e-3, 0.]])
    col = np.array([[[1e30, 1e30, 1e30], [1e14, 1e14, 0.]]])
    rad, tau, src = oracle.los_rt([lut], n_steps, temp, pres, col, None, materialise=True)
    assert np.all(np.isfinite(rad))
    rad2 = oracle.los_layers(tau, src, n_steps)
    assert rel_err(rad2, rad) < 1e-12
    k = np.argmax(tau[0, 2])
    assert tau[0, 2, k] > 50 and abs(rad[0, k] / src[0, 2, k] - 1) < 1e-12   # thick: I = S(last)
    thin = tau[1, :2].sum(axis=0) < 1e-6
    assert thin.any()
    lin = (tau[1, :2] * src[1, :2]).sum(axis=0)
    assert rel_err(rad[1][thin], lin[thin], floor_rel=1e-12) < 1e-5          # thin: I = sum J


def test_oracle_step_builder_identities(oracle):
    """The oracle's LOS step builder pinned by what must hold for any correct one:
    a horizontal ray is symmetric about its tangent point; the step columns add up to the column of
    the whole path (additivity of the Curtis-Godson integrals, curgods.f); Curtis-Godson T and P
    lie inside the range of the step; for a constant VMR the gas column is vmr x air column."""
    from spectrobot_b200 import synthetic as S
    atm = S.titan_atmosphere(n_bands=1)
    z = atm["z"]
    vmr = np.full((1, 1, len(z)), 0.015)
    org = np.array([[1.0e5, 0.0, 0.0]])
    tg = np.array([0.0, 0.0, 2575.0 + 600.0])
    d = (tg - org[0]) / np.linalg.norm(tg - org[0])
    fine = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr, org, [d], max_T_variation=2.0,
                                  max_Plog_variation=0.3)[0]
    one = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr, org, [d], max_T_variation=1e9,
                                 max_Plog_variation=1e9)[0]
    assert one["n_steps"] == 1 and fine["n_steps"] > 10
    col = np.array(fine["column"][0])
    assert abs(col.sum() - one["column"][0][0]) < 1e-10 * col.sum()
    assert np.allclose(col, col[::-1], rtol=1e-6)                    # symmetric about the tangent point
    assert np.allclose(fine["temp"], fine["temp"][::-1], rtol=1e-8)
    assert min(atm["temp"][0]) <= min(fine["temp"]) and max(fine["temp"]) <= max(atm["temp"][0])
    assert np.all(np.diff(fine["pres"][:fine["n_steps"] // 2]) > 0)  # inbound half: P rises
    # tangent altitude above the top: no step; below the surface: the ray stops at the ground
    miss = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr, org,
                                  [(np.array([0, 0, 2575.0 + 1600.0]) - org[0]) / 1.0e5])[0]
    assert miss["n_steps"] == 0
    dn = np.array([0.0, 0.0, 2575.0 - 800.0]) - org[0]
    hit = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr, org, [dn / np.linalg.norm(dn)])[0]
    assert hit["n_steps"] > 3 and hit["pres"][0] > 1000.0            # first (far) step is at the surface


def test_oracle_step_builder_latitude_linear(oracle):
    """lat_centres mode of the oracle's step builder (`['lin', ...]` profiles of
    radtran_3Dvs2D_radtrans_new.py:82-111): identical rows give the tables of the 1-D planet; a
    temperature that rises linearly with latitude gives, on a ray inside one meridian plane slab,
    step temperatures between those of the two bracketing rows; outside the first / last centre
    the profile is constant."""
    from spectrobot_b200 import synthetic as S
    atm = S.titan_atmosphere(n_bands=1)
    z = atm["z"]
    cen = np.array([-60.0, 0.0, 60.0])
    rep = lambda a: np.repeat(np.asarray(a), 3, axis=0)   # noqa: E731
    vmr1 = np.full((1, 1, len(z)), 0.015)
    vmr3 = np.full((1, 3, len(z)), 0.015)
    org = np.array([[1.0e5, 0.0, 0.0]])
    tg = np.array([0.0, 2575.0 + 500.0, 900.0])       # tangent point at about 16 deg latitude
    d = (tg - org[0]) / np.linalg.norm(tg - org[0])
    one = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr1, org, [d])[0]
    same = oracle.los_steps_build(z, rep(atm["temp"]), rep(atm["pres"]), vmr3, org, [d],
                                  lat_centres=cen)[0]
    assert same["n_steps"] == one["n_steps"] > 3
    for k in ("temp", "pres"):
        assert np.allclose(same[k], one[k], rtol=1e-12)
    assert np.allclose(same["column"][0], one["column"][0], rtol=1e-12)
    # rows 0 K, +10 K, +20 K warmer: every step lies between the 1-D value + 10 and + 20 K
    # (all samples of this ray are between 0 and 60 degrees north)
    t3 = np.stack([atm["temp"][0], atm["temp"][0] + 10.0, atm["temp"][0] + 20.0])
    warm = oracle.los_steps_build(z, t3, rep(atm["pres"]), vmr3, org, [d], lat_centres=cen,
                                  max_T_variation=1e9, max_Plog_variation=0.2)[0]
    base = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr1, org, [d],
                                  max_T_variation=1e9, max_Plog_variation=0.2)[0]
    assert warm["n_steps"] == base["n_steps"]
    dt = np.array(warm["temp"]) - np.array(base["temp"])
    assert np.all(dt > 10.0) and np.all(dt < 20.0)
    # a ray over the pole region beyond the last centre sees the last row only
    tgp = np.array([0.0, 300.0, 2575.0 + 500.0])
    dp = (tgp - org[0]) / np.linalg.norm(tgp - org[0])
    pole = oracle.los_steps_build(z, t3, rep(atm["pres"]), vmr3, org, [dp], lat_centres=cen,
                                  max_T_variation=1e9, max_Plog_variation=0.2)[0]
    pole1 = oracle.los_steps_build(z, atm["temp"] + 20.0, atm["pres"], vmr1, org, [dp],
                                   max_T_variation=1e9, max_Plog_variation=0.2)[0]
    assert pole["n_steps"] == pole1["n_steps"]
    mid = pole["n_steps"] // 2
    assert abs(pole["temp"][mid] - pole1["temp"][mid]) < 1e-9    # tangent region: > 60 deg north
