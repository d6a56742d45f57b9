"""GPU tests of the reference-shaped Python API (spect_classes / spect_main_module /
spect_base_module) against the oracle: the per-line path (calc_shapes_lines -> BuildCoeff ->
sum_all_lines), the batched LUT builder, the coefficient assembly, the instrument convolution and
the radtrans driver."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL_XS, TOL_RAD = 1e-6, 1e-5


@pytest.fixture(scope="module")
def world():
    import torch
    from spectrobot_b200 import engine, spect_base_module as sbm, spect_classes as spcl
    from spectrobot_b200 import spect_main_module as smm, synthetic as S
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    tab = S.line_table(160, 2993.0, 3007.0, n_levels=5, seed=21, frac_unlinked=0.05)
    lines = S.spect_lines(tab)
    planet = S.titan_planet(tab["level_energies"])
    sp = smm.prepare_spe_grid([2996.0, 3004.0]).spectral_grid
    return dict(torch=torch, engine=engine, sbm=sbm, spcl=spcl, smm=smm, S=S, tab=tab,
                lines=lines, planet=planet, sp=sp, im=planet.gases['CH4'].iso_1)


def test_reference_shaped_cell_path_matches_batched_builder_and_oracle(world, oracle):
    """calc_shapes_lines + SpectralGcoeff.BuildCoeff(preCalc_shapes=True) (the reference's
    LutSet.add_PT path through humliv_bb/sum_all_lines) == LookUpTable.make == oracle."""
    spcl, smm, S, im, sp = world["spcl"], world["smm"], world["S"], world["im"], world["sp"]
    P, T = 0.02, 158.0
    proc = spcl.calc_shapes_lines(sp, world["lines"], T, P, im)
    assert len(proc) == int(np.sum(world["tab"]["up_set"] >= 0))
    assert abs(proc[0].shape.integrate() - 1.0) < 2e-3           # "integral(shape)=1", spcl:1994
    ref = oracle.gcoeff_cell(world["tab"], sp.grid, T, P, S.CH4_MM, 5)
    lut = smm.LookUpTable(im, [2996.0, 3004.0], LTE=False)
    lut.make(sp, world["lines"], [[P, T], [P, T + 5.0]])
    g32 = lut.g32.cpu().numpy()
    for s, lev in enumerate(im.levels):
        st = smm.LutSet(6, 1, im.MM, level=getattr(im, lev))
        set_ = st.add_PT(sp, proc, P, T, keep_memory=True)
        for k, ct in enumerate(spcl.CTYPES):
            got = set_[ct].spectrum
            if np.max(np.abs(ref[s, k])) == 0.0:
                assert np.all(got == 0.0)
                continue
            assert rel_err(got, ref[s, k]) < TOL_XS, (lev, ct)
            assert rel_err(g32[0, s, k], ref[s, k].astype(np.float32)) < 2e-7, (lev, ct)


def test_make_abscoeff_LUTS_fast_matches_literal_restatement(world, oracle):
    smm, spcl, im, sp = world["smm"], world["spcl"], world["im"], world["sp"]
    PT = [[p, float(t)] for p in (1e-3, 1e-2, 1e-1) for t in (150., 155., 160., 165.)]
    lut = smm.LookUpTable(im, [2996.0, 3004.0], LTE=False)
    lut.make(sp, world["lines"], PT)
    Temps, Press = [152.3, 158.8, 163.0], [5e-4, 3e-3, 0.04]
    tv = np.array([[t + 3.0 * s for t in Temps] for s in range(5)])
    for s, lev in enumerate(im.levels):
        getattr(im, lev).local_vibtemp = list(tv[s])
    allL = {(im.mol_name, im.iso): lut}
    a, e = smm.make_abscoeff_LUTS_fast(sp, im, Temps, Press, LTE=False, allLUTs=allL)
    olut = dict(g32=lut.g32.cpu().numpy(), pt=np.array(PT), level_energy=im.level_energies(),
                mol=6, iso=1, lte_unidentified=False)
    ra, re = oracle.make_abscoeff_LUTS_fast(olut, Temps, Press, tvib=tv)
    for k in range(3):
        assert rel_err(a[k].spectrum, ra[k]) < 1e-9
        assert rel_err(e[k].spectrum, re[k]) < 1e-9
    # the host LutSet.calculate path gives the same interpolated G coefficients
    Gco = lut.sets['lev_00'].calculate(Press[1], Temps[1])
    assert Gco['sp_emission'] is None                  # level 0 is never an upper level
    sets = [olut["g32"][c, 0, 2].astype(float) for c in range(len(PT))]
    assert np.array_equal(Gco['absorption'].spectrum,
                          oracle.LutSet_calculate(PT, sets, Press[1], Temps[1]))


def test_convolve_lowres_matches_reference_formula(world, oracle):
    eng, spcl = world["engine"], world["spcl"]
    rng = np.random.default_rng(9)
    x = np.sort(np.concatenate([np.arange(3000.0, 3010.0, 5e-4), rng.uniform(3000.0, 3010.0, 500)]))
    y = rng.uniform(0.0, 1.0, (3, len(x))) * np.exp(-((x - 3004.0) / 2.0) ** 2)
    centres = np.array([2990.0, 3000.2, 3003.0, 3005.5, 3009.9, 3010.0 + 5e-4 * 0.4])
    widths = np.array([0.5, 0.3, 0.8, 0.05, 0.4, 1e-5])
    got = eng.convolve_lowres_host(x, y, centres, widths)
    for s in range(3):
        ref = oracle.convolve_to_grid_from_irregular(x, y[s], centres, widths)
        assert got[s, 0] == 0.0 and ref[0] == 0.0               # no hi-res point in the window
        assert rel_err(got[s], ref, floor_rel=1e-12) < 1e-12
    so = spcl.SpectralIntensity(y[0], spcl.SpectralGrid(x, units='cm_1'))
    obs = spcl.SpectralIntensity(np.zeros(6), spcl.SpectralGrid(centres, units='cm_1'))
    low = so.hires_to_lowres(obs, spectral_widths=list(widths))
    assert np.array_equal(low.spectrum, got[0])


def test_radtrans_driver_matches_oracle_pipeline(world, oracle):
    """smm.radtrans: pixels -> 3 LOS -> Curtis-Godson steps -> batched GPU radiances in two
    wavenumber chunks -> low-res channels; checked against the oracle run on the same step tables
    followed by the literal convolution."""
    smm, S, planet, sp, sbm = world["smm"], world["S"], world["planet"], world["sp"], world["sbm"]
    centres = np.linspace(2997.0, 3003.0, 7)
    widths = np.full(7, 0.6)
    pixels = S.vims_pixels([450.0, 700.0], channels=centres, widths=widths)
    inputs = dict(n_split=2, cart_LUTS=None, out_dir=None, n_threads=8)
    LUTopt = dict(pres_step_log=1.0, temp_step=5.0)
    sims, rt, single = smm.radtrans(inputs, planet, world["lines"], pixels, sp_gri=sp,
                                    radtran_opt=dict(max_T_variation=5., max_Plog_variation=1.),
                                    LUTopt=LUTopt)
    assert len(sims) == 2 and len(rt) == 6 and sorted(single) == [('CH4', 'iso_1')]
    # oracle: rebuild the same LOS, steps and LUT cells; hi-res on the CPU, then convolve
    pix = sorted(pixels, key=lambda p: p.limb_tg_alt)
    loss = []
    for p in pix:
        loss += [p.low_LOS(), p.LOS(), p.up_LOS()]
    for los in loss:
        los.calc_atm_intersections(planet)
        los.calc_radtran_steps(planet, None, max_T_variation=5., max_Plog_variation=1.)
    gi, steps = smm.los_step_tables(loss, planet)
    max_p = max(planet.atmosphere.calc(p.low_LOS().get_tangent_point(), 'pres') for p in pix)
    PT = smm.calc_PT_couples_atmosphere(world["lines"], list(planet.gases.values()),
                                        planet.atmosphere, max_pres=max_p, **LUTopt)
    lut = smm.LookUpTable(world["im"], [2996.0, 3004.0], LTE=False)
    lut.make(sp, world["lines"], PT)
    olut = dict(g32=lut.g32.cpu().numpy(), pt=np.array(PT),
                level_energy=world["im"].level_energies(), mol=6, iso=1,
                iso_ratio=S.CH4_RATIO, lte_unidentified=False)
    hi = oracle.los_rt([olut], steps.n_steps, steps.temp, steps.pres, steps.column, steps.tvib)
    assert hi.max() > 0
    for i, los in enumerate(loss):
        ref = oracle.convolve_to_grid_from_irregular(sp.grid, hi[i], centres, widths)
        got = rt['LOS%02d' % i].spectrum
        assert rel_err(got, ref, floor_rel=1e-9) < TOL_RAD, i
    # sims: FOV integral of the pixel's three LOS (FOV_integr_1D, smm:3273-3277) against the
    # literal spline + quad restatement
    three = np.array([rt['LOS%02d' % i].spectrum for i in range(3)])
    rot = getattr(pix[0], 'pixel_rot', 0.0) or 0.0
    assert np.allclose(sims[0].spectrum, oracle.FOV_integr_1D(three, centres, rot), rtol=2e-7, atol=0)
    # single-LOS reference-shaped call gives the same hi-res spectrum as the batch
    one = loss[1].radtran_fast(sp, planet, LUTS={(world["im"].mol_name, 1): lut})
    assert rel_err(one[0].spectrum, hi[1]) < TOL_RAD


def test_lut_level_files_round_trip_through_the_device(world, tmp_path):
    """LookUpTable.make -> export_levels (reference per-level pickle streams) -> import_levels
    into a fresh table: identical float32 numbers, and the reloaded table drives the LOS path."""
    smm, im, sp, eng = world["smm"], world["im"], world["sp"], world["engine"]
    PT = [[p, float(t)] for p in (1e-3, 1e-2) for t in (150., 155., 160.)]
    lut = smm.LookUpTable(im, [2996.0, 3004.0], LTE=False)
    lut.make(sp, world["lines"], PT)
    files = lut.export_levels(str(tmp_path), stamp='_rt')
    back = smm.LookUpTable(im, [2996.0, 3004.0], LTE=False).import_levels(files, sp)
    assert back.PTcouples == lut.PTcouples
    assert np.array_equal(back.g32.cpu().numpy(), lut.g32.cpu().numpy())
    assert eng.lut_row_stride(back.g32) % 32 == 0
    steps = eng.LosSteps([2], [[152.0, 157.0]], [[2e-3, 5e-3]], [[[1e17, 2e17]]], None)
    a = eng.los_rt_lut([lut.device_lut()], steps).cpu().numpy()
    b = eng.los_rt_lut([back.device_lut()], steps).cpu().numpy()
    assert a.max() > 0 and np.array_equal(a, b)


def _bayes(smm, planet, nodes=(300., 500., 700., 900., 1100.), vmr=0.015):
    z = planet.atmosphere.grid.coords['alt']
    prof = smm.LinearProfile_1D_new('CH4', z, list(nodes), [vmr] * len(nodes), [0.3 * vmr] * len(nodes))
    bs = smm.BayesSet('ch4')
    bs.add_set(prof)
    return bs


def test_inversion_fast_limb_forward_and_jacobian(world):
    """smm.inversion_fast_limb (one forward + Jacobian evaluation, smm:2598-2956): the forward part
    equals smm.radtrans; the derivative columns of the step builder add up to the column
    (linearity of curgod_fort_2 in the VMR); the FOV-integrated analytic Jacobian matches central
    finite differences of the whole host path on an LTE planet (where a step's T, P and T_vib do
    not depend on the VMR)."""
    smm, S = world["smm"], world["S"]
    planet = S.titan_planet(world["tab"]["level_energies"], nonlte=False)
    sp = world["sp"]
    centres = np.linspace(2997.0, 3003.0, 7)
    widths = np.full(7, 0.6)
    inputs = dict(n_split=1, cart_LUTS=None, out_dir=None, n_threads=8)
    LUTopt = dict(pres_step_log=1.0, temp_step=5.0)
    opt = dict(max_T_variation=5., max_Plog_variation=1.)
    bs = _bayes(smm, planet)

    def run(bset):
        pixels = S.vims_pixels([450.0, 700.0], channels=centres, widths=widths)
        return smm.forward_jacobian_limb(inputs, planet, world["lines"], bset, pixels, sp_gri=sp,
                                       radtran_opt=opt, LUTopt=dict(LUTopt))

    sims, rt, derivs = run(bs)
    assert len(sims) == 2 and len(rt) == 6 and len(derivs) == 6 * bs.n_tot
    pixels = S.vims_pixels([450.0, 700.0], channels=centres, widths=widths)
    sims_f, rt_f, _ = smm.radtrans(inputs, planet, world["lines"], pixels, sp_gri=sp,
                                   radtran_opt=opt, LUTopt=dict(LUTopt))
    for a, b in zip(sims, sims_f):
        assert rel_err(a.spectrum, b.spectrum) < 1e-7     # radtrans' low-res sink keeps its layers in float32 (~1e-9)
    # node at 300 km lies below both tangent heights (first node: mask = 1 below it, so it is
    # still involved through its triangle up to 500 km); the 1100 km node only feeds the top
    jac = bs.build_jacobian()
    assert jac.shape == (2 * 7, bs.n_tot) and np.all(np.isfinite(jac))
    assert bs.n_used_par() == bs.n_tot
    # derivative columns: sum_p value_p * d u/d p == u for every step of every LOS
    pix = sorted(pixels, key=lambda p: p.limb_tg_alt)
    los = pix[0].LOS()
    los.calc_atm_intersections(planet)
    los.calc_radtran_steps(planet, None, calc_derivatives=True, bayes_set=bs, **opt)
    for st in los.radtran_steps['step']:
        tot = sum(par.value * st['dcolumns'][(par.nameset, par.key)] for par in bs.params())
        assert abs(tot - st['columns']['CH4']) <= 1e-10 * st['columns']['CH4']
    # finite differences of the full host path
    for q in (1, 2, 3):
        up, dn = _bayes(smm, planet), _bayes(smm, planet)
        h = 1e-3 * up.params()[q].value
        up.params()[q].value += h
        dn.params()[q].value -= h
        s_up, _, _ = run(up)
        s_dn, _, _ = run(dn)
        for k in range(2):
            fd = (s_up[k].spectrum - s_dn[k].spectrum) / (2 * h)
            got = bs.params()[q].derivatives[k].spectrum
            scale = max(np.abs(fd).max(), np.abs(jac[k * 7:(k + 1) * 7]).max() * 1e-3)
            assert np.abs(got - fd).max() < 2e-5 * scale, (q, k)
    # reference-shaped single-LOS call: hires_deriv attached to the returned parameter set
    max_p = max(planet.atmosphere.calc(p.low_LOS().get_tangent_point(), 'pres') for p in pix)
    PT = smm.calc_PT_couples_atmosphere(world["lines"], list(planet.gases.values()),
                                        planet.atmosphere, max_pres=max_p, **LUTopt)
    LUTS = smm.check_and_build_allluts(inputs, sp, world["lines"], list(planet.gases.values()),
                                       PTcouples=PT, LUTopt=LUTopt)
    one = los.radtran_fast(sp, planet, LUTS=LUTS, calc_derivatives=True, bayes_set=bs)
    assert len(one[2].params()) == bs.n_tot
    assert one[2].params()[2].hires_deriv.spectrum.shape == one[0].spectrum.shape


def test_device_step_builder_matches_host_builder(world):
    """sr_los_steps_build (geometry + adaptive merge + Curtis-Godson integrals for a whole batch,
    SURVEY 8f row 4) against sbm.LineOfSight.calc_atm_intersections / calc_radtran_steps run per
    LOS: same step counts, tables equal to rounding, incl. a 7-band atmosphere, non-LTE
    vibrational temperatures, a ray that hits the surface, one that misses the atmosphere, and the
    derivative-column table of a VMR parameter set."""
    smm, S, sbm = world["smm"], world["S"], world["sbm"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=7)
    bs = _bayes(smm, planet)
    planet.gases['CH4'].add_clim(bs.sets['CH4'].profile())
    obs = sbm.Coords([5.0, 90.0, 1.0e5], s_ref='Spherical')
    loss = []
    for alt, lat in ((-300.0, 12.0), (150.0, -50.0), (420.0, 70.0), (777.7, 28.0), (1200.0, -80.0),
                     (1499.0, 0.0), (1600.0, 10.0)):
        loss.append(sbm.LineOfSight(obs, sbm.Coords([lat, 3.0, alt], s_ref='Spherical')))
    opt = dict(max_T_variation=4.0, max_Plog_variation=0.8)
    gi, steps, dfrac = smm.los_step_tables_device(loss, planet, bayes_set=bs, set_name='CH4', **opt)
    assert steps.n_steps[-1] == 0 and steps.n_steps[0] > 0
    host = []
    for los in loss[:-1]:
        los2 = sbm.LineOfSight(los.starting_point, los.second_point)
        los2.calc_atm_intersections(planet)
        los2.calc_radtran_steps(planet, None, calc_derivatives=True, bayes_set=bs, **opt)
        host.append(los2)
    gi_h, steps_h = smm.los_step_tables(host, planet)
    assert gi_h == gi
    nmax = steps_h.n_steps_max
    assert np.array_equal(steps.n_steps[:-1], steps_h.n_steps) and steps.n_steps_max >= nmax
    tol = dict(rtol=1e-10, atol=0)
    assert np.allclose(steps.temp[:-1, :nmax], steps_h.temp, **tol)
    assert np.allclose(steps.pres[:-1, :nmax], steps_h.pres, **tol)
    assert np.allclose(steps.column[:, :-1, :nmax], steps_h.column, **tol)
    assert np.allclose(steps.tvib[:, :, :-1, :nmax], steps_h.tvib, **tol)
    dfrac_h = smm.los_jac_tables(host, bs, 'CH4', nmax)
    assert np.allclose(dfrac[:-1, :nmax], dfrac_h, rtol=1e-9, atol=1e-14)
    assert np.all(dfrac[-1] == 0.0)
    for l, los in enumerate(host):
        for par in bs.params():
            assert loss[l].involved_retparams[(par.nameset, par.key)] == \
                los.involved_retparams[(par.nameset, par.key)]
    # LOS blocks of the per-point scratch are invisible
    import os
    os.environ["SR_STEPS_BLOCK"] = "3"
    try:
        _, st_b, df_b = smm.los_step_tables_device(loss, planet, bayes_set=bs, set_name='CH4', **opt)
    finally:
        del os.environ["SR_STEPS_BLOCK"]
    assert np.array_equal(st_b.n_steps, steps.n_steps) and np.array_equal(st_b.temp, steps.temp)
    assert np.array_equal(st_b.column, steps.column) and np.array_equal(st_b.tvib, steps.tvib)
    assert np.array_equal(df_b, dfrac)
    # a narrow table is widened by the wrapper (SR_ERR_LIMIT -> retry)
    atm = smm.planet_atmosphere_tables(planet, gi)
    org = np.array([l.starting_point.Cartesian() for l in loss])
    drc = np.array([l.direction for l in loss])
    st2, _ = world["engine"].los_steps_build(atm, org, drc, n_steps_max=2, **opt)
    assert np.array_equal(st2.n_steps, steps.n_steps)
    assert np.array_equal(st2.temp[:, :nmax], steps.temp[:, :nmax])


def test_device_step_builder_matches_oracle(world, oracle):
    """sr_los_steps_build against the CPU restatement (oracle.los_steps_build: NumPy geometry and
    merge, the oracle's own C curgod integrals)."""
    eng, S = world["engine"], world["S"]
    atm = S.titan_atmosphere()
    z, nb = atm["z"], len(atm["temp"])
    energies = world["tab"]["level_energies"]
    rng = np.random.default_rng(5)
    vmr = np.stack([np.full((nb, len(z)), 0.015) * (1 + 0.2 * np.sin(z / 150.0)),
                    np.full((nb, len(z)), 2e-6) * np.exp(-z / 900.0)])
    tvib = np.stack([np.stack([np.stack(S.vib_temperatures(z, atm["temp"][b], energies, 40.0 + 5 * b))
                               for b in range(nb)], axis=1)] * 2)           # [gas][set][band][z]
    tvib_on = np.ones((2, len(energies)), dtype=np.int32)
    tvib_on[1, 2:] = -1
    tvib_on[1, 1] = 0
    masks = np.stack([np.clip(1 - np.abs(z - c) / 200.0, 0, 1) for c in (300., 500., 700.)])
    n_los = 9
    th = rng.uniform(-1.2, 1.2, n_los)
    org = 1.0e5 * np.stack([np.cos(th), np.zeros(n_los), np.sin(th)], axis=1)
    tgt_alt = np.array([-500., 80., 350., 500., 650., 800., 1000., 1400., 1700.])
    drc = []
    for l in range(n_los):     # aim at a point at the wanted tangent altitude, off the origin's meridian
        t = np.array([-np.sin(th[l]), 0.3, np.cos(th[l])])
        t /= np.linalg.norm(t)
        tg = (2575.0 + tgt_alt[l]) * t
        d = tg - org[l]
        drc.append(d / np.linalg.norm(d))
    drc = np.array(drc)
    A = eng.Atmosphere(z, atm["temp"], atm["pres"], vmr, tvib=tvib, tvib_on=tvib_on,
                       lat_edges=atm["lat_edges"])
    opt = dict(delta_x=7.5, max_T_variation=3.0, max_Plog_variation=0.7)
    steps, dfrac = eng.los_steps_build(A, org, drc, masks=masks, jac_gas=0, **opt)
    ref = oracle.los_steps_build(z, atm["temp"], atm["pres"], vmr, org, drc, tvib=tvib,
                                 tvib_on=tvib_on, lat_edges=atm["lat_edges"], masks=masks,
                                 jac_gas=0, **opt)
    assert steps.n_steps.max() > 20 and steps.n_steps.min() == 0
    for l, r in enumerate(ref):
        n = r["n_steps"]
        assert steps.n_steps[l] == n, l
        if n == 0:
            continue
        tol = dict(rtol=1e-10, atol=0)
        assert np.allclose(steps.temp[l, :n], r["temp"], **tol)
        assert np.allclose(steps.pres[l, :n], r["pres"], **tol)
        for m in range(2):
            assert np.allclose(steps.column[m, l, :n], r["column"][m], **tol)
            for j in range(len(energies)):
                assert np.allclose(steps.tvib[m, j, l, :n], r["tvib"][m][j], **tol), (l, m, j)
        assert np.allclose(dfrac[l, :n], np.array(r["dfrac"]), rtol=1e-9, atol=1e-14)
        assert np.all(steps.column[:, l, n:] == 0.0) and np.all(steps.temp[l, n:] == 100.0)


def test_radtrans_group_observations(world):
    """radtrans(group_observations=True): the pixels' spectra interpolated from a ladder of
    simulated LOS (smm:3056-3058, 3263-3272) agree with the per-pixel simulation to the accuracy of
    the quadratic spline on a 12 km ladder, and equal the literal spline + FOV recomputation."""
    smm, S, planet, sp = world["smm"], world["S"], world["planet"], world["sp"]
    centres = np.linspace(2997.0, 3003.0, 7)
    widths = np.full(7, 0.6)
    inputs = dict(n_split=1, cart_LUTS=None, out_dir=None, n_threads=8)
    LUTopt = dict(pres_step_log=1.0, temp_step=5.0)
    opt = dict(max_T_variation=5., max_Plog_variation=1.)
    mk = lambda: S.vims_pixels([450.0, 520.0, 700.0], channels=centres, widths=widths)   # noqa: E731
    sims, rt, _ = smm.radtrans(inputs, planet, world["lines"], mk(), sp_gri=sp, radtran_opt=opt,
                               LUTopt=dict(LUTopt))
    pix_g = mk()
    sims_g, rt_g, _ = smm.radtrans(inputs, planet, world["lines"], pix_g, sp_gri=sp, radtran_opt=opt,
                                   LUTopt=dict(LUTopt), group_observations=True, alt_step_sims=12.,
                                   alt_first_los=400.)
    assert len(sims_g) == 3 and len(rt_g) > 20
    for a, b in zip(sims, sims_g):
        assert rel_err(b.spectrum, a.spectrum) < 1e-2
    pix_s = sorted(pix_g, key=lambda p: p.limb_tg_alt)
    loss, alts, _, _ = smm.make_group_observations(pix_s, alt_step=12., alt_first_los=400.)
    spl = smm.make_radtran_spline(alts, [rt_g['LOS%02d' % i] for i in range(len(alts))])
    three = np.array([spl(al).spectrum for al in smm._pixel_los_altitudes(pix_s[1])])
    assert np.allclose(sims_g[1].spectrum, smm.fov_integrate(three, 0.0), rtol=1e-13)


def test_inversion_fast_limb_group_observations(world):
    """The ladder variant of the forward + Jacobian evaluation (smm:2904-2929): radiances and
    derivative spectra interpolated in altitude agree with the per-pixel evaluation to the accuracy
    of the spline, and the Jacobian keeps its shape."""
    smm, S = world["smm"], world["S"]
    planet = S.titan_planet(world["tab"]["level_energies"], nonlte=False)
    sp = world["sp"]
    centres = np.linspace(2997.0, 3003.0, 7)
    widths = np.full(7, 0.6)
    inputs = dict(n_split=1, cart_LUTS=None, out_dir=None, n_threads=8)
    LUTopt = dict(pres_step_log=1.0, temp_step=5.0)
    opt = dict(max_T_variation=5., max_Plog_variation=1.)
    mk = lambda: S.vims_pixels([470.0, 650.0], channels=centres, widths=widths)   # noqa: E731
    bs_a, bs_b = _bayes(smm, planet), _bayes(smm, planet)
    sims_a, _, _ = smm.forward_jacobian_limb(inputs, planet, world["lines"], bs_a, mk(), sp_gri=sp,
                                           radtran_opt=opt, LUTopt=dict(LUTopt))
    sims_b, rt_b, _ = smm.forward_jacobian_limb(inputs, planet, world["lines"], bs_b, mk(), sp_gri=sp,
                                              radtran_opt=opt, LUTopt=dict(LUTopt),
                                              group_observations=True, alt_step_sims=12.)
    assert len(rt_b) > 15
    Ja, Jb = bs_a.build_jacobian(), bs_b.build_jacobian()
    assert Ja.shape == Jb.shape == (14, bs_a.n_tot)
    for a, b in zip(sims_a, sims_b):
        assert rel_err(b.spectrum, a.spectrum) < 1e-2
    scale = np.abs(Ja).max(axis=0, keepdims=True)
    assert np.all(np.abs(Jb - Ja) <= 2e-2 * scale)


def test_convolve_lowres_many_spectra(world, oracle):
    """More spectra than one launch pair of k_convolve_lowres takes (slabs of 512): every spectrum
    gets the same channels as when it is convolved alone."""
    eng, torch = world["engine"], world["torch"]
    rng = np.random.default_rng(12)
    x = np.arange(3000.0, 3004.0, 5e-4)
    y = rng.uniform(0.0, 1.0, (700, len(x)))
    centres = np.linspace(3000.5, 3003.5, 9)
    widths = np.full(9, 0.2)
    xd, yd = torch.as_tensor(x, device="cuda"), torch.as_tensor(y, device="cuda")
    all_ = eng.convolve_lowres(xd, yd, centres, widths).cpu().numpy()
    for i in (0, 511, 512, 699):
        one = eng.convolve_lowres(xd, yd[i:i + 1].contiguous(), centres, widths).cpu().numpy()
        assert np.array_equal(all_[i], one[0]), i
    for i in (3, 600):
        ref = oracle.convolve_to_grid_from_irregular(x, y[i], centres, widths)
        assert rel_err(all_[i], ref, floor_rel=1e-12) < 1e-12


def test_device_step_builder_latitude_box_parameters(world):
    """Parameters with latitude boxes (LinearProfile_2D): the device step builder's derivative
    table against the per-LOS host builder on the 7-band atmosphere, for rays that cross bands."""
    smm, S, sbm = world["smm"], world["S"], world["sbm"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=7)
    atm = planet.atmosphere
    starts = list(atm.grid.coords['lat'][:-1])
    nodes = [300., 550., 800., 1100.]
    vals = [[0.015 * (1 + 0.1 * b)] * 4 for b in range(7)]
    prof = smm.LinearProfile_2D('CH4', atm, nodes, starts, vals, [[1e-3] * 4] * 7)
    bs = smm.BayesSet('3d')
    bs.add_set(prof)
    p = prof.profile()
    p.values['vmr'][6] = p.values['vmr'][5]          # keep the (quirky, empty) last box physical
    planet.gases['CH4'].add_clim(p)
    obs = sbm.Coords([20.0, 90.0, 1.0e5], s_ref='Spherical')
    loss = [sbm.LineOfSight(obs, sbm.Coords([lat, 2.0, alt], s_ref='Spherical'))
            for alt, lat in ((400.0, 58.0), (650.0, -31.0), (900.0, 74.0), (500.0, 10.0))]
    opt = dict(max_T_variation=5.0, max_Plog_variation=1.0)
    gi, steps, dfrac = smm.los_step_tables_device(loss, planet, bayes_set=bs, set_name='CH4', **opt)
    host = []
    for los in loss:
        l2 = sbm.LineOfSight(los.starting_point, los.second_point)
        l2.calc_atm_intersections(planet)
        l2.calc_radtran_steps(planet, None, calc_derivatives=True, bayes_set=bs, **opt)
        host.append(l2)
    _, steps_h = smm.los_step_tables(host, planet)
    nmax = steps_h.n_steps_max
    assert np.array_equal(steps.n_steps, steps_h.n_steps)
    assert np.allclose(steps.column[:, :, :nmax], steps_h.column, rtol=1e-10, atol=0)
    dfrac_h = smm.los_jac_tables(host, bs, 'CH4', nmax)
    assert np.allclose(dfrac[:, :nmax], dfrac_h, rtol=1e-9, atol=1e-14)
    used_boxes = {bs.params()[q].key[0] for q in range(bs.n_tot) if np.any(dfrac[:, :, q] != 0.0)}
    assert len(used_boxes) >= 3                        # the rays cross several latitude boxes
    for l, los in enumerate(host):
        for par in bs.params():
            assert loss[l].involved_retparams[(par.nameset, par.key)] == \
                los.involved_retparams[(par.nameset, par.key)]
