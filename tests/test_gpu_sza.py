"""GPU tests of the 3-D part of the LOS path: solar-zenith-
angle dependent vibrational temperatures along the LOS, use_tangent_sza, LOS_order='photon' /
invert_LOS_direction, max_opt_depth, single_rads / track_levels, an nm observation, and the
retrieval loop of inversion_fast_limb.  The reference's own code for these lives in the missing
spect_base_module (DESIGN.md section 6 is the specification), so the checks are: device builder ==
host builder == oracle restatement, and physical invariants."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    import torch
    from spectrobot_b200 import engine, spect_base_module as sbm, spect_classes as spcl
    from spectrobot_b200 import spect_main_module as smm, synthetic as S
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    tab = S.line_table(140, 2993.0, 3007.0, n_levels=5, seed=33, frac_unlinked=0.0)
    lines = S.spect_lines(tab)
    sp = smm.prepare_spe_grid([2996.0, 3004.0]).spectral_grid
    return dict(torch=torch, engine=engine, sbm=sbm, spcl=spcl, smm=smm, S=S, tab=tab, lines=lines,
                sp=sp)


def _host_steps(world, planet, loss, ssps=None, fszas=None, LOS_order='radtran', lines=None, **opt):
    sbm, smm = world["sbm"], world["smm"]
    host = []
    for i, los in enumerate(loss):
        l2 = sbm.LineOfSight(los.starting_point, los.second_point)
        l2.calc_atm_intersections(planet, LOS_order=LOS_order)
        if fszas is not None:
            l2.szas = np.ones(len(l2.intersections)) * fszas[i]
        elif ssps is not None:
            l2.calc_SZA_along_los(planet, ssps[i])
        l2.calc_radtran_steps(planet, lines, **opt)
        host.append(l2)
    return smm.los_step_tables(host, planet)


def _assert_same_tables(a, b, tol=1e-10):
    n = b.n_steps_max
    assert np.array_equal(a.n_steps, b.n_steps)
    kw = dict(rtol=tol, atol=0)
    assert np.allclose(a.temp[:, :n], b.temp, **kw) and np.allclose(a.pres[:, :n], b.pres, **kw)
    assert np.allclose(a.column[:, :, :n], b.column, **kw)
    assert np.allclose(a.tvib[:, :, :, :n], b.tvib, **kw)


def test_sza_along_los_device_host_oracle(world, oracle):
    """T_vib(lat, SZA, z) evaluated at the SZA of every sample of the ray (calc_SZA_along_los), or
    at the tangent SZA everywhere (use_tangent_sza): sr_los_steps_build_rays == the per-LOS host
    methods == the oracle restatement; and the two options really differ."""
    S, sbm, smm, eng = world["S"], world["sbm"], world["smm"], world["engine"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=7, sza_nodes=S.SZA_NODES)
    pixels = S.vims_pixels([380.0, 610.0, 905.0], lat=25.0, sza=[35.0, 58.0, 77.0])
    loss, ssps, fszas = [], [], []
    for p in pixels:
        loss += [p.low_LOS(), p.LOS(), p.up_LOS()]
        ssps += 3 * [p.sub_solar_point()]
        fszas += 3 * [p.limb_tg_sza]
    opt = dict(max_T_variation=5.0, max_Plog_variation=1.0)
    gi, st_sun, _ = smm.los_step_tables_device(loss, planet, ssps=ssps, fszas=fszas, **opt)
    _assert_same_tables(st_sun, _host_steps(world, planet, loss, ssps=ssps, **opt)[1])
    _, st_fix, _ = smm.los_step_tables_device(loss, planet, ssps=ssps, fszas=fszas,
                                              use_tangent_sza=True, **opt)
    _assert_same_tables(st_fix, _host_steps(world, planet, loss, fszas=fszas, **opt)[1])
    assert np.array_equal(st_sun.temp, st_fix.temp)                       # geometry is the same
    excited = world["tab"]["level_energies"] > 0
    d = np.abs(st_sun.tvib - st_fix.tvib)[0][excited]
    assert d.max() > 0.5                                                  # ... T_vib is not (K)
    with pytest.raises(ValueError):
        smm.los_step_tables_device(loss, planet, **opt)                   # 3-D T_vib needs the Sun
    # oracle restatement on the same tables
    atm = smm.planet_atmosphere_tables(planet, gi)
    org = np.array([l.starting_point.Cartesian() for l in loss])
    drc = np.array([l.direction for l in loss])
    sun = np.array([p.Cartesian() for p in ssps])
    ref = oracle.los_steps_build(atm.z, atm.temp, atm.pres, atm.vmr, org, drc, tvib=atm.tvib,
                                 tvib_on=atm.tvib_on, lat_edges=atm.lat_edges, sza_nodes=atm.sza_nodes,
                                 sun=sun, **opt)
    for l, r in enumerate(ref):
        n = r["n_steps"]
        assert n == st_sun.n_steps[l]
        assert np.allclose(st_sun.temp[l, :n], r["temp"], rtol=1e-10, atol=0)
        assert np.allclose(st_sun.column[0, l, :n], r["column"][0], rtol=1e-10, atol=0)
        for j in range(atm.n_sets_max):
            assert np.allclose(st_sun.tvib[0, j, l, :n], r["tvib"][0][j], rtol=1e-10, atol=0)
    # a plain engine call with one Sun vector for the whole batch
    st_one, _ = eng.los_steps_build(atm, org, drc, sun=sun[0], **opt)
    assert np.array_equal(st_one.tvib[:, :, :3], st_sun.tvib[:, :, :3])


def test_photon_order_and_opt_depth_limit(world, oracle):
    """LOS_order='photon' reverses the samples (steps are rebuilt from the observer's side);
    max_opt_depth adds the optical-depth merge limit: device == host == oracle for both."""
    S, sbm, smm = world["S"], world["sbm"], world["smm"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=7, sza_nodes=S.SZA_NODES)
    pix = S.vims_pixels([300.0, 520.0, 840.0], lat=-40.0, sza=[44.0, 61.0, 72.0])
    loss = [p.LOS() for p in pix]
    ssps = [p.sub_solar_point() for p in pix]
    opt = dict(max_T_variation=5.0, max_Plog_variation=1.0)
    gi, st_r, _ = smm.los_step_tables_device(loss, planet, ssps=ssps, **opt)
    _, st_p, _ = smm.los_step_tables_device(loss, planet, ssps=ssps, LOS_order='photon', **opt)
    _assert_same_tables(st_p, _host_steps(world, planet, loss, ssps=ssps, LOS_order='photon', **opt)[1])
    for l in range(len(loss)):   # same total columns, opposite order of the layers
        n_r, n_p = st_r.n_steps[l], st_p.n_steps[l]
        assert st_r.column[0, l, :n_r].sum() == pytest.approx(st_p.column[0, l, :n_p].sum(), rel=1e-9)
        assert n_r > 2 and n_p > 2
    # optical-depth limit: thinner steps, device == host == oracle
    sig = sbm.peak_cross_sections(planet, world["lines"])['CH4']
    tau_tot = sig * st_r.column[0].sum(axis=1)
    lim = float(tau_tot.max()) / 40.0
    _, st_t, _ = smm.los_step_tables_device(loss, planet, ssps=ssps, max_opt_depth=lim,
                                            lines=world["lines"], **opt)
    assert st_t.n_steps.sum() > st_r.n_steps.sum()
    _assert_same_tables(st_t, _host_steps(world, planet, loss, ssps=ssps, lines=world["lines"],
                                          max_opt_depth=lim, **opt)[1])
    atm = smm.planet_atmosphere_tables(planet, gi)
    ref = oracle.los_steps_build(atm.z, atm.temp, atm.pres, atm.vmr,
                                 np.array([l.starting_point.Cartesian() for l in loss]),
                                 np.array([l.direction for l in loss]), tvib=atm.tvib,
                                 tvib_on=atm.tvib_on, lat_edges=atm.lat_edges, sza_nodes=atm.sza_nodes,
                                 sun=np.array([p.Cartesian() for p in ssps]), max_opt_depth=lim,
                                 sigma_peak=[sig], **opt)
    assert [r["n_steps"] for r in ref] == list(st_t.n_steps)
    with pytest.raises(ValueError):
        smm.los_step_tables_device(loss, planet, ssps=ssps, max_opt_depth=lim, **opt)   # no lines


def test_radtrans_3d_options_and_single_rads(world, oracle, tmp_path):
    """smm.radtrans on a 3-D non-LTE planet: use_tangent_sza and invert_LOS_direction change the
    result (they are honoured, not ignored), the radiances agree with the oracle LOS integral run
    on the same step tables, single_rads / track_levels contributions add up to the total, and
    save_hires writes the hi-res spectra the low-res ones were convolved from."""
    import pickle
    S, smm = world["S"], world["smm"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=7, sza_nodes=S.SZA_NODES)
    centres, widths = np.linspace(2997.0, 3003.0, 7), np.full(7, 0.6)
    inputs = dict(n_split=1, cart_LUTS=None, out_dir=str(tmp_path), n_threads=8)
    LUTopt = dict(pres_step_log=1.0, temp_step=5.0)
    opt = dict(max_T_variation=5., max_Plog_variation=1.)
    mk = lambda: S.vims_pixels([520.0, 760.0], lat=20.0, sza=[38.0, 74.0], channels=centres,  # noqa: E731
                               widths=widths)
    kw = dict(sp_gri=world["sp"], radtran_opt=opt, LUTopt=LUTopt, save_hires=False)
    trk = {('CH4', 'iso_1'): ['lev_01', 'lev_03']}
    sims, rt, single = smm.radtrans(inputs, planet, world["lines"], mk(), track_levels=trk, **kw)
    sims_t, rt_t, _ = smm.radtrans(inputs, planet, world["lines"], mk(), use_tangent_sza=True, **kw)
    sims_i, rt_i, _ = smm.radtrans(inputs, planet, world["lines"], mk(), invert_LOS_direction=True, **kw)
    a = np.array([s.spectrum for s in sims])
    assert np.all(np.isfinite(a)) and a.max() > 0
    assert rel_err(np.array([s.spectrum for s in sims_t]), a) > 1e-4      # SZA gradient matters
    d_inv = rel_err(np.array([s.spectrum for s in sims_i]), a)
    assert d_inv > 1e-9        # reversed layers: the observer now looks at the other hemisphere
    # emitters add up: sum over levels == the isotopologue == the total (one gas here)
    tot = np.array([rt[t].spectrum for t in sorted(rt)])
    iso = np.array([single[('CH4', 'iso_1')][t].spectrum for t in sorted(rt)])
    assert rel_err(iso, tot) < 1e-12
    allev = smm.track_all_levels(planet)
    _, rt2, single2 = smm.radtrans(inputs, planet, world["lines"], mk(), track_levels=allev, **kw)
    lev_sum = sum(np.array([single2[('CH4', 'iso_1', lev)][t].spectrum for t in sorted(rt2)])
                  for lev in allev[('CH4', 'iso_1')])
    assert rel_err(lev_sum, tot) < 1e-7      # (float32 layer scratch of the low-res sink: ~1e-9 per run)
    assert np.all(np.array([single[('CH4', 'iso_1', 'lev_01')][t].spectrum for t in sorted(rt)]) >= 0)
    # hi-res on disk + oracle LOS integral on the same step tables
    kw["save_hires"] = True
    sims_h, rt_h, _ = smm.radtrans(inputs, planet, world["lines"], mk(), nome_inv='hr', **kw)
    assert rel_err(np.array([s.spectrum for s in sims_h]), a) < 1e-7    # hi-res path: FP64 layers; low-res sink: float32
    nsp, hires = pickle.load(open(str(tmp_path / 'hires_radtran_hr.pic'), 'rb'))
    assert nsp == 0 and sorted(hires) == sorted(rt_h)
    low = oracle.convolve_to_grid_from_irregular(world["sp"].grid, hires['LOS01'].spectrum, centres, widths)
    assert rel_err(rt_h['LOS01'].spectrum, low) < 1e-7
    pix = sorted(mk(), key=lambda p: p.limb_tg_alt)
    loss = [pix[0].low_LOS(), pix[0].LOS(), pix[0].up_LOS()]
    gi, steps, _ = smm.los_step_tables_device(loss, planet, ssps=[pix[0].sub_solar_point()] * 3, **opt)
    max_p = max(planet.atmosphere.calc(p.low_LOS().get_tangent_point(), 'pres') for p in pix)
    PT = smm.calc_PT_couples_atmosphere(world["lines"], list(planet.gases.values()), planet.atmosphere,
                                        max_pres=max_p, **LUTopt)
    LUTS = smm.check_and_build_allluts(inputs, world["sp"], world["lines"], list(planet.gases.values()),
                                       PTcouples=PT, LUTopt=LUTopt)
    L = LUTS[('CH4', 1)]
    olut = dict(g32=L.g32.cpu().numpy(), pt=np.array(L.PTcouples),
                level_energy=world["tab"]["level_energies"], mol=6, iso=1, iso_ratio=S.CH4_RATIO,
                lte_unidentified=False)
    want = oracle.los_rt([olut], steps.n_steps, steps.temp, steps.pres, steps.column, steps.tvib)
    for i, tag in enumerate(('LOS00', 'LOS01', 'LOS02')):
        assert rel_err(hires[tag].spectrum, want[i]) < 1e-5


def test_nm_observation_through_radtrans(world, oracle):
    """An observation on a wavelength axis in nm and in Wm2 (the reference's VIMS pixels): radtrans
    converts per point and convolves on the wavelength axis on the device, like
    SpectralIntensity.hires_to_lowres does (spcl:1180-1191, 771-797; pinned by
    tests/test_gpu_ref_golden.py::test_convolution_on_device)."""
    S, smm, spcl = world["S"], world["smm"], world["spcl"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=1)
    c_cm, w_cm = np.linspace(2997.0, 3003.0, 7), np.full(7, 0.6)
    c_nm = np.sort(1.e7 / c_cm)
    w_nm = np.full(7, 0.6 * 1.e7 / 3000.0 ** 2)
    inputs = dict(n_split=1, cart_LUTS=None, out_dir=None, n_threads=8)
    kw = dict(sp_gri=world["sp"], radtran_opt=dict(max_T_variation=5., max_Plog_variation=1.),
              LUTopt=dict(pres_step_log=1.0, temp_step=5.0), save_hires=False)
    pix_nm = S.vims_pixels([600.0], channels=c_nm, widths=w_nm, units='nm', obs_units='Wm2')
    sims, rt, _ = smm.radtrans(inputs, planet, world["lines"], pix_nm, **kw)
    assert sims[0].units == 'Wm2' and sims[0].spectral_grid.units == 'nm'
    los = pix_nm[0].LOS()
    los.calc_atm_intersections(planet)
    los.calc_radtran_steps(planet, None, max_T_variation=5., max_Plog_variation=1.)
    gases = list(planet.gases.values())
    max_p = planet.atmosphere.calc(pix_nm[0].low_LOS().get_tangent_point(), 'pres')
    PT = smm.calc_PT_couples_atmosphere(world["lines"], gases, planet.atmosphere, max_pres=max_p,
                                        pres_step_log=1.0, temp_step=5.0)
    LUTS = smm.check_and_build_allluts(inputs, world["sp"], world["lines"], gases, PTcouples=PT)
    hi = los.radtran_fast(world["sp"], planet, LUTS=LUTS)[0]
    x = world["sp"].grid
    want = oracle.convolve_to_grid_from_irregular((1.e7 / x)[::-1], (hi.spectrum * x ** 2 * 1e-7)[::-1],
                                                 c_nm, w_nm) * 1e-3
    assert rel_err(rt['LOS01'].spectrum, want) < 1e-7
    low = hi.hires_to_lowres(pix_nm[0].observation, spectral_widths=w_nm)
    assert rel_err(low.spectrum, want) < 1e-7 and low.units == 'Wm2'
    # micron axis: the same channels, per micron
    pix_um = S.vims_pixels([600.0], channels=c_nm * 1e-3, widths=w_nm * 1e-3, units='mum', obs_units='Wm2')
    _, rt_um, _ = smm.radtrans(inputs, planet, world["lines"], pix_um, **kw)
    assert rel_err(rt_um['LOS01'].spectrum, want * 1e3) < 1e-7
    pix_hz = S.vims_pixels([600.0], channels=c_cm * 3e10, widths=w_cm * 3e10, units='hz')
    with pytest.raises(ValueError):
        smm.radtrans(inputs, planet, world["lines"], pix_hz, **kw)


def test_inversion_fast_limb_retrieves_vmr(world):
    """The reference-shaped retrieval loop (smm:2598-2987): synthetic observations made with a
    known CH4 profile, a first guess 30 % off; inversion_fast_limb iterates forward model +
    analytic Jacobians + Levenberg-Marquardt steps, chi drops and the profile moves towards the
    truth.  solo_simulation returns None after one simulation."""
    S, smm, spcl = world["S"], world["smm"], world["spcl"]
    planet = S.titan_planet(world["tab"]["level_energies"], n_bands=1, nonlte=False)
    centres, widths = np.linspace(2996.6, 3003.4, 18), np.full(18, 0.35)
    inputs = dict(n_split=1, cart_LUTS=None, out_dir=None, n_threads=8)
    kw = dict(sp_gri=world["sp"], radtran_opt=dict(max_T_variation=5., max_Plog_variation=1.),
              LUTopt=dict(pres_step_log=1.0, temp_step=5.0))
    tg = [420.0, 560.0, 700.0, 840.0]
    nodes = [300., 500., 700., 900., 1100.]
    z = planet.atmosphere.grid.coords['alt']
    truth = [0.015] * 5

    def bayes(values):
        prof = smm.LinearProfile_1D_new('CH4', z, nodes, truth, [0.5 * v for v in truth],
                                        first_guess_prof=values)
        bs = smm.BayesSet('ch4')
        bs.add_set(prof)
        return bs

    pixels = S.vims_pixels(tg, channels=centres, widths=widths)
    sims_true, _, _ = smm.forward_jacobian_limb(inputs, planet, world["lines"], bayes(truth), pixels, **kw)
    for pix, sim in zip(sorted(pixels, key=lambda p: p.limb_tg_alt), sims_true):
        pix.observation.spectrum = sim.spectrum.copy()
        pix.observation.noise = spcl.SpectralObject(np.full(18, 0.01 * sim.spectrum.max()),
                                                    pix.observation.spectral_grid)
        pix.observation.mask = np.ones(18)
    guess = [0.7 * v for v in truth]
    assert smm.inversion_fast_limb(inputs, planet, world["lines"], bayes(guess), pixels,
                                   solo_simulation=True, **kw) is None
    bs = bayes(guess)
    first = smm.chicalc([p.observation for p in pixels],
                        smm.forward_jacobian_limb(inputs, planet, world["lines"], bayes(guess), pixels, **kw)[0],
                        [p.observation.noise for p in pixels], [p.observation.mask for p in pixels], 0)
    out = smm.inversion_fast_limb(inputs, planet, world["lines"], bs, pixels, max_it=6,
                                  lambda_LM=0.01, **kw)
    assert out is not None
    chi, obs, sims, bs_out = out
    assert chi < 0.05 * first
    got = np.array([p.value for p in bs_out.params()])
    assert np.abs(got[1:4] / 0.015 - 1.0).max() < 0.1          # nodes the four tangent heights see
    assert bs_out.av_kernel.shape == (5, 5) and len(bs_out.old_params) >= 1
    with pytest.raises(NotImplementedError):
        smm.inversion_fast_limb(inputs, planet, world["lines"], bayes(guess), pixels, save_hires=True, **kw)
