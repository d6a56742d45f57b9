"""CPU test: the parameter space and the Levenberg-Marquardt update that smm.inversion_fast_limb
runs around the GPU forward model, against fixtures produced by EXECUTING the reference's own
LinearProfile_1D_new / RetParam / BayesSet / genvec / chicalc / inversion_algebra
(tests/golden/make_ref_golden3.py; smm:163-257, 450-656, 3399-3469)."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "ref_golden3.npz"))


def _case(ref):
    from spectrobot_b200 import spect_base_module as sbm, spect_classes as spcl, spect_main_module as smm
    z = ref["z"]
    alt_gri = sbm.AtmGrid('alt', z)
    bs = smm.BayesSet(tag='fixture')
    n1, n2 = np.arange(450., 1051., 100.), np.arange(550., 1051., 125.)
    bs.add_set(smm.LinearProfile_1D_new('CH4', alt_gri, n1, 0.015 + 1e-5 * n1, 0.5 * (0.015 + 1e-5 * n1)))
    bs.add_set(smm.LinearProfile_1D_new('HCN', alt_gri, n2, 2.e-6 + 0 * n2, 3.e-4 + 0 * n2,
                                        first_guess_prof=1.5e-6 + 0 * n2))
    grid = spcl.SpectralGrid(np.linspace(3280., 3320., ref["obs"].shape[1]), units='nm')
    obs = [spcl.SpectralIntensity(v, grid, units='Wm2') for v in ref["obs"]]
    sims = [spcl.SpectralIntensity(v, grid, units='Wm2') for v in ref["sims"]]
    noise = [spcl.SpectralObject(v, grid) for v in ref["noise"]]
    masks = [m for m in ref["masks"]]
    for q, par in enumerate(bs.params()):
        for k, d in enumerate(ref["derivs"][q]):
            par.store_deriv(spcl.SpectralIntensity(d, grid, units='Wm2'), num=k)
    return smm, bs, obs, sims, noise, masks


def test_parameter_space_matches_reference(ref):
    smm, bs, obs, sims, noise, masks = _case(ref)
    pars = bs.params()
    assert [p.key for p in pars] == list(ref["keys"]) and [p.nameset for p in pars] == list(ref["namesets"])
    assert np.array_equal(np.array([p.maskgrid.mask for p in pars if p.nameset == 'CH4']), ref["masks_par"])
    assert np.array_equal(bs.apriori_vector(), ref["apriori"])
    assert np.array_equal(bs.param_vector(), ref["values0"])
    assert np.array_equal(bs.VCM_apriori(), ref["vcm_apriori"])
    assert bs.n_tot == len(pars) == 12 and bs.order == ['CH4', 'HCN']
    assert np.array_equal(bs.build_jacobian(), ref["jac_nomask"])
    assert np.array_equal(bs.build_jacobian(masks=masks), ref["jac_mask"])
    pars[0].store_deriv(pars[0].derivatives[1], num=1)             # overwrite, not append (smm:651-656)
    assert len(pars[0].derivatives) == ref["derivs"].shape[1]


def test_levenberg_marquardt_step_matches_reference(ref):
    """genvec / chicalc bit-identical; two consecutive inversion_algebra steps (lambda 0.1, 3.0):
    parameters, averaging kernel and covariance at the rounding level of the matrix inverses, the
    halving rule of a parameter that would turn negative (RetParam.update_par, smm:633-641)."""
    smm, bs, obs, sims, noise, masks = _case(ref)
    assert np.array_equal(np.array(smm.genvec(obs, sims, noise, masks=masks)), ref["genvec"])
    assert smm.chicalc(obs, sims, noise, masks, 5) == ref["chi"][0]
    assert smm.chicalc(obs, sims, noise, None, 0) == ref["chi"][1]
    for tag, lam in (('a', 0.1), ('b', 3.0)):
        smm.inversion_algebra(obs, sims, noise, bs, lambda_LM=lam, masks=masks)
        # (bit-identical here; the tolerances leave room for another BLAS on an ill-conditioned
        # matrix: CH4 parameters ~1e-2, HCN ~1e-6)
        assert np.allclose(bs.param_vector(), ref["values_" + tag], rtol=1e-8, atol=0.0), tag
        assert np.allclose(bs.av_kernel, ref["avk_" + tag], rtol=1e-6, atol=1e-7), tag
        assert np.allclose(bs.VCM, ref["vcm_" + tag], rtol=1e-6, atol=0.0), tag
    assert np.all(bs.param_vector() > 0)
    assert np.allclose(np.array(bs.old_params), ref["old_params"], rtol=1e-8, atol=0.0)
    assert np.allclose(np.array(bs.params()[-1].old_values), ref["old_values_last"], rtol=1e-8, atol=0.0)
    # the last parameter's raw step was negative beyond its value: it was halved, not clipped
    assert bs.params()[-1].value < ref["values0"][-1]
    bs.update_parerror()
    assert bs.params()[3].ret_error == pytest.approx(np.sqrt(ref["vcm_b"][3, 3]), rel=1e-8)


def test_involvement_spline_and_level_filters_match_reference(ref):
    """LinearProfile_1D_new.check_involved (smm:492-501), make_radtran_spline (smm:3377-3396),
    check_lines_mols / track_all_levels / keep_levels_wlines / keep_levels (smm:69-160)."""
    from spectrobot_b200 import spect_base_module as sbm, spect_classes as spcl, spect_main_module as smm
    smm_, bs, *_ = _case(ref)
    ch4 = bs.sets['CH4']
    got = np.array([[ch4.check_involved(key, {'alt': list(r)}) for key in ch4.alts]
                    for r in ref["involved_ranges"]])
    assert np.array_equal(got, ref["involved"])
    grid = spcl.SpectralGrid(np.linspace(3280., 3320., 9), units='nm')
    ladder = [spcl.SpectralIntensity(v, grid, units='Wm2') for v in ref["spline_in"]]
    spl = smm.make_radtran_spline(ref["spline_alts"], ladder)
    got = np.array([spl(x).spectrum for x in ref["spline_probes"]])
    assert np.allclose(got, ref["spline_out"], rtol=1e-13, atol=0.0)
    assert np.allclose(got[0], ref["spline_in"][0], rtol=1e-12)            # interpolating at a node

    lines = spcl.read_line_database(os.path.join(GOLD, "ref_lines.par"))
    levels = ['0 0 0 0 1A1', '0 0 1 0 1F2', '0 1 0 0 1E', '1 0 0 0 1A1', '0 0 0 1 1F2']
    ch4m = sbm.Molec(6, 'CH4')
    ch4m.add_iso(1).add_levels(levels, [0.0, 3019.4935, 1533.3326, 2916.5, 1310.8])
    ch4m.add_iso(2)
    hcn = sbm.Molec(23, 'HCN')
    hcn.add_iso(1)
    planet = sbm.Titan(1500.)
    planet.gases = {'CH4': ch4m, 'HCN': hcn}
    ok = smm.check_lines_mols(lines, [ch4m, hcn])
    assert [l.Freq for l in ok] == list(ref["filter_lines_ok"])
    got = sorted('%s/%s/%s' % (g, i, lev) for (g, i), levs in smm.track_all_levels(planet).items() for lev in levs)
    assert got == list(ref["track_all"])
    smm.keep_levels_wlines(planet, lines)
    assert ch4m.iso_1.levels == list(ref["levels_wlines"])
    smm.keep_levels(planet, {('CH4', 'iso_1'): ['lev_00', 'lev_02'], ('CH4', 'iso_2'): [], ('HCN', 'iso_1'): []})
    assert ch4m.iso_1.levels == list(ref["levels_kept"])
    ok = smm.check_lines_mols(lines, [ch4m])
    assert [l.Freq for l in ok] == list(ref["filter_lines_ok2"])
    with pytest.raises(KeyError):
        smm.keep_levels(planet, {})                                          # like the reference (:144)
