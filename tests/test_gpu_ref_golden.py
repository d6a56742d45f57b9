"""GPU tests: the CUDA library against fixtures produced by EXECUTING the reference's own Python
(tests/golden/make_ref_golden.py): cross-sections within 1e-6, coefficient assembly and
convolution far inside the 1e-5 radiance tolerance.  See tests/test_ref_golden.py for what the
fixtures pin and what they cannot (the Fortran itself, the missing LOS module)."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CTYPES = ['sp_emission', 'ind_emission', 'absorption']
LEVELS = ['0 0 0 0 1A1', '0 0 1 0 1F2', '0 1 0 0 1E']
ENERGIES = [0.0, 3019.4935, 1533.3326]
TOL_XS = 1e-6


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "ref_golden.npz"))


@pytest.fixture(scope="module")
def case():
    import torch
    from spectrobot_b200 import engine, spect_base_module as sbm, spect_classes as spcl
    from spectrobot_b200 import spect_main_module as smm
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    lines = spcl.read_line_database(os.path.join(GOLD, "ref_lines.par"))
    iso1 = sbm.IsoMolec(6, 1)
    iso1.add_levels(LEVELS, ENERGIES)
    iso1.is_in_LTE = False
    iso2 = sbm.IsoMolec(6, 2)
    return dict(lines=lines, iso1=iso1, iso2=iso2, spcl=spcl, sbm=sbm, smm=smm, engine=engine,
                torch=torch)


def test_cuda_cells_match_reference_built_lut(ref, case):
    """K1/K2 (sr_gcoeff_cells_dev) against the LUT the reference built and pickled: non-LTE
    isotopologue (3 levels x 3 ctypes, two cells incl. a Lorentz-dominated one) and LTE 'all'."""
    spcl, engine = case["spcl"], case["engine"]
    grid = ref["grid"]
    l1 = [l for l in case["lines"] if l.Iso == 1]
    tab = spcl.line_table(l1, case["iso1"])
    ls = engine.LineSet(tab, grid, case["iso1"].MM, 3)
    got = ls.gcoeff_cells(ref["cells_PT"]).cpu().numpy()
    want = ref["cells_nonlte"]
    for s in range(3):
        for k in range(3):
            if not np.any(want[s, k]):
                assert not np.any(got[0, s, k]), (s, k)
            else:
                assert rel_err(got[0, s, k], want[s, k]) < TOL_XS, (s, k)
    assert rel_err(got[1][:, :, ::8], ref["cells_nonlte_b8"]) < TOL_XS
    g32 = ls.gcoeff_cells_f32(ref["cells_PT"][:1]).cpu().numpy()
    assert rel_err(g32[0], want.astype(np.float32)) < 2e-7
    l2 = [l for l in case["lines"] if l.Iso == 2]
    ls2 = engine.LineSet(spcl.line_table(l2, None), grid, case["iso2"].MM, 1)
    got = ls2.gcoeff_cells([ref["cells_lte_PT"]]).cpu().numpy()[0, 0]
    assert rel_err(got, ref["cells_lte"]) < TOL_XS


def test_reference_shaped_api_matches_reference(ref, case):
    """calc_shapes_lines (shapes, window position, G coefficients) and LookUpTable.make /
    LutSet.add_PT through the product's reference-shaped classes."""
    spcl, smm = case["spcl"], case["smm"]
    sp = smm.prepare_spe_grid([2998.0, 3006.0]).spectral_grid
    assert np.array_equal(sp.grid, ref["grid"])
    P, T = ref["cells_PT"][0]
    proc = spcl.calc_shapes_lines(sp, [l for l in case["lines"] if l.Iso == 1], T, P, case["iso1"])
    assert [l.Freq for l in proc] == list(ref["shape_freq"])
    for i, l in enumerate(proc):
        assert l.shape.spectral_grid.grid[0] == ref["shape_first"][i]
        for k, ct in enumerate(CTYPES):
            assert l.G_coeffs[ct] == pytest.approx(ref["shape_gcoeff"][i, k], rel=1e-12)
    for j, i in enumerate(ref["shape_pick"]):
        assert rel_err(proc[int(i)].shape.spectrum, ref["shape_spectra"][j]) < TOL_XS
    st = smm.LutSet(6, 1, case["iso1"].MM, level=case["iso1"].lev_01)
    set_ = st.add_PT(sp, proc, P, T, keep_memory=True)
    for k, ct in enumerate(CTYPES):
        assert rel_err(set_[ct].spectrum, ref["cells_nonlte"][1, k]) < TOL_XS
    lut = smm.LookUpTable(case["iso1"], [2998.0, 3006.0], LTE=False)
    lut.make(sp, case["lines"], [list(pt) for pt in ref["cells_PT"]])
    g32 = lut.g32.cpu().numpy()
    assert rel_err(g32[0], ref["cells_nonlte"].astype(np.float32)) < 2e-7


def test_abscoeff_and_lut_interpolation_on_device(ref, case):
    """A10 + A11 on the device: sr_los_abs_emi_dev (k_step_weights + k_los_mma) against the
    reference's make_abscoeff_LUTS_fast, non-LTE and LTE populations; all-zero rows = None."""
    engine = case["engine"]
    g32 = engine.lut_from_host(ref["interp_g32"])
    lut = engine.Lut(g32, ref["interp_PT"], 6, 1, 1.0, level_energies=np.array(ENERGIES))
    temps, press = ref["interp_probes"][:, 1], ref["interp_probes"][:, 0]
    n = len(temps)
    for tv, want in ((ref["abscoeff_tvib"], ref["abscoeff_nonlte"]), (None, ref["abscoeff_lte"])):
        tvib = None if tv is None else tv[None, :, None, :]
        steps = engine.LosSteps([n], temps[None], press[None], np.ones((1, 1, n)), tvib)
        a, e = engine.los_abs_emi([lut], steps)
        assert rel_err(a.cpu().numpy()[0], want[0], 1e-9) < 1e-9
        assert rel_err(e.cpu().numpy()[0], want[1]) < 1e-9


def test_convolution_on_device(ref, case):
    """f1: k_convolve_lowres against the reference's hires_to_lowres, observation in cm-1 and in
    nm (grid reversal + radiance Jacobian + intensity units, spcl:771-797, 1180-1191)."""
    engine, spcl = case["engine"], case["spcl"]
    got = engine.convolve_lowres_host(ref["conv_grid"], ref["conv_spec"], ref["conv_cm_centres"],
                                      ref["conv_cm_widths"])[0]
    assert rel_err(got, ref["conv_cm_result"]) < 1e-11
    hires = spcl.SpectralIntensity(ref["conv_spec"], spcl.SpectralGrid(ref["conv_grid"], units='cm_1'),
                                   units='ergscm2')
    obs = spcl.SpectralIntensity(np.zeros(len(ref["conv_nm_centres"])),
                                 spcl.SpectralGrid(ref["conv_nm_centres"], units='nm'), units='Wm2')
    low = hires.hires_to_lowres(obs, spectral_widths=ref["conv_nm_widths"])
    assert [low.units, low.spectral_grid.units] == list(ref["conv_nm_units"])
    assert rel_err(low.spectrum, ref["conv_nm_result"]) < 1e-11


def test_split_file_round_trip(ref, case, tmp_path):
    """f3: the reference-written split file loads into a resident table; the product's own
    split_and_compress_LUTS writes files the same reader (and the fixture comparison) accepts."""
    smm = case["smm"]
    sp = smm.prepare_spe_grid([2998.0, 3006.0]).spectral_grid
    lut = smm.LookUpTable(case["iso1"], [2998.0, 3006.0], LTE=False)
    lut.make(sp, case["lines"], [list(pt) for pt in ref["cells_PT"]])
    full = lut.g32.cpu().numpy()
    allL, n_split, grids = smm.split_and_compress_LUTS(sp, {('CH4', 1): lut}, str(tmp_path), 2, n_split=3)
    assert [len(g.grid) for g in grids] == list(ref["split_lens"])
    mine = smm.read_split_file(lut.splitfiles[1])
    theirs = smm.read_split_file(os.path.join(GOLD, "ref_LUT_csplit01_mol06_iso1_nonLTE.pic"))
    for lev in theirs:
        for c in range(2):
            for ct in CTYPES:
                a, b = mine[lev].sets[c][ct], theirs[lev].sets[c][ct]
                assert (a is None) == (b is None)
                if a is not None:
                    assert a.spectrum.dtype == np.float32
                    assert rel_err(a.spectrum, b.spectrum) < 3e-7
    lut.splitfiles[1] = os.path.join(GOLD, "ref_LUT_csplit01_mol06_iso1_nonLTE.pic")
    lut.load_split(1)
    lo = int(ref["split_lens"][0])
    assert rel_err(lut.g32.cpu().numpy(), full[..., lo:lo + int(ref["split_lens"][1])]) < 3e-7
