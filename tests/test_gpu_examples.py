"""GPU tests: the driver-shaped acceptance scripts under examples/ (Python-3 copies of the
reference's radtran_test_CO.py, radtran_3D_ch4.py and run_0607_lut*.py on synthetic inputs, using
only the reference's module / class / function names) run end to end, and their products agree
with the oracle pipeline (VERDICT row G; BASELINE.json configs[0..2])."""
import importlib
import os
import sys

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
EX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")


@pytest.fixture(scope="module")
def examples(tmp_path_factory):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    os.environ["SR_EXAMPLE_SMALL"] = "1"
    os.environ["SR_EXAMPLE_DIR"] = str(tmp_path_factory.mktemp("examples"))
    sys.path.insert(0, EX)
    yield lambda name: importlib.import_module(name)
    sys.path.remove(EX)
    for k in ("SR_EXAMPLE_SMALL", "SR_EXAMPLE_DIR"):
        os.environ.pop(k, None)


def test_run_0607_lut(examples, oracle):
    """check_and_build_allluts for every gas of the planet; two cells of every LUT against the
    oracle's whole-cell routine on the lines the script read from its HITRAN-format file."""
    from spectrobot_b200 import spect_classes as spcl, spect_main_module as smm
    planet, linee, allLUTS, wn_ranges, LUTopt = examples("run_0607_lut_py3").main(small=True)
    assert sorted(allLUTS) == [('CH4', 1), ('CH4', 2), ('HCN', 1)]
    for (name, iso), L in allLUTS.items():
        im = getattr(planet.gases[name], 'iso_%d' % iso)
        w0, w1 = wn_ranges[name]
        mine = [l for l in smm.check_lines_mols(linee, [planet.gases[name]])
                if l.Iso == iso and w0 <= l.Freq <= w1]
        tab = spcl.line_table(mine, im if len(im.levels) else None)
        g32 = L.g32.cpu().numpy()
        assert g32.shape[0] == len(L.PTcouples) > 10 and np.any(g32)
        for c in (0, len(L.PTcouples) - 1):
            P, T = L.PTcouples[c]
            want = oracle.gcoeff_cell(tab, L.spectral_grid.grid, T, P, im.MM, tab["n_sets"])
            assert rel_err(g32[c], want.astype(np.float32)) < 2e-7, (name, iso, c)
        assert L.LTE == (len(im.levels) == 0)


def test_radtran_3D_ch4(examples):
    """The 3-D retrieval driver: LUTs, SZA-dependent T_vib along the LOS, an nm / Wm2 observation,
    inversion_fast_limb(g3D=True): chi drops and the retrieved CH4 moves from the a-priori towards
    the profile the 'observations' were simulated with."""
    result, truth, sims_true, planet, linee, pixels = examples("radtran_3D_ch4_py3").main(small=True)
    assert result is not None
    chi, obs, sims, bayes = result
    assert np.isfinite(chi) and chi < 5.0
    assert all(s.units == 'Wm2' and s.spectral_grid.units == 'nm' for s in sims)
    used = [(p, t) for p, t in zip(bayes.params(), truth.params()) if p.is_used]
    assert len(used) >= 3
    # parameters the three tangent heights constrain: closer to the truth than the a-priori was
    err0 = np.array([abs(p.apriori / t.value - 1) for p, t in used])
    err1 = np.array([abs(p.value / t.value - 1) for p, t in used])
    assert np.median(err1) < 0.7 * np.median(err0)
    for a, b in zip(sims, sims_true):
        assert rel_err(a.spectrum, b.spectrum, 1e-3) < 0.1


def test_radtran_test_CO(examples, oracle):
    """The per-LOS driver (smm.inversion -> LineOfSight.radtran): runs with LUTs; one LOS of it
    without LUTs (useLUTs=False, line-by-line at every step) agrees with the LUT path within the
    LUT's interpolation error and with the oracle's layer recursion."""
    from spectrobot_b200 import spect_main_module as smm
    mod = examples("radtran_test_CO_py3")
    result, truth, sims_true, planet, linee, pixels = mod.main(small=True)
    chi, obs, sims, bayes = result
    assert np.isfinite(chi) and chi < 5.0
    got = np.array([p.value for p in bayes.params()])
    assert np.all(got > 0)
    err0 = np.abs(np.array([p.apriori for p in bayes.params()]) / 65.e-6 - 1)
    err1 = np.abs(got / 65.e-6 - 1)
    assert np.median(err1) < np.median(err0)
    # LUT vs line-by-line on one LOS
    pix = pixels[0]
    wn_range = [2140., 2150.]
    opt = dict(max_T_variation=5., max_Plog_variation=1.0)
    los_a, los_b = pix.LOS(), pix.LOS()
    with_lut = los_a.radtran(wn_range, planet, linee, radtran_opt=opt,
                             LUTopt=dict(temp_step=5., pres_step_log=1.0))[0]
    no_lut = los_b.radtran(wn_range, planet, linee, useLUTs=False, radtran_opt=opt)[0]
    scale = np.abs(no_lut.spectrum).max()
    assert np.abs(with_lut.spectrum - no_lut.spectrum).max() < 0.05 * scale
    assert np.abs(with_lut.spectrum - no_lut.spectrum).max() > 0.0


def test_radtran_3Dvs2D_radtrans_new(examples):
    """The 3-D vs 2-D driver (BASELINE configs[3]) on a planet whose profiles are latitude-LINEAR
    (host step builder): radtrans with the altitude ladder, tracked levels and hi-res files in the
    three geometries (SZA along the LOS, tangent SZA, inverted LOS), then the three retrievals."""
    results, truth, sims_true, planet, linee, pixels = \
        examples("radtran_3Dvs2D_radtrans_new_py3").main(small=True)
    sims3, rt3, single3 = results['radtran_tracklevels_szavar_all']
    sims2, rt2, single2 = results['radtran_tracklevels_noszavar_short']
    simsi, rti, singlei = results['radtran_3D_tracklevels_inverseLOS_short']
    assert len(sims3) == len(pixels) == 3 and len(rt3) >= 8          # the ladder, not 3 LOS per pixel
    assert all(s.units == 'Wm2' and s.spectral_grid.units == 'nm' for s in sims3 + sims2 + simsi)
    tags = sorted(rt3)
    ch4_levels = planet.gases['CH4'].iso_1.levels
    assert set(single3) == set([('CH4', 'iso_1'), ('HCN', 'iso_1')] +
                               [('CH4', 'iso_1', l) for l in ch4_levels] +
                               [('HCN', 'iso_1', l) for l in planet.gases['HCN'].iso_1.levels])
    assert set(k for k in single2 if len(k) == 3) == set(
        [('CH4', 'iso_1', l) for l in ch4_levels[1:3]] + [('HCN', 'iso_1', planet.gases['HCN'].iso_1.levels[1])])
    for tag in (tags[1], tags[len(tags) // 2], tags[-2]):
        total = rt3[tag].spectrum
        gases = sum(single3[k][tag].spectrum for k in single3 if len(k) == 2)
        assert rel_err(gases, total, 1e-3) < 1e-4                     # the source is linear in the emitters
        levels = sum(single3[('CH4', 'iso_1', l)][tag].spectrum for l in ch4_levels)
        assert rel_err(levels, single3[('CH4', 'iso_1')][tag].spectrum, 1e-3) < 1e-4
    a3 = np.array([s.spectrum for s in sims3])
    a2 = np.array([s.spectrum for s in sims2])
    ai = np.array([s.spectrum for s in simsi])
    assert np.all(np.isfinite(a3)) and np.all(a3 > 0)
    assert 0.0 < np.abs(a2 - a3).max() < 0.5 * a3.max()               # the SZA along the LOS matters
    assert np.abs(ai - a3).max() > 0.0                                 # so does the direction
    out = os.environ["SR_EXAMPLE_DIR"]
    for teag in ('tracklevels_szavar_all', 'tracklevels_noszavar_short'):
        assert os.path.exists(os.path.join(out, 'out', 'hires_radtran_%s.pic' % teag))
        assert os.path.exists(os.path.join(out, 'out', 'radtran_%s.pic' % teag))
    for teag in ('2Dvs3D_szavar_lin', '2Dvs3D_noszavar_lin', '2Dvs3D_inverseLOS_lin'):
        chi, obs, sims, bayes = results['out_' + teag]
        assert np.isfinite(chi) and len(sims) == 3
        assert os.path.getsize(os.path.join(out, 'out', 'out_%s.pic' % teag)) > 0
    chi, obs, sims, bayes = results['out_2Dvs3D_szavar_lin']          # the geometry the truth was made with
    used = [(p, t) for p, t in zip(bayes.params(), truth.params()) if p.is_used]
    assert len(used) >= 3
    err0 = np.array([abs(p.apriori / t.value - 1) for p, t in used])
    err1 = np.array([abs(p.value / t.value - 1) for p, t in used])
    assert np.median(err1) < np.median(err0)


def test_inversion_sequences_20067(examples):
    """The sequence-retrieval driver (BASELINE configs[4] shape): every sequence retrieved with
    inversion_fast_limb(group_observations=True, alt_first_los=300., check_log=...), results and
    log written like the reference's."""
    import pickle
    num, sequences, results_tot, truth, planet = examples("inversion_sequences_20067_py3").main(small=True)
    assert num == 2 and len(sequences) == len(results_tot) == 2
    out = os.path.join(os.environ["SR_EXAMPLE_DIR"], 'out')
    log = open(os.path.join(out, 'check_log_allinv.dat')).read()
    assert log.count('SEQ: n_pix 3') == 2 and log.count('Iteration  0: chi is') == 2
    assert 'LATITUDE -20 - -10' in log and 'LATITUDE 10 - 20' in log and 'Fine!' in log
    with open(os.path.join(out, 'results_inversion_0607.pic'), 'rb') as f:
        n2, seq2, res2 = pickle.load(f)
    assert n2 == 2 and len(seq2[0]['pixels']) == 3
    for bayes, bayes2 in zip(results_tot, res2):
        got = np.array([p.value for p in bayes.params()])
        assert np.array_equal(got, np.array([p.value for p in bayes2.params()]))
        used = [(p, t) for p, t in zip(bayes.params(), truth.params()) if p.is_used]
        assert len(used) >= 3 and np.all(got > 0)
        err0 = np.array([abs(p.apriori / t.value - 1) for p, t in used])
        err1 = np.array([abs(p.value / t.value - 1) for p, t in used])
        assert np.median(err1) < np.median(err0)
    for k in (1, 2):
        assert os.path.getsize(os.path.join(out, 'out_inversion_0607_seq_%03d.pic' % k)) > 0
