"""GPU tests of the reference-shaped callers around the hot path against fixtures produced by
EXECUTING the reference's own Python (tests/golden/make_ref_golden2.py): make_abscoeff_LUTS_fast
with tracked levels (AbsSetLOS in memory and streamed to disk), the slow line-by-line twin
make_abscoeff_isomolec, PrepareCalcShapes / do_for_th_calc, LutSet.make."""
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
def ref2():
    return np.load(os.path.join(GOLD, "ref_golden2.npz"))


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


class Box(object):
    def put(self, item):
        self.item = item


def _synthetic_lut(ref, case):
    """The small synthetic LUT of the A10/A11 fixtures as a resident LookUpTable."""
    smm, spcl, engine = case["smm"], case["spcl"], case["engine"]
    g32 = ref["interp_g32"]
    sg = spcl.SpectralGrid(np.arange(g32.shape[3]) * 5e-4 + 3000.0, units='cm_1')
    lut = smm.LookUpTable(case["iso1"], sg.wn_range(), False)
    lut.spectral_grid = sg
    lut.PTcouples = [list(map(float, pt)) for pt in ref["interp_PT"]]
    lut.g32 = engine.lut_from_host(g32)
    for s, lev in enumerate(case["iso1"].levels):
        st = smm.LutSet(6, 1, case["iso1"].MM, level=getattr(case["iso1"], lev))
        st.PTcouples, st.spectral_grid, st._table = lut.PTcouples, sg, (lut, s)
        lut.sets[lev] = st
    return lut, sg


def test_abscoeff_tracked_levels_match_reference(ref, ref2, case, tmp_path):
    """make_abscoeff_LUTS_fast(track_levels=...) (smm:2134-2299): totals, per-level emission, the
    reference's tracked-absorption quirk, file names; in memory and streamed to cartDROP."""
    smm, iso1 = case["smm"], case["iso1"]
    lut, sg = _synthetic_lut(ref, case)
    temps, press = ref["interp_probes"][:, 1], ref["interp_probes"][:, 0]
    for s, lev in enumerate(iso1.levels):
        getattr(iso1, lev).local_vibtemp = list(ref["abscoeff_tvib"][s])
    allL = {(iso1.mol_name, iso1.iso): lut}
    track = [str(v) for v in ref2["track_levels"]]
    drop = str(tmp_path) + '/'
    a, e, et, at = smm.make_abscoeff_LUTS_fast(sg, iso1, temps, press, LTE=False, allLUTs=allL,
                                               cartDROP=drop, track_levels=track)
    assert isinstance(a, smm.AbsSetLOS) and a.counter == len(temps) and len(a.set) == len(temps)
    assert rel_err(np.array([x.spectrum for x in a.set]), ref["abscoeff_nonlte"][0], 1e-9) < 1e-9
    assert rel_err(np.array([x.spectrum for x in e.set]), ref["abscoeff_nonlte"][1]) < 1e-9
    for j, lev in enumerate(track):
        assert rel_err(np.array([x.spectrum for x in et[lev].set]), ref2["track_emi"][j]) < 1e-9
        assert rel_err(np.array([x.spectrum for x in at[lev].set]), ref2["track_abs"][j]) < 1e-9
    # the tracked emissions of all emitting levels add up to the total emission
    tot = sum(np.array([x.spectrum for x in et[lev].set]) for lev in track)
    assert rel_err(tot, ref["abscoeff_nonlte"][1]) < 1e-9
    names = [os.path.basename(f) for f in (a.filename, e.filename, et['lev_01'].filename,
                                           at['lev_02'].filename)]
    assert names == [str(v) for v in ref2["track_names"]]
    assert not os.listdir(drop)                       # nothing written unless store_in_memory
    # streamed to disk: the reference's AbsSetLOS protocol reads it back
    a2, e2 = smm.make_abscoeff_LUTS_fast(sg, iso1, temps, press, LTE=True, allLUTs=allL,
                                         cartDROP=drop, store_in_memory=True, tagLOS='LOS07')
    assert a2.set == [] and a2.counter == len(temps) and os.path.basename(a2.filename) == \
        'abscoeff_LOS07_mol_6_iso_1.pic'
    back_a = np.array([a2.read_one().spectrum for _ in temps])
    back_e = np.array([e2.read_one().spectrum for _ in temps])
    a2.finalize_IO(), e2.finalize_IO()
    assert rel_err(back_a, ref["abscoeff_lte"][0], 1e-9) < 1e-9
    assert rel_err(back_e, ref["abscoeff_lte"][1]) < 1e-9
    assert smm.make_abscoeff_LUTS_fast(sg, iso1, temps, press, allLUTs={(iso1.mol_name, 1): None},
                                       track_levels=track) == (None, None, None, None)


def test_slow_line_by_line_abscoeff_matches_reference(ref2, case, tmp_path):
    """make_abscoeff_isomolec(useLUTs=False) (smm:1880-2131): the reference's per-step
    calc_shapes_lines -> add_PT -> populations chain against ONE batched K1 call; non-LTE
    isotopologue with a tracked level and an LTE isotopologue without levels."""
    smm, iso1, iso2 = case["smm"], case["iso1"], case["iso2"]
    T, P = list(ref2["slow_T"]), list(ref2["slow_P"])
    for s, lev in enumerate(iso1.levels):
        getattr(iso1, lev).local_vibtemp = list(ref2["slow_tvib"][s])
    drop = str(tmp_path) + '/'
    a, e, et, at = smm.make_abscoeff_isomolec([2998.0, 3006.0], iso1, T, P, LTE=False, useLUTs=False,
                                              lines=case["lines"], cartDROP=drop, track_levels=['lev_01'])
    assert rel_err(np.array([x.spectrum for x in a.set]), ref2["slow_abs"], 1e-9) < TOL_XS
    assert rel_err(np.array([x.spectrum for x in e.set]), ref2["slow_emi"]) < TOL_XS
    assert rel_err(np.array([x.spectrum for x in et['lev_01'].set]), ref2["slow_track_emi"]) < TOL_XS
    assert rel_err(np.array([x.spectrum for x in at['lev_01'].set]), ref2["slow_track_emi"]) < TOL_XS
    a2, e2 = smm.make_abscoeff_isomolec([2998.0, 3006.0], iso2, T[:2], P[:2], LTE=True, useLUTs=False,
                                        lines=case["lines"], cartDROP=drop)
    assert rel_err(np.array([x.spectrum for x in a2.set]), ref2["slow_abs_iso2"], 1e-9) < TOL_XS
    assert rel_err(np.array([x.spectrum for x in e2.set]), ref2["slow_emi_iso2"]) < TOL_XS
    with pytest.raises(ValueError):
        smm.make_abscoeff_isomolec([2998.0, 3006.0], iso1, T, P, useLUTs=False)
    # useLUTs=True takes the LUT path on the LUT's own grid
    sp = smm.prepare_spe_grid([2998.0, 3006.0]).spectral_grid
    lut = smm.LookUpTable(iso1, [2998.0, 3006.0], LTE=False)
    lut.make(sp, case["lines"], [[p, t] for p in (0.01, 0.1, 2.0) for t in (150., 160., 170., 180.)])
    a3, e3 = smm.make_abscoeff_isomolec([2998.0, 3006.0], iso1, T, P, LTE=False, useLUTs=True,
                                        allLUTs={(iso1.mol_name, iso1.iso): lut}, cartDROP=drop)
    assert len(a3.set) == 3 and len(a3[0].spectrum) == len(sp.grid)
    # a LUT is an interpolation of the same physics: close to the exact evaluation, not equal
    assert rel_err(a3[2].spectrum, ref2["slow_abs"][2], 1e-3) < 0.2


def test_prepare_calc_shapes_and_worker(ref, case):
    """PrepareCalcShapes (:1440-1462) gives EVERY line a shape and G coefficients (no level
    filter; E_vib = 0 for lines it cannot link), do_for_th_calc (:1418-1437) slices like the
    reference's worker."""
    spcl, smm, iso1 = case["spcl"], case["smm"], case["iso1"]
    sp = smm.prepare_spe_grid([2998.0, 3006.0]).spectral_grid
    T, P = ref["phys_PT"][0]
    l1 = [l for l in case["lines"] if l.Iso == 1]
    idx = [i for i, l in enumerate(case["lines"]) if l.Iso == 1]
    out = spcl.PrepareCalcShapes(sp, l1, T, P, iso1)
    assert len(out) == len(l1) and out[0] is l1[0]
    for lin, i in zip(out, idx):
        for k, ct in enumerate(CTYPES):
            assert lin.G_coeffs[ct] == pytest.approx(ref["phys_gcoeff"][0, i, k], rel=1e-12), (i, ct)
        assert len(lin.shape.spectrum) == 13010
        assert abs(lin.shape.spectral_grid.grid[6505] - sp.grid[np.argmin(np.abs(sp.grid - lin.Freq))]) < 1e-9
    P2, T2 = ref["cells_PT"][0]
    out = spcl.PrepareCalcShapes(sp, l1, T2, P2, iso1)
    by_freq = dict((l.Freq, l) for l in out)
    for j, i in enumerate(ref["shape_pick"]):
        lin = by_freq[float(ref["shape_freq"][int(i)])]
        assert rel_err(lin.shape.spectrum, ref["shape_spectra"][j]) < TOL_XS
    parts = []
    for i in range(3):
        box = Box()
        spcl.do_for_th_calc(sp, l1, T2, P2, iso1, i, box, n_threads=3)
        parts.append(box.item)
    n = len(l1) // 3
    assert [len(p) for p in parts] == [n, n, len(l1) - 2 * n]
    assert [l.Freq for p in parts for l in p] == [l.Freq for l in l1]
    co = spcl.SpectralGcoeff('absorption', sp, 6, 1, iso1.MM, iso1.lev_00.minimal_level_string())
    assert len(co.calc_shapes(l1, T2, P2, iso1)) == len(ref["shape_freq"])


def test_lutset_make_matches_reference_cells(ref, case):
    """LutSet.make (smm:1069-1119, the whole-set builder): the cells of one level equal the
    reference-built LUT's."""
    smm, iso1 = case["smm"], case["iso1"]
    sp = smm.prepare_spe_grid([2998.0, 3006.0]).spectral_grid
    # a LutSet knows one level only: the lines whose other level is unknown to the isotopologue
    # (dropped by LookUpTable.make through calc_shapes_lines, :1384-1388) are taken out here
    linked = [l for l in case["lines"] if l.Iso == 2 or l.LinkToMolec(iso1)]
    for s, lev in enumerate(iso1.levels):
        st = smm.LutSet(6, 1, iso1.MM, level=getattr(iso1, lev))
        sets = st.make(sp, linked, [list(pt) for pt in ref["cells_PT"]])
        assert len(sets) == 2 and st.find(*ref["cells_PT"][1]) == 1
        for k, ct in enumerate(CTYPES):
            want = ref["cells_nonlte"][s, k]
            if not np.any(want):
                assert not np.any(sets[0][ct].spectrum), (lev, ct)
            else:
                assert rel_err(sets[0][ct].spectrum, want) < TOL_XS, (lev, ct)
            assert rel_err(sets[1][ct].spectrum[::8], ref["cells_nonlte_b8"][s, k]) < TOL_XS
    st = smm.LutSet(6, 2, case["iso2"].MM, level=None)
    sets = st.make(sp, case["lines"], [list(ref["cells_lte_PT"])])
    for k, ct in enumerate(CTYPES):
        assert rel_err(sets[0][ct].spectrum, ref["cells_lte"][k]) < TOL_XS
