"""GPU parity of the Tier-1 drop-ins (sr_humliv_bb, sr_sum_all_lines, sr_curgod_1..4, through the
C ABI) against values produced by EXECUTING the reference's own Fortran source
(tests/golden/f77_golden.npz, made by tests/golden/make_f77_golden.py with the mechanical
FORTRAN 77 executor tests/golden/f77_exec.py).  The inputs are those of tests/test_gpu_voigt.py
(::test_humliv_bb_inside_branch, ::test_humliv_bb_other_branches, ::test_curgod); tolerance 1e-6
relative on cross sections (north_star).  Needs neither /root/reference nor oracle/."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_XS = 1e-6
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "f77_golden.npz")


@pytest.fixture(scope="module")
def gold():
    from spectrobot_b200 import _lib
    assert _lib.cuda_available(), "libspectrobot.so sees no CUDA device"
    return np.load(GOLD)


def test_humliv_bb_matches_the_executed_fortran(gold):
    """lineshape.f:226-569: 15 (P,T) cases with the line inside its window, x0 left / right of the
    window, two [i1, i2] sub-ranges."""
    from spectrobot_b200 import lineshape
    lin, keep = gold["lin"], gold["keep"]
    on_gpu = np.flatnonzero(gold["hv_gpu"] == 1)
    assert len(on_gpu) == 19
    for k in on_gpu:
        c, i1, i2, x0, lw, dw = [float(v) for v in gold["hv_in"][k]]
        got = lineshape.humliv_bb(lin + c, int(i1), int(i2), x0, lw, dw)
        assert rel_err(got[keep], gold["hv_y"][k]) < TOL_XS, str(gold["hv_tags"][k])
        assert np.all(got[:int(i1) - 1] == 0) and np.all(got[int(i2):] == 0)
    tags = [str(t) for t in gold["hv_tags"]]
    for t, tag in enumerate(gold["hv_whole_tags"]):
        k = tags.index(str(tag))
        if gold["hv_gpu"][k] != 1:
            continue
        c, i1, i2, x0, lw, dw = [float(v) for v in gold["hv_in"][k]]
        got = lineshape.humliv_bb(lin + c, int(i1), int(i2), x0, lw, dw)
        assert rel_err(got, gold["hv_whole_y"][t]) < TOL_XS, str(tag)


def test_sum_all_lines_matches_the_executed_fortran(gold):
    """lineshape.f:2-25.  The terms span twelve decades and cancel, so the bound is the one any
    order of summation satisfies: |got - ref| <= 1e-13 x (|spe_ini| + sum of |terms|) per point."""
    from spectrobot_b200 import lineshape
    m = np.asfortranarray(gold["sal_matrix"].astype(np.float64))
    spe = gold["sal_spe"].astype(np.float64)
    init, fin = gold["sal_init"].astype(np.int32), gold["sal_fin"].astype(np.int32)
    n_lines = 40
    got = lineshape.sum_all_lines(spe, m, init, fin, n_lines, len(spe))
    scale = np.abs(spe).copy()
    for l in range(n_lines):
        n = int(fin[l]) - int(init[l]) + 1
        scale[int(init[l]) - 1:int(fin[l])] += np.abs(m[l, :n])
    assert got.shape == gold["sal_res"].shape
    assert np.all(np.abs(got - gold["sal_res"]) <= 1e-13 * scale)


def test_curgod_matches_the_executed_fortran(gold):
    """curgods.f:2-98 through sr_curgod_1..4 (the f2py names)."""
    from spectrobot_b200 import curgods
    k = int(gold["cg_gpu_case"])
    n_p = int(gold["cg_n_p"][k])
    nd, vmr, f, x = [np.ascontiguousarray(a[:n_p], dtype=np.float64) for a in gold["cg_in"][k]]
    got = [curgods.curgod_fort_1(nd, x, n_p), curgods.curgod_fort_2(nd, vmr, x, n_p),
           curgods.curgod_fort_3(nd, vmr, f, x, n_p), curgods.curgod_fort_4(nd, vmr, f, x, n_p)]
    ref = [float(v) for v in gold["cg_res"][k]]
    for j in range(4):
        assert abs(got[j] - ref[j]) <= 1e-9 * abs(ref[j]), (j, got[j], ref[j])
