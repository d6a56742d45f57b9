"""CPU check of the arithmetic the kernels are built from.  The product's device-function header
(spectrobot_b200/csrc/sr_device.cuh: the Humlicek regions of humliv_bb in their literal and their
hot-loop forms, the Curtis-Godson segments, exp / expm1 / layer update of the LOS recursion) is
compiled FOR THE HOST by g++ through tests/device_math/shim/cuda_runtime.h - a test-only stand-in
that maps the four CUDA intrinsics the header uses and models MUFU.RCP64H as a 20-bit reciprocal -
and compared with the reference's Fortran executed from source (tests/golden/f77_golden.npz) and
with libm.  This covers the formulas, not the kernels around them (indexing, tiling, FMA
contraction by nvcc): those are the -m gpu tests.  Nothing here is shipped or loaded by the
product, which has no CPU path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


def P(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def dm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("device_math") / "libdevice_math_host.so")
    src = os.path.join(ROOT, "tests", "device_math")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-shared", "-fPIC",
                    "-I", os.path.join(src, "shim"), "-I", os.path.join(ROOT, "spectrobot_b200", "csrc"),
                    "-o", out, os.path.join(src, "device_math_host.cpp")], check=True)
    lib = C.CDLL(out)
    lib.dm_curgod.restype = C.c_double
    return lib


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "f77_golden.npz"))


def test_humlicek_regions_against_the_executed_fortran(dm, gold, oracle):
    """15 (P,T) cases, line inside its window: region boundaries by the Fortran's NINT index rules;
    literal forms (Tier-1 drop-in): the float32-CMPLX core is BIT-IDENTICAL to lineshape.f, the
    rationals agree to the difference between a running abscissa (lineshape.f:467) and a closed
    form; hot-loop forms (tile kernel: direct complex division, region 1 through the 20-bit
    reciprocal + Newton step): 1e-13 in the core, 1e-10 overall - four orders below the 1e-6 gate."""
    lin, keep = gold["lin"], gold["keep"]
    n = 0
    for k in np.flatnonzero(gold["hv_gpu"] == 1):
        c, i1, i2, x0, lw, dw = gold["hv_in"][k]
        if not str(gold["hv_tags"][k]).startswith("inside"):
            continue
        x = np.ascontiguousarray(lin + c)
        ref = gold["hv_y"][k]
        for mode in (0, 1):
            y = np.zeros(13010)
            reg = np.zeros(4, dtype=np.int32)
            rc = dm.dm_humliv_inside(P(x), 13010, C.c_double(x0), C.c_double(lw), C.c_double(dw), mode,
                                     P(y), reg.ctypes.data_as(C.POINTER(C.c_int)))
            assert rc == 0
            assert np.array_equal(reg, oracle.humliv_regions(x, 1, 13010, x0, lw, dw))
            il, ir, il2, ir2 = reg
            core = (keep + 1 > il2) & (keep + 1 < ir2)
            assert core.sum() > 20
            err = np.abs(y[keep] - ref) / ref
            if mode == 0:
                assert np.array_equal(y[keep][core], ref[core])
            assert err[core].max() < 1e-13 and err.max() < 1e-10, (gold["hv_tags"][k], mode)
        n += 1
    assert n == 15


def test_region1_hot_loop_forms(dm):
    """(u + 1) / (u^2 + c2) through MUFU.RCP64H + one Newton step, three instruction orders."""
    rng = np.random.default_rng(5)
    for ry in (1e-4, 0.05, 1.0, 20.0):
        x = np.r_[rng.uniform(15.0 + ry, 4000.0, 5000), 15.0 + ry]
        u = np.ascontiguousarray(x * x + ry * ry - 0.5)
        c2 = 2.0 * ry * ry
        exact = (u + 1.0) / (u * u + c2)
        got = [np.zeros_like(u) for _ in range(3)]
        dm.dm_reg1_forms(P(u), len(u), C.c_double(c2), P(got[0]), P(got[1]), P(got[2]))
        for g in got:
            assert np.max(np.abs(g - exact) / exact) < 2e-12


def test_curtis_godson_segments_against_the_executed_fortran(dm, gold):
    """curgod_seg1..4 summed over the segments = curgod_fort_1..4 (curgods.f:2-98)."""
    for k, n_p in enumerate(gold["cg_n_p"]):
        nd, vmr, f, x = [np.ascontiguousarray(a[:n_p]) for a in gold["cg_in"][k]]
        for j in (1, 2, 3, 4):
            got = dm.dm_curgod(j, P(nd), P(vmr), P(f), P(x), int(n_p))
            ref = gold["cg_res"][k][j - 1]
            assert abs(got - ref) <= 1e-14 * abs(ref), (k, j)


def test_exp_and_layer_update_forms(dm):
    """exp_pair / exp_phi (one range reduction, one polynomial) against libm, and every short form
    of the layer update I <- I e^-t + J (1 - e^-t)/t within the bound its comment states."""
    rng = np.random.default_rng(3)
    xs = np.r_[rng.uniform(-700, 700, 20000), rng.uniform(-1, 1, 20000), 10.0 ** rng.uniform(-300, -1, 2000),
               -10.0 ** rng.uniform(-300, -1, 2000), 0.0]
    ex, em = np.zeros_like(xs), np.zeros_like(xs)
    dm.dm_exp_pair(P(xs), len(xs), P(ex), P(em))
    rem = np.expm1(xs)
    assert np.max(np.abs(ex - np.exp(xs)) / np.exp(xs)) < 1e-15
    assert np.max(np.abs(em - rem) / np.where(rem == 0, 1.0, np.abs(rem))) < 1e-15 and em[-1] == 0.0
    ts = np.r_[rng.uniform(-5, 700, 20000), rng.uniform(-1, 1, 20000), 10.0 ** rng.uniform(-300, -1, 2000),
               -10.0 ** rng.uniform(-300, -1, 2000), 0.0]
    ex, ph = np.zeros_like(ts), np.zeros_like(ts)
    dm.dm_exp_phi(P(ts), len(ts), P(ex), P(ph))

    def phi(t):
        return np.where(t == 0, 1.0, -np.expm1(-t) / np.where(t == 0, 1.0, t))

    assert np.max(np.abs(ex - np.exp(-ts)) / np.exp(-ts)) < 1e-15
    assert np.max(np.abs(ph - phi(ts)) / phi(ts)) < 2e-15 and ph[-1] == 1.0
    # forms: 0 full, 1 |t| < 1e-2 (degree 4), 2 |t| < ln2/2 (degree 9), 3 float32-valued inputs
    for form, tmax, tol in ((0, 700.0, 2e-15), (1, 1e-2, 2e-13), (2, 0.3465, 3e-11), (3, 700.0, 5e-12)):
        n = 40000
        lo = -tmax if form in (1, 2) else -3.0        # tau < 0: population inversion
        t = np.r_[rng.uniform(lo, tmax, n // 2), 10.0 ** rng.uniform(-12, np.log10(tmax), n // 2)]
        I, J = rng.uniform(0, 2, n), rng.uniform(0, 2, n) * np.abs(t)
        if form == 3:
            t, J = t.astype(np.float32).astype(float), J.astype(np.float32).astype(float)
        out = np.zeros(n)
        dm.dm_layer_update(P(I), P(t), P(J), n, form, 0, P(out))
        ref = I * np.exp(-t) + J * phi(t)
        assert np.max(np.abs(out - ref) / np.abs(ref)) < tol, form
        dm.dm_layer_update(P(I), P(t), P(J), n, form, 1, P(out))          # solo_absorption
        assert np.max(np.abs(out - I * np.exp(-t)) / (I * np.exp(-t))) < max(tol, 1e-12), form
