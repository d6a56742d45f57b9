"""The C-ABI library loads on a CPU-only box and exports every symbol include/spectrobot.h declares
(no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "spectrobot.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from spectrobot_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libspectrobot.so does not export %s" % n
    # and the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_host_only_entry_points_work_without_gpu():
    from spectrobot_b200 import _lib, fparts_mod
    L = _lib.lib()
    assert L.sr_version() >= 100
    gi, t, q = fparts_mod.bd_tips_2003(6, 1)
    assert gi == 1.0 and len(t) == 119 and q[0] == float(np.float32(54.791))
    with pytest.raises(_lib.SpectrobotError) as e:
        fparts_mod.bd_tips_2003(99, 1)
    assert e.value.code == _lib.SR_ERR_TABLE
    c = _lib.sr_consts()
    L.sr_default_consts(ctypes.byref(c))
    p = _lib.python_consts()
    for f, _ in _lib.sr_consts._fields_:
        assert getattr(c, f) == getattr(p, f), f        # scipy CODATA == the exact SI values


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the compute entry points fail loudly (SR_ERR_CUDA), they do not
    fall back to a CPU implementation."""
    from spectrobot_b200 import _lib, lineshape
    if _lib.cuda_available():
        pytest.skip("CUDA device present")
    with pytest.raises(_lib.SpectrobotError) as e:
        lineshape.humliv_bb(np.linspace(0, 1, 13010), 1, 13010, 0.5, 1e-3, 1e-3)
    assert e.value.code == _lib.SR_ERR_CUDA


def test_curgod_entry_points_under_their_f2py_names():
    """sr_curgod_1..4 (one symbol per curgods.f routine) reject bad arguments with
    SR_ERR_ARG before any device work, and fail loudly without a device otherwise."""
    from spectrobot_b200 import _lib, curgods
    L = _lib.lib()
    a = np.linspace(1.0, 2.0, 8)
    res = np.zeros(1)
    assert L.sr_curgod_1(_lib.dptr(a), _lib.dptr(a), 0, _lib.dptr(res)) == _lib.SR_ERR_ARG
    assert L.sr_curgod_2(_lib.dptr(a), None, _lib.dptr(a), 8, _lib.dptr(res)) == _lib.SR_ERR_ARG
    assert L.sr_curgod_3(_lib.dptr(a), _lib.dptr(a), None, _lib.dptr(a), 8, _lib.dptr(res)) == _lib.SR_ERR_ARG
    assert L.sr_curgod_4(None, _lib.dptr(a), _lib.dptr(a), _lib.dptr(a), 8, _lib.dptr(res)) == _lib.SR_ERR_ARG
    if not _lib.cuda_available():
        with pytest.raises(_lib.SpectrobotError) as e:
            curgods.curgod_fort_2(a, a, a, 8)
        assert e.value.code == _lib.SR_ERR_CUDA


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "spectrobot_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "cpu_oracle" not in txt and "sr_oracle" not in txt and "orc_" not in txt, f
