"""Drop-in for the reference's f2py module `curgods` (curgods.f): Curtis-Godson column integrals
over piecewise-exponential number density (linear vmr / f).  n_p = number of valid points."""
import numpy as np

from ._lib import as_f64, check, dptr, lib


def _run(k, nd, vmr, f, x, n_p):
    """One integral through the entry point of its own name, sr_curgod_<k> (include/spectrobot.h)."""
    n_p = int(n_p)
    arrs = [as_f64(np.asarray(a, dtype=float)[:n_p]) for a in (nd, vmr, f, x) if a is not None]
    res = np.empty(1)
    fn = getattr(lib(), "sr_curgod_%d" % k)
    check(fn(*([dptr(a) for a in arrs] + [n_p, dptr(res)])))
    return float(res[0])


def curgod_fort_1(nd, x, n_p):
    """curgods.f:2-21"""
    return _run(1, nd, None, None, x, n_p)


def curgod_fort_2(nd, vmr, x, n_p):
    """curgods.f:24-45"""
    return _run(2, nd, vmr, None, x, n_p)


def curgod_fort_3(nd, vmr, f, x, n_p):
    """curgods.f:48-73"""
    return _run(3, nd, vmr, f, x, n_p)


def curgod_fort_4(nd, vmr, f, x, n_p):
    """curgods.f:76-97"""
    return _run(4, nd, vmr, f, x, n_p)


def curgod_batch(k, nd, vmr, f, x):
    """Batched form: arrays [n_batch, n_p] -> res[n_batch] (one launch)."""
    nd = as_f64(nd)
    n_batch, n_p = nd.shape
    arrs = [nd] + [None if a is None else as_f64(a) for a in (vmr, f)] + [as_f64(x)]
    res = np.empty(n_batch)
    p = [None if a is None else dptr(a) for a in arrs]
    check(lib().sr_curgod(k, p[0], p[1], p[2], p[3], n_p, n_batch, dptr(res)))
    return res
