"""Drop-in for the reference's f2py module `lineshape` (lineshape.f), backed by libspectrobot.so.

Same call signatures and array conventions as the f2py wrappers: inputs are copied,
every output is a new array, indices are 1-based inclusive.  Where the Fortran executes `stop`
(which kills the reference's worker process) these raise SpectrobotError instead.
"""
import numpy as np

from ._lib import IMXSIG, IMXLINES, IMXSIG_LONG, as_f64, as_i32, check, dptr, iptr, lib


def humliv_bb(x, i1, i2, x0, lw, dw):
    """y = lineshape.humliv_bb(x,i1,i2,x0,lw,dw)  [lineshape.f:226-240; spect_classes.py:1999].
    x must have exactly imxsig = 13010 elements, like the f2py signature."""
    x = as_f64(x)
    if x.shape != (IMXSIG,):
        raise ValueError("0-th dimension must be fixed to %d but got %s" % (IMXSIG, x.shape))
    y = np.zeros(IMXSIG)
    check(lib().sr_humliv_bb(dptr(x), IMXSIG, int(i1), int(i2), float(x0), float(lw), float(dw),
                             dptr(y)))
    return y


def sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe):
    """spe_fin = lineshape.sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe)
    [lineshape.f:2-13; spect_classes.py:1092].  matrix: (ld, n_win) Fortran-contiguous (the
    reference passes (40000, 13010)); init/fin: 1-based inclusive; n_spe is unused, as in the
    Fortran."""
    spe_ini = as_f64(spe_ini)
    m = np.asfortranarray(matrix, dtype=np.float64)
    init = as_i32(init)
    fin = as_i32(fin)
    out = np.empty_like(spe_ini)
    check(lib().sr_sum_all_lines(dptr(spe_ini), dptr(m), iptr(init), iptr(fin), int(n_lines),
                                 m.shape[0], m.shape[1], len(spe_ini), dptr(out)))
    return out
