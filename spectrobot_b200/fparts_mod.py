"""Drop-in for the reference's f2py module `fparts_mod` (fparts_mod.f)."""
import ctypes as C

import numpy as np

from ._lib import check, dptr, lib


def bd_tips_2003(mol, iso):
    """gi, t_grid, QT_grid = fparts_mod.bd_tips_2003(mol, iso)  [fparts_mod.f:33-53]."""
    gi = C.c_double()
    t = np.empty(119)
    q = np.empty(119)
    check(lib().sr_bd_tips_2003(int(mol), int(iso), C.byref(gi), dptr(t), dptr(q)))
    return gi.value, t, q
