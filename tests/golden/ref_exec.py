"""Executes the reference's OWN Python (spect_classes.py, spect_main_module.py) under Python 3.

The reference is Python 2 + three f2py modules + a module that is not in its tree.
Both source files parse under the Python-3 grammar, so they are compiled from where they lie
(`/root/reference`, nothing is copied) after ONE syntactic rewrite - every `/` becomes a call that
keeps Python-2 semantics (floor division when both operands are integers) - and executed with
stand-ins for what does not exist here:

    lineshape, fparts_mod   the C restatement of the Fortran (oracle/sr_oracle.c), f2py-shaped
    spect_base_module       tests/golden/ref_sbm_stub.py (missing upstream)
    cPickle                 pickle
    matplotlib, memory_profiler   empty modules

What this pins: every Python-level formula and every piece of book-keeping of SURVEY 8a rows
A1, A3-A11 and 8f rows 1 and 3 (widths, G coefficients, window placement and clipping, level
selection, the zero-padded staging matrix, LUT files, LutSet.calculate, make_abscoeff_LUTS_fast,
calc_PT_couples_atmosphere, convolution, FOV integration).  The Fortran itself (A2, A6's inner
loop, A9's table, A13) is pinned separately by f77_exec.py, which executes the reference's Fortran
source text; `with fortran_from_source():` below swaps the C stand-ins for those executed routines,
and tests/test_f77_golden.py uses it to show that the reference's Python together with the
reference's Fortran reproduce the committed fixtures bit for bit.  What stays unpinned: A12 (source
missing upstream).

Only tests/golden/make_ref_golden.py and tests/ use this module; /root/reference does not exist on
the GPU box, so the fixtures it produces are committed.
"""
import ast
import os
import pickle
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get('SPECTROBOT_REFERENCE', '/root/reference')

_LOADED = None


def available():
    return os.path.exists(os.path.join(REFERENCE, 'spect_classes.py'))


def py2div(a, b):
    """Python-2 `/`: floor division for two integers (spect_classes.py:1050, 1423), true division
    otherwise."""
    ints = (int, np.integer)
    if isinstance(a, ints) and isinstance(b, ints) and not isinstance(a, bool):
        return a // b
    return a / b


class _Py2Division(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            call = ast.Call(func=ast.Name(id='__py2div__', ctx=ast.Load()),
                            args=[node.left, node.right], keywords=[])
            return ast.copy_location(call, node)
        return node


def _py2open(name, mode='r', *a, **kw):
    """Python-2 `open`: no text/binary distinction on Linux, so pickles written with mode 'w' /
    read with 'r' (spect_main_module.py:1632, 1724) work; here '.pic' files get the 'b' flag."""
    if str(name).endswith('.pic') and 'b' not in mode:
        mode += 'b'
    return open(name, mode, *a, **kw)


def _stub(name, **names):
    mod = types.ModuleType(name)
    mod.__dict__.update(names)
    sys.modules[name] = mod
    return mod


def _exec_reference(name):
    path = os.path.join(REFERENCE, name + '.py')
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')          # invalid escape sequences in docstrings
        tree = _Py2Division().visit(ast.parse(open(path).read(), filename=path))
    ast.fix_missing_locations(tree)
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod.__dict__['__py2div__'] = py2div
    mod.__dict__['open'] = _py2open
    sys.modules[name] = mod
    exec(compile(tree, path, 'exec'), mod.__dict__)
    return mod


def load(imxlines=256):
    """(spect_classes, spect_main_module, spect_base_module stub) of the reference.

    imxlines: the Fortran array bound of sum_all_lines (parameters.inc:64, mirrored at
    spect_classes.py:28).  add_lines_to_spectrum zero-pads its staging matrix to
    imxlines x 13010 doubles (4.16 GB at the original 40000); the bound does not enter the
    arithmetic, so it is lowered here to keep fixture generation light."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError('reference tree not found at ' + REFERENCE)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import cpu_oracle as O
    import importlib.util

    _stub('matplotlib')
    _stub('matplotlib.pyplot')
    _stub('matplotlib.cm')
    _stub('memory_profiler', profile=lambda f: f)
    sys.modules['cPickle'] = pickle

    def sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe):
        return O.sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe)

    _stub('lineshape', humliv_bb=O.humliv_bb, sum_all_lines=sum_all_lines)
    _stub('fparts_mod', bd_tips_2003=O.bd_tips_2003)
    spec = importlib.util.spec_from_file_location('spect_base_module',
                                                  os.path.join(HERE, 'ref_sbm_stub.py'))
    sbm = importlib.util.module_from_spec(spec)
    sys.modules['spect_base_module'] = sbm
    spec.loader.exec_module(sbm)

    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        spcl = _exec_reference('spect_classes')
        spcl.imxlines = int(imxlines)
        if not hasattr(np, 'trapz'):
            np.trapz = np.trapezoid
        smm = _exec_reference('spect_main_module')
    _LOADED = (spcl, smm, sbm)
    return _LOADED


class quiet(object):
    """Silences the reference's progress prints (and its `echo >> control_spectrobot`)."""

    def __enter__(self):
        self._out = sys.stdout
        sys.stdout = open(os.devnull, 'w')
        return self

    def __exit__(self, *a):
        sys.stdout.close()
        sys.stdout = self._out


class fortran_from_source(object):
    """Context manager: while active, the `lineshape` and `fparts_mod` stand-ins execute the
    reference's OWN Fortran source (lineshape.f, fparts_mod.f, read where they lie and run by the
    mechanical FORTRAN 77 executor f77_exec.py) instead of the C restatement - so that the
    reference's Python and the reference's Fortran run together with nothing of this repository
    in the arithmetic.  f2py shapes: every output is a new array."""

    def __enter__(self):
        if HERE not in sys.path:
            sys.path.insert(0, HERE)
        import f77_exec as F
        ls = F.load(os.path.join(REFERENCE, 'lineshape.f'))
        fp = F.load(os.path.join(REFERENCE, 'fparts_mod.f'))

        def humliv_bb(x, i1, i2, x0, lw, dw):
            x = np.ascontiguousarray(x, dtype=float)
            y = np.zeros(len(x))
            ls['humliv_bb'](x, int(i1), int(i2), float(x0), float(lw), float(dw), y)
            return y

        def sum_all_lines(spe_ini, matrix, init, fin, n_lines, n_spe):
            spe_ini = np.ascontiguousarray(spe_ini, dtype=float)
            out = np.empty_like(spe_ini)
            ls['sum_all_lines'](spe_ini, np.asarray(matrix), np.asarray(init).astype(np.int32),
                                np.asarray(fin).astype(np.int32), int(n_lines), int(n_spe), out)
            return out

        def bd_tips_2003(mol, iso):
            t, q = np.zeros(119), np.zeros(119)
            gi = fp['bd_tips_2003'](int(mol), int(iso), 0.0, t, q)['gi']
            return gi, t, q

        self._saved = []
        for mod, name, fn in (('lineshape', 'humliv_bb', humliv_bb),
                              ('lineshape', 'sum_all_lines', sum_all_lines),
                              ('fparts_mod', 'bd_tips_2003', bd_tips_2003)):
            self._saved.append((mod, name, getattr(sys.modules[mod], name)))
            setattr(sys.modules[mod], name, fn)
        return self

    def __exit__(self, *a):
        for mod, name, fn in self._saved:
            setattr(sys.modules[mod], name, fn)
