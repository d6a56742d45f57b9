"""CPU tests: the oracle's restatement of the reference's FORTRAN (oracle/sr_oracle.c: humliv_bb,
humli_bb, sum_all_lines, curgod_fort_1..4, the TIPS tables) against fixtures produced by EXECUTING the reference's
own Fortran source text with the mechanical FORTRAN 77 executor tests/golden/f77_exec.py
(tests/golden/make_f77_golden.py -> f77_golden.npz).  Bit for bit: both sides perform the same
IEEE operations in the same order.  Also: the executor's language rules on small programs written
here, and - when /root/reference is present - that the committed fixture is what the source
computes now."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("SPECTROBOT_REFERENCE", "/root/reference")
IMXSIG = 13010


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "f77_golden.npz"))


@pytest.fixture(scope="module")
def f77():
    sys.path.insert(0, GOLD)
    import f77_exec
    return f77_exec


def test_humliv_bb_restatement_is_bit_identical_to_the_executed_fortran(oracle, gold):
    """lineshape.f:226-569, all three branches (x0 left of / right of / inside the window),
    sub-ranges of [i1, i2], widths from 1e-9 to 3 cm-1: 61 cases."""
    lin, keep = gold["lin"], gold["keep"]
    assert len(gold["hv_tags"]) >= 60
    for k, tag in enumerate(gold["hv_tags"]):
        c, i1, i2, x0, lw, dw = gold["hv_in"][k]
        y = oracle.humliv_bb(lin + c, int(i1), int(i2), x0, lw, dw)
        assert np.array_equal(y[keep], gold["hv_y"][k]), tag
        touched = (keep + 1 >= i1) & (keep + 1 <= i2)
        assert np.all(gold["hv_y"][k][~touched] == 0.0), tag          # untouched outside [i1, i2]
        assert np.all(gold["hv_y"][k][touched] > 0.0), tag
    for t, tag in enumerate(gold["hv_whole_tags"]):
        k = list(gold["hv_tags"]).index(tag)
        c, i1, i2, x0, lw, dw = gold["hv_in"][k]
        assert np.array_equal(oracle.humliv_bb(lin + c, int(i1), int(i2), x0, lw, dw), gold["hv_whole_y"][t])
    # the two STOP statements (lineshape.f:253-264)
    msgs = list(gold["hv_stop_msgs"])
    assert "i1 > i2" in msgs[0] and "dw <=0" in msgs[1] and "dw <=0" in msgs[2]
    x = np.linspace(0, 1, IMXSIG)
    for args in ((10, 5, 0.5, 1e-3, 1e-3), (1, IMXSIG, 0.5, 1e-3, 0.0), (1, IMXSIG, 0.5, 1e-3, -1.0)):
        with pytest.raises(RuntimeError):
            oracle.humliv_bb(x, *args)


def test_humli_bb_restatement(oracle, gold):
    """lineshape.f:150-205 (scalar form, exported by the f2py module)."""
    got = np.array([oracle.humli_bb(a, b) for a, b in zip(gold["hb_rx"], gold["hb_ry"])])
    assert np.array_equal(got, gold["hb_rre"])


def test_sum_all_lines_restatement(oracle, gold):
    """lineshape.f:2-25: overlapping windows added in line order (rounding depends on it), windows
    touching the first and the last point of the spectrum."""
    m = np.asfortranarray(gold["sal_matrix"].astype(float))
    spe = gold["sal_spe"].astype(float)
    init, fin = gold["sal_init"], gold["sal_fin"]
    assert init.min() == 1 and fin.max() == len(spe)
    got = oracle.sum_all_lines(spe, m, init, fin, 40)
    assert np.array_equal(got, gold["sal_res"])


def test_curgod_restatement(oracle, gold):
    """curgods.f:2-98, n_p = 2, 3, 17, 120."""
    for k, n_p in enumerate(gold["cg_n_p"]):
        nd, vmr, f, x = [a[:n_p].copy() for a in gold["cg_in"][k]]
        got = [oracle.curgod(1, nd, x), oracle.curgod(2, nd, vmr, x), oracle.curgod(3, nd, vmr, f, x),
               oracle.curgod(4, nd, vmr, f, x)]
        assert got == list(gold["cg_res"][k]), (k, got, gold["cg_res"][k])


def test_tips_tables_of_oracle_and_product_are_the_executed_fortran(oracle, gold):
    """fparts_mod.f bd_tips_2003 -> QT_*: gi, the 119-node temperature grid and Q(T) of every
    (mol, iso) the dispatcher reaches, for the oracle's table and for the PRODUCT's
    sr_bd_tips_2003 (a host-only entry point of libspectrobot.so: runs without a GPU)."""
    from spectrobot_b200 import fparts_mod
    keys = gold["tips_keys"]
    assert len(keys) >= 100 and [6, 1] in keys.tolist() and [23, 1] in keys.tolist() and [26, 1] in keys.tolist()
    for k, (mol, iso) in enumerate(keys):
        q = gold["tips_q"][k].astype(float)
        for impl in (oracle.bd_tips_2003, fparts_mod.bd_tips_2003):
            gi, t, qq = impl(int(mol), int(iso))
            assert gi == gold["tips_gi"][k], (mol, iso)
            assert np.array_equal(np.asarray(t), gold["tips_t"]) and np.array_equal(np.asarray(qq), q), (mol, iso)


# ---------------------------------------------------------------------------------------------
# the executor itself: language rules on programs written here
# ---------------------------------------------------------------------------------------------
_PROGRAM = """\
C     a comment line
      subroutine rules(n, a, v, ires, res)
      implicit none
      integer*4 n, ires(10), i, k
      real*8 a, v(5), res(10), s
      real*4 q
      complex*16 z, w
* integer division truncates towards zero; mixed mode promotes
      ires(1) = 7/2
      ires(2) = (-7)/2
      ires(3) = nint(2.5d0) + 10*nint(-2.5d0)
      ires(4) = nint(0.49999999999999994d0)
      res(1) = 7/2*a
      res(2) = 0.1                      ! REAL*4 literal widened
      res(3) = 0.1d0
      res(4) = a**3
      res(5) = -a**2
      q = a
      res(6) = q
* DO: trip count fixed at entry, variable one step past the end afterwards
      k = 0
      do i = 1, n
        k = k + i
        n = 1
      end do
      ires(5) = k
      ires(6) = i
      do i = 5, 1
        k = -1
      enddo
      ires(7) = i
      do i = 10, 1, -3
        ires(8) = i
      end do
      ires(9) = i
* default-kind CMPLX rounds both parts to REAL*4; complex*16 arithmetic on the result
      z = cmplx(a, -a)
      res(7) = dble(z)
      res(8) = dimag(z)
      w = (1.5 + z*(2.d0 + z))/
     &    (0.25 + z)
      res(9) = dble(w)
      res(10) = dimag(w)
      s = 0.d0
      i = 1
      do while ((s .lt. 2.5d0) .and. (i .le. 5))
        s = s + v(i)
        i = i + 1
      end do
      if (s .ge. 2.5d0) then
        v(1) = s
      else if (s .gt. 1.d0) then
        v(1) = -s
      else
        v(1) = 0.d0
      endif
      if (i .eq. 4) v(2) = 99.d0
      if (a .lt. 0.d0) stop 'negative a'
      end
"""


def test_executor_language_rules(f77, tmp_path):
    fn = tmp_path / "rules.f"
    fn.write_text(_PROGRAM)
    rules = f77.load(str(fn))["rules"]
    a = 1.0 / 3.0
    ires = np.zeros(10, dtype=np.int64)
    res = np.zeros(10)
    v = np.array([1.0, 1.0, 1.0, 1.0, 1.0])
    out = rules(4, a, v, ires, res)
    assert out["n"] == 1                                     # assigned inside the loop
    assert list(ires[:9]) == [3, -3, 3 - 30, 0, 10, 5, 5, 1, -2]
    f32 = lambda t: float(np.float32(t))   # noqa: E731
    assert res[0] == 3 * a and res[1] == f32(0.1) and res[2] == 0.1
    assert res[3] == (a * a) * a and res[4] == -(a * a) and res[5] == f32(a)
    assert res[6] == f32(a) and res[7] == f32(-a)
    z = complex(f32(a), f32(-a))
    num = complex(f32(1.5), 0.0) + z * (complex(2.0, 0.0) + z)
    den = complex(0.25, 0.0) + z
    w = f77._cdiv(complex(num.real, num.imag), den)
    assert abs(res[8] - w.real) <= 2e-16 and abs(res[9] - w.imag) <= 2e-16
    assert abs(complex(res[8], res[9]) - num / den) < 1e-15
    assert v[0] == 3.0 and v[1] == 99.0
    with pytest.raises(f77.F77Stop, match="negative a"):
        rules(1, -1.0, v, ires, res)
    # correctly rounded REAL*4 literals, Smith division against exact rational arithmetic
    assert f77.f4_literal("36183.31") == f32(36183.31) and f77.f4_literal(".56419") == f32(0.56419)
    assert f77.f4_literal("16777217") == 16777216.0           # tie -> even
    from fractions import Fraction as Fr
    a_, b_ = complex(3.0, -7.0), complex(0.5, 4.0)
    q = f77._cdiv(a_, b_)
    d = Fr(b_.real) ** 2 + Fr(b_.imag) ** 2
    er = (Fr(a_.real) * Fr(b_.real) + Fr(a_.imag) * Fr(b_.imag)) / d
    ei = (Fr(a_.imag) * Fr(b_.real) - Fr(a_.real) * Fr(b_.imag)) / d
    assert abs(Fr(q.real) - er) < Fr(1, 2 ** 50) and abs(Fr(q.imag) - ei) < Fr(1, 2 ** 50)


def test_executor_refuses_what_it_does_not_know(f77, tmp_path):
    """Labels other than on RETURN, jumps, undeclared names under IMPLICIT NONE, assignment to a
    DATA variable: the unit is left out, nothing is guessed; a CALL of such a unit fails when it
    is reached."""
    fn = tmp_path / "no.f"
    fn.write_text("      subroutine a(x)\n      real*8 x\n      goto 10\n 10   x = 1.d0\n      end\n"
                  "      subroutine b(x)\n      implicit none\n      real*8 x\n      x = y\n      end\n"
                  "      subroutine c(x)\n      real*8 x\n      call a(x)\n      end\n"
                  "      subroutine d(x)\n      real*8 x\n      x = 2.d0*x\n      end\n"
                  "      subroutine e(x)\n      real*8 x, t(2)\n      data t/1.,2./\n      t(1) = x\n      end\n")
    got = f77.load(str(fn))
    assert sorted(got) == ["c", "d"] and got["d"](1.5)["x"] == 3.0
    with pytest.raises(f77.F77Unsupported):
        got["c"](1.0)


_PROGRAM2 = """\
      subroutine outer(k, g, tab)
      integer*4 k
      real*8 g, tab(4), loc(3)
      data loc/ 0.1, 2*0.25E+01/
      tab(4) = loc(1) + loc(3)
      if (k .eq. 1) then
      call inner(k, g, tab)
      go to 100
      endif
      g = -1.d0
 100  return
      end
      subroutine inner(iso, gsi, q)
      implicit double precision (a-h,o-z)
      dimension xg(2), qq(2,3), q(4)
      data xg/ 1.,6./
      data (qq( 1,j),j=1,3)/ 0.54791E+02, 0.1, 3./
      data (qq( 2,j),j=1,3)/ 1., 2., 3./
      eps = 0.01
      gsi = xg(iso+1)
\tdo i=1,3
\t  q(i) = qq(iso,i)
\tenddo
   99 return
      end
"""


def test_executor_data_call_implicit_and_return_labels(f77, tmp_path):
    """The constructs of fparts_mod.f: IMPLICIT typing, DIMENSION, DATA (lists, repeat counts,
    implied DO over a row), CALL with scalars copied back, `go to` a labelled RETURN, tab-indented
    statements."""
    fn = tmp_path / "tips.f"
    fn.write_text(_PROGRAM2)
    units = f77.load(str(fn))
    tab = np.zeros(4)
    out = units["outer"](1, 0.0, tab)
    f32 = lambda t: float(np.float32(t))   # noqa: E731
    assert out == {"k": 1, "g": 6.0}
    assert list(tab) == [f32(54.791), f32(0.1), 3.0, f32(0.1) + 2.5]
    assert units["outer"](2, 0.0, tab)["g"] == -1.0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not on this box")
def test_f77_fixture_is_what_the_reference_source_computes_now(f77, gold):
    """Re-executes the reference's Fortran live (this container only): the committed fixture is its
    output, not a hand edit; and the executor took the routines from lineshape.f / curgods.f."""
    ls = f77.load(os.path.join(REF, "lineshape.f"))
    cg = f77.load(os.path.join(REF, "curgods.f"))
    assert {"humliv_bb", "humli_bb", "sum_all_lines"} <= set(ls)
    assert {"curgod_fort_1", "curgod_fort_2", "curgod_fort_3", "curgod_fort_4"} <= set(cg)
    assert ls["humliv_bb"].source.count("_cdiv(") == 6          # 3 branches x (region 3, region 4)
    lin, keep = gold["lin"], gold["keep"]
    for k in (0, 7, 14, 15, 16, 17, 18, 22, 30, 45, 60):
        c, i1, i2, x0, lw, dw = gold["hv_in"][k]
        y = np.zeros(IMXSIG)
        ls["humliv_bb"](lin + c, int(i1), int(i2), x0, lw, dw, y)
        assert np.array_equal(y[keep], gold["hv_y"][k]), gold["hv_tags"][k]
    m = np.asfortranarray(gold["sal_matrix"].astype(float))
    res = np.empty(len(gold["sal_spe"]))
    ls["sum_all_lines"](gold["sal_spe"].astype(float), m, gold["sal_init"], gold["sal_fin"], 40,
                        len(res), res)
    assert np.array_equal(res, gold["sal_res"])
    k = 11
    n_p = int(gold["cg_n_p"][k])
    nd, vmr, f, x = [a[:n_p].copy() for a in gold["cg_in"][k]]
    assert cg["curgod_fort_3"](nd, vmr, f, x, n_p, 0.0)["res"] == gold["cg_res"][k][2]
    assert ls["humli_bb"](gold["hb_rx"][3], gold["hb_ry"][3], 0.0)["rre"] == gold["hb_rre"][3]
    fp = f77.load(os.path.join(REF, "fparts_mod.f"))
    assert len(fp) >= 40 and "qt_ch4" in fp and "qt_h3p" not in fp
    for mol, iso in ((6, 1), (6, 3), (23, 2), (26, 1), (5, 4)):
        k = gold["tips_keys"].tolist().index([mol, iso])
        t, q = np.zeros(119), np.zeros(119)
        assert fp["bd_tips_2003"](mol, iso, 0.0, t, q)["gi"] == gold["tips_gi"][k]
        assert np.array_equal(q, gold["tips_q"][k].astype(float)) and np.array_equal(t, gold["tips_t"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not on this box")
def test_reference_python_with_reference_fortran_builds_the_committed_lut(f77, monkeypatch):
    """The reference's Python (LookUpTable.make -> LutSet.add_PT -> BuildCoeff ->
    add_lines_to_spectrum -> prepare_fortran_sum, CalcPartitionSum) run TOGETHER WITH the
    reference's Fortran (humliv_bb, sum_all_lines, bd_tips_2003 executed from lineshape.f /
    fparts_mod.f by f77_exec) - nothing of this repository in the arithmetic - reproduces, bit for
    bit, the LUT cell, shapes and partition sums of the first fixture set (ref_golden.npz, which
    was generated with the C restatement standing in for the f2py modules)."""
    import shutil
    import tempfile
    sys.path.insert(0, GOLD)
    import ref_exec as R
    import make_ref_golden as M
    ref = np.load(os.path.join(GOLD, "ref_golden.npz"))
    spcl, smm, sbm = R.load()
    lines = spcl.read_line_database(os.path.join(GOLD, "ref_lines.par"))
    dec = (lambda v: v.decode() if isinstance(v, bytes) else v)
    for l in lines:
        for k in ('Up_lev_str', 'Lo_lev_str'):
            setattr(l, k, dec(getattr(l, k)))
    iso1, _ = M.case_isomolecs(sbm)
    sp = smm.prepare_spe_grid(M.WN_RANGE).spectral_grid
    work = tempfile.mkdtemp() + os.sep
    monkeypatch.chdir(work)                  # the reference appends to ./control_spectrobot
    try:
        with R.fortran_from_source(), R.quiet():
            assert sys.modules['lineshape'].humliv_bb.__qualname__.startswith('fortran_from_source')
            q = [[spcl.CalcPartitionSum(int(m), int(i), temp=92.3) for m, i in ref["q_molisos"]]]
            shp = spcl.calc_shapes_lines(sp, [l for l in lines if l.Iso == 1], M.CELLS[0][1],
                                         M.CELLS[0][0], iso1, n_threads=1)
            lut = smm.LookUpTable(iso1, sp.wn_range(), False)
            lut.make(sp, lines, M.CELLS[:1], cartLUTs=work, n_threads=1)
        assert np.array_equal(np.array(q)[0], ref["q_values"][:, 4])
        pick = ref["shape_pick"]
        assert np.array_equal(np.array([shp[i].shape.spectrum for i in pick]), ref["shape_spectra"])
        n_set = 0
        for s, lev in enumerate(iso1.levels):
            pts, sets = M.read_stream(lut.sets[lev].filename)
            for k, ct in enumerate(M.CTYPES):
                assert np.array_equal(sets[0][ct].spectrum, ref["cells_nonlte"][s, k]), (lev, ct)
                n_set += bool(np.any(sets[0][ct].spectrum))
        assert n_set >= 6
    finally:
        shutil.rmtree(work, ignore_errors=True)


_REGEN = r"""
import sys, warnings
warnings.simplefilter('ignore')
gold, root = sys.argv[1], sys.argv[2]
sys.path.insert(0, gold)
sys.path.insert(1, root)
import ref_exec as R
R.load()
for name in ('make_ref_golden', 'make_ref_golden2', 'make_ref_golden3'):
    M = __import__(name)
    with R.fortran_from_source(), R.quiet():
        M.main()
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not on this box")
def test_every_reference_executed_fixture_set_regenerates_with_the_executed_fortran(tmp_path):
    """All three generators of the reference-executed fixtures (make_ref_golden{,2,3}.py: LUT
    builds, shapes, LTE / non-LTE cells, split files, prepare_fortran_sum, the slow line-by-line
    coefficients, ...) re-run in a scratch copy with the reference's Fortran executed from source
    in place of the C stand-ins: every array of ref_golden{,2,3}.npz comes out bit-identical.  The
    committed fixtures are therefore outputs of the reference's Python AND Fortran."""
    import shutil
    import subprocess
    scratch = tmp_path / "tests" / "golden"
    shutil.copytree(GOLD, str(scratch), ignore=shutil.ignore_patterns("__pycache__"))
    script = tmp_path / "regen.py"
    script.write_text(_REGEN)
    subprocess.run([sys.executable, str(script), str(scratch), ROOT], check=True, cwd=str(tmp_path),
                   stdout=subprocess.DEVNULL)
    n = 0
    for f in ("ref_golden.npz", "ref_golden2.npz", "ref_golden3.npz"):
        a, b = np.load(str(scratch / f), allow_pickle=True), np.load(os.path.join(GOLD, f), allow_pickle=True)
        assert sorted(a.files) == sorted(b.files)
        for k in a.files:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (f, k)
            same = np.array_equal(a[k], b[k]) if a[k].dtype.kind in "USO" else np.array_equal(a[k], b[k], equal_nan=True)
            assert same, (f, k)
            n += 1
    assert n >= 190
