"""Fourth set of reference-executed fixtures (tests/golden/f77_golden.npz): the reference's
FORTRAN routines on the hot path - lineshape.f humliv_bb (A2), sum_all_lines (inner loop of A6),
humli_bb, curgods.f curgod_fort_1..4 (A13) and fparts_mod.f bd_tips_2003 with the QT_* routines it
calls (the table of A9) - executed FROM THEIR OWN SOURCE TEXT by the
mechanical FORTRAN 77 executor f77_exec.py (no Fortran compiler exists in this image; see its
header for the language rules it applies).  Run from the repo root in a container that has
/root/reference:

    python tests/golden/make_f77_golden.py

Inputs are stored next to the outputs, so the tests need neither /root/reference nor this script.
humliv_bb outputs are stored on the index subset `keep` (every 32nd point, the first and last
600 points, the 1210 points around the window centre) to keep the file small; three cases are
stored whole.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import f77_exec as F  # noqa: E402

REF = os.environ.get("SPECTROBOT_REFERENCE", "/root/reference")
IMXSIG = 13010


def keep_indices():
    return np.unique(np.r_[np.arange(0, IMXSIG, 32), np.arange(0, 600), np.arange(IMXSIG - 600, IMXSIG),
                           np.arange(5900, 7110)])


def window(w0, w1, step=5.e-4):
    """Grid and window offsets exactly as the reference forms them (smm:1267, spcl:1445-1446)."""
    grid = np.arange(w0, w1 + step / 2, step, dtype=float)
    st = grid[1] - grid[0]
    lin = np.arange(-IMXSIG * st / 2, IMXSIG * st / 2, st, dtype=float)
    assert len(lin) == IMXSIG
    return grid, lin


def humliv_cases():
    """(tag, centre, i1, i2, x0, lw, dw, on_gpu).  The first 15 + 4 are the inputs of
    tests/test_gpu_voigt.py::test_humliv_bb_inside_branch / ::test_humliv_bb_other_branches (the
    CUDA drop-in is compared with these reference-executed values on the B200); the rest widen the
    parameter range for the CPU restatement."""
    from oracle import cpu_oracle as O
    cases = []
    grid, lin = window(2990.0, 3010.0)
    nu0 = 3000.1234567
    c = grid[np.argmin(np.abs(grid - nu0))]
    for T in (70.0, 150.0, 296.0):
        for P in (1e-6, 1e-3, 0.1, 2.5, 1000.0):
            lw, dw = O.widths_c(nu0, 0.06, 0.7, T, P, 16.0313)     # inputs only: two numbers
            cases.append(("inside_P%g_T%g" % (P, T), c, 1, IMXSIG, nu0, lw, dw / np.sqrt(np.log(2.0)), 1))
    x = lin + 3000.0
    lw, dw = 2.0e-3, 4.0e-3
    cases += [("forward", 3000.0, 1, IMXSIG, x[0] - 0.0107, lw, dw, 1),
              ("backward", 3000.0, 1, IMXSIG, x[-1] + 0.0031, lw, dw, 1),
              ("sub_forward", 3000.0, 6600, 12000, 3000.00013, lw, dw, 1),
              ("sub_inside", 3000.0, 3000, 9000, 3000.00013, lw, dw, 1)]
    rng = np.random.default_rng(20067)
    for k in range(10):                                      # wider: CPU restatement only
        x0 = float(rng.uniform(2853.3, 3446.7))
        cc = float(np.round(x0 / 5e-4) * 5e-4)
        lw = float(10.0 ** rng.uniform(-9.0, 0.5))
        dw = float(rng.uniform(1.5e-3, 7e-3))
        cases.append(("wide_%d" % k, cc, 1, IMXSIG, x0, lw, dw, 0))
    for k, (lw, dw) in enumerate([(1e-3, 3e-3), (0.08, 2.5e-3), (1e-6, 5e-3), (0.5, 3e-3)]):
        cases += [("left_%d" % k, 3000.0, 1, IMXSIG, x[0] - 0.01, lw, dw, 0),
                  ("left_edge_%d" % k, 3000.0, 1, IMXSIG, x[0], lw, dw, 0),
                  ("left_sub_%d" % k, 3000.0, 100, 9000, x[99] - 1e-3, lw, dw, 0),
                  ("right_%d" % k, 3000.0, 1, IMXSIG, x[-1] + 0.01, lw, dw, 0),
                  ("right_edge_%d" % k, 3000.0, 1, IMXSIG, x[-1], lw, dw, 0),
                  ("right_sub_%d" % k, 3000.0, 200, 12000, x[11999] + 2e-3, lw, dw, 0),
                  ("off_centre_%d" % k, 3000.0, 1, IMXSIG, 2998.2, lw, dw, 0),
                  ("narrow_sub_%d" % k, 3000.0, 6400, 6600, 3000.0, lw, dw, 0)]
    return lin, cases


def main():
    ls = F.load(os.path.join(REF, "lineshape.f"))
    cg = F.load(os.path.join(REF, "curgods.f"))
    out = dict()

    # --- humliv_bb (lineshape.f:226-569)
    lin, cases = humliv_cases()
    keep = keep_indices()
    out["lin"] = lin
    out["keep"] = keep
    out["hv_tags"] = np.array([c[0] for c in cases])
    out["hv_in"] = np.array([c[1:7] for c in cases], dtype=float)     # centre, i1, i2, x0, lw, dw
    out["hv_gpu"] = np.array([c[7] for c in cases], dtype=np.int8)
    ys = np.zeros((len(cases), len(keep)))
    whole = dict()
    for k, (tag, c, i1, i2, x0, lw, dw, _) in enumerate(cases):
        y = np.zeros(IMXSIG)
        ls["humliv_bb"](lin + c, i1, i2, x0, lw, dw, y)
        ys[k] = y[keep]
        if tag in ("inside_P0.1_T150", "forward", "wide_3"):
            whole[tag] = y
    out["hv_y"] = ys
    out["hv_whole_tags"] = np.array(sorted(whole))
    out["hv_whole_y"] = np.array([whole[t] for t in sorted(whole)])
    stops = []
    for args in ((10, 5, 0.5, 1e-3, 1e-3), (1, IMXSIG, 0.5, 1e-3, 0.0), (1, IMXSIG, 0.5, 1e-3, -1.0)):
        try:
            ls["humliv_bb"](np.linspace(0, 1, IMXSIG), *args, np.zeros(IMXSIG))
            stops.append("")
        except F.F77Stop as e:
            stops.append(str(e))
    out["hv_stop_msgs"] = np.array(stops)

    # --- humli_bb (lineshape.f:150-205; exported by the f2py module, unused by the Python layer)
    rng = np.random.default_rng(7)
    rx = np.r_[rng.uniform(-25.0, 25.0, 150), 0.0, 5.4, 5.6, 14.9, 15.1]
    ry = np.r_[10.0 ** rng.uniform(-6.0, 1.3, 150), 1e-3, 0.05, 0.05, 0.05, 0.05]
    out["hb_rx"], out["hb_ry"] = rx, ry
    out["hb_rre"] = np.array([ls["humli_bb"](a, b, 0.0)["rre"] for a, b in zip(rx, ry)])

    # --- sum_all_lines (lineshape.f:2-25): overlapping windows, line order matters for rounding
    rng = np.random.default_rng(11)
    n_lines, n_win, ld, n_spe = 40, 600, 48, 5000
    m = np.zeros((ld, n_win), order="F")
    m[:n_lines] = rng.standard_normal((n_lines, n_win)).astype(np.float32)
    m[:n_lines] *= 10.0 ** rng.uniform(-6, 6, (n_lines, 1)).astype(np.float32)
    m[...] = m.astype(np.float32)                            # float32-valued, stored losslessly below
    init = rng.integers(1, n_spe - n_win, n_lines).astype(np.int32)
    fin = (init + rng.integers(0, n_win, n_lines)).astype(np.int32)
    init[:4] = 1                                              # windows that start at the first point
    fin[:4] = init[:4] + np.array([0, 1, 299, 599])
    fin[4:6] = n_spe                                          # and end at the last one
    init[4:6] = n_spe - np.array([0, 599])
    spe = rng.standard_normal(n_spe).astype(np.float32).astype(float)
    res = np.empty(n_spe)
    ls["sum_all_lines"](spe, m, init, fin, n_lines, n_spe, res)
    assert np.array_equal(m.astype(np.float32).astype(float), m)
    out["sal_matrix"] = m.astype(np.float32)
    out["sal_init"], out["sal_fin"], out["sal_spe"], out["sal_res"] = init, fin, spe.astype(np.float32), res
    res0 = np.empty(n_spe)
    ls["sum_all_lines"](spe, m, init, fin, 0, n_spe, res0)   # n_lines = 0: a copy
    assert np.array_equal(res0, spe)

    # --- curgod_fort_1..4 (curgods.f:2-98)
    rng = np.random.default_rng(13)
    cg_in, cg_res = [], []
    for n_p in (2, 3, 17, 120):
        for trial in range(3):
            x = np.cumsum(rng.uniform(1.0e5, 1.0e6, n_p))                 # cm
            nd = 1e13 * np.exp(-x / 6.0e6) * rng.uniform(0.9, 1.1, n_p)
            vmr = 1e-2 * (1 + 0.3 * np.sin(x / 5.0e6)) * rng.uniform(0.95, 1.05, n_p)
            f = 150.0 + 20 * np.cos(x / 7.0e6) + rng.uniform(-1, 1, n_p)
            pad = lambda a: np.r_[a, np.zeros(120 - n_p)]   # noqa: E731
            cg_in.append(np.array([pad(nd), pad(vmr), pad(f), pad(x)]))
            cg_res.append([cg["curgod_fort_1"](nd, x, n_p, 0.0)["res"],
                           cg["curgod_fort_2"](nd, vmr, x, n_p, 0.0)["res"],
                           cg["curgod_fort_3"](nd, vmr, f, x, n_p, 0.0)["res"],
                           cg["curgod_fort_4"](nd, vmr, f, x, n_p, 0.0)["res"]])
    # the inputs of tests/test_gpu_voigt.py::test_curgod (compared with the CUDA drop-in on the B200)
    rng = np.random.default_rng(2)
    n_p = 120
    x = np.cumsum(rng.uniform(1.0, 10.0, n_p))
    nd = 1e12 * np.exp(-x / 80.0) * rng.uniform(0.9, 1.1, n_p)
    vmr = 1e-2 * (1 + 0.3 * np.sin(x / 50.0))
    f = 150.0 + 20 * np.cos(x / 70.0)
    cg_in.append(np.array([nd, vmr, f, x]))
    cg_res.append([cg["curgod_fort_1"](nd, x, n_p, 0.0)["res"],
                   cg["curgod_fort_2"](nd, vmr, x, n_p, 0.0)["res"],
                   cg["curgod_fort_3"](nd, vmr, f, x, n_p, 0.0)["res"],
                   cg["curgod_fort_4"](nd, vmr, f, x, n_p, 0.0)["res"]])
    out["cg_n_p"] = np.r_[np.repeat([2, 3, 17, 120], 3), 120]
    out["cg_gpu_case"] = np.array(len(cg_in) - 1)
    out["cg_in"] = np.array(cg_in)
    out["cg_res"] = np.array(cg_res)

    # --- bd_tips_2003 (fparts_mod.f:33-295 + the QT_* routines it calls): every (mol, iso) the
    # dispatcher reaches; Q values are REAL*4 literals stored in DOUBLE PRECISION arrays, so they
    # are float32-valued and stored as float32 without loss
    fp = F.load(os.path.join(REF, "fparts_mod.f"))
    keys, gis, qs, t_grid = [], [], [], None
    for mol in range(1, 60):
        for iso in range(1, 20):
            t, q = np.zeros(119), np.full(119, -7.0)
            try:
                r = fp["bd_tips_2003"](mol, iso, -7.0, t, q)
            except IndexError:                   # iso beyond the DIMENSION of the routine's tables
                break
            except F.F77Unsupported:             # QT_H3P (mol 42): COMMON storage association
                break
            if np.all(q == -7.0):                # no branch of the dispatcher took this molecule
                break
            assert np.array_equal(q.astype(np.float32).astype(float), q)
            keys.append((mol, iso))
            gis.append(r["gi"])
            qs.append(q.astype(np.float32))
            assert t_grid is None or np.array_equal(t, t_grid)
            t_grid = t
    out["tips_keys"] = np.array(keys, dtype=np.int32)
    out["tips_gi"] = np.array(gis)
    out["tips_q"] = np.array(qs)
    out["tips_t"] = t_grid

    fn = os.path.join(HERE, "f77_golden.npz")
    np.savez_compressed(fn, **out)
    print("wrote", fn, os.path.getsize(fn), "bytes;", len(cases), "humliv_bb cases,", len(keys), "TIPS tables")


if __name__ == "__main__":
    main()
