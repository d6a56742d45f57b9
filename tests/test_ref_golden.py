"""CPU tests against fixtures produced by EXECUTING the reference's own Python
(tests/golden/make_ref_golden.py, ref_exec.py): the oracle restatement and the host-side product
code must reproduce what spect_classes.py / spect_main_module.py themselves compute.

Rows pinned this way: A1 window placement, A3 widths, A4 G coefficients, A5/A6 level
selection + clipping + staging-matrix sum, A7 LUT stream, A8 PT ladder, A9 Lagrange interpolation,
A10 LutSet.calculate, A11 make_abscoeff_LUTS_fast, f1 convolution / FOV / masks, f3 file formats.
The Fortran under them (A2, A13, the TIPS table) is restated, not executed - no Fortran compiler.
"""
import os
import sys

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CTYPES = ['sp_emission', 'ind_emission', 'absorption']
LEVELS = ['0 0 0 0 1A1', '0 0 1 0 1F2', '0 1 0 0 1E']
ENERGIES = [0.0, 3019.4935, 1533.3326]


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "ref_golden.npz"))


@pytest.fixture(scope="module")
def case():
    """The fixture's line list and isotopologues as PRODUCT objects (read from the same HITRAN
    file the reference read)."""
    from spectrobot_b200 import spect_base_module as sbm, spect_classes as spcl
    lines = spcl.read_line_database(os.path.join(GOLD, "ref_lines.par"))
    iso1 = sbm.IsoMolec(6, 1)
    iso1.add_levels(LEVELS, ENERGIES)
    iso1.is_in_LTE = False
    iso2 = sbm.IsoMolec(6, 2)
    return dict(lines=lines, iso1=iso1, iso2=iso2, spcl=spcl, sbm=sbm)


def test_line_reader_matches_reference_reader(ref, case):
    """read_line_database (spcl:1532-1601): same records, same numbers, same strings."""
    spcl = case["spcl"]
    assert len(case["lines"]) == len(ref["db_Freq"])
    for k in spcl.cose_hit:
        want = ref["db_" + k]
        got = np.array([getattr(l, k) for l in case["lines"]])
        if want.dtype.kind in "US":
            assert [str(w) for w in want] == [str(g) for g in got], k
        else:
            assert np.array_equal(got.astype(float), want.astype(float)), k


def test_molparam_matches_reference_file(ref, case):
    assert case["iso1"].ratio == float(ref["iso_ratio"])
    assert [case["iso1"].MM, case["iso2"].MM] == list(ref["iso_MM"])


def test_widths_and_gcoeffs_host_and_oracle(ref, case, oracle):
    """A3 + A4: CheckWidths / Calc_Gcoeffs of the product's SpectLine and of the oracle (Python and
    C) against the reference's values."""
    iso1, iso2 = case["iso1"], case["iso2"]
    for n, (T, P) in enumerate(ref["phys_PT"]):
        for i, lin in enumerate(case["lines"]):
            MM = iso1.MM if lin.Iso == 1 else iso2.MM
            dw, lw, sh = ref["phys_widths"][n, i]
            got = lin.CheckWidths(T, P, MM)
            assert got[0] == pytest.approx(dw, rel=1e-15) and got[1] == pytest.approx(lw, rel=1e-15)
            assert got[2] == pytest.approx(sh, rel=1e-15)
            lw_c, dw_c = oracle.widths_c(lin.Freq, lin.Air_broad, lin.T_dep_broad, T, P, MM)
            assert dw_c == pytest.approx(dw, rel=1e-14) and lw_c == pytest.approx(lw, rel=1e-14)
            G = lin.Calc_Gcoeffs(T, isomolec=iso1 if lin.Iso == 1 else None)
            want = ref["phys_gcoeff"][n, i]
            for j, ct in enumerate(CTYPES):
                assert G[ct] == pytest.approx(want[j], rel=1e-13), (i, ct)
            ok = bool(ref["phys_link_ok"][i])
            assert (lin.Up_lev_id is not None and lin.Lo_lev_id is not None) == ok
            evu, evl = (lin.E_vib_up, lin.E_vib_lo) if ok else (0.0, 0.0)
            gc = oracle.gcoeffs_c(lin.Freq, lin.A_coeff, lin.E_lower, lin.g_up, lin.g_lo, evu, evl, T)
            assert np.allclose(gc, want, rtol=1e-12, atol=0.0)


def test_partition_sum_and_strengths(ref, case, oracle):
    """A9: CalcPartitionSum (4-point / 3-point Lagrange) - oracle Python, oracle C and the
    library's host table; CalcStrength_at_T, CalcStrength_from_Einstein, calc_A_coeff_from_strength
    of the product's SpectLine."""
    from spectrobot_b200 import engine
    for (m, i), row in zip(ref["q_molisos"], ref["q_values"]):
        for t, q in zip(ref["q_temps"], row):
            assert oracle.CalcPartitionSum(int(m), int(i), float(t)) == pytest.approx(q, rel=1e-13)
            # the reference evaluates scipy's poly1d (monomial coefficients): its own rounding
            # noise against the direct Lagrange form used in C is ~1e-11 relative
            assert oracle.partition_sum_c(int(m), int(i), float(t)) == pytest.approx(q, rel=1e-9)
            assert engine.partition_sum(int(m), int(i), float(t)) == pytest.approx(q, rel=1e-9)
    iso1 = case["iso1"]
    for n, (T, P) in enumerate(ref["phys_PT"]):
        for i, lin in enumerate(case["lines"]):
            assert lin.CalcStrength(T) == pytest.approx(ref["phys_strength_T"][n, i], rel=1e-10)
            want = ref["phys_strength_einstein"][n, i]
            if np.isnan(want[0]):
                continue
            lin.E_vib_up = lin.E_vib_lo = None
            got = lin.CalcStrength_from_Einstein(T, isomolec=iso1 if lin.Iso == 1 else None)
            assert got[0] == pytest.approx(want[0], rel=1e-10)
            assert got[1] == pytest.approx(want[1], rel=1e-10)
    for i, lin in enumerate(case["lines"]):
        assert lin.calc_A_coeff_from_strength() == pytest.approx(ref["phys_A_from_strength"][i],
                                                                 rel=1e-10)


def _tab(case, iso):
    spcl = case["spcl"]
    lines = [l for l in case["lines"] if l.Iso == iso]
    return lines, spcl.line_table(lines, case["iso1"] if iso == 1 else None)


def test_oracle_cell_matches_reference_built_lut(ref, case, oracle):
    """A1 + A5 + A6 + A7: the C oracle's whole-cell routine (window on the nearest grid point,
    level selection by integer id, clipping at the spectrum edges, summation) against the LUT the
    reference built through calc_shapes_lines -> LutSet.add_PT -> BuildCoeff ->
    add_lines_to_spectrum -> prepare_fortran_sum -> sum_all_lines and wrote to its pickle stream."""
    grid = ref["grid"]
    lines, tab = _tab(case, 1)
    assert np.array_equal(tab["up_set"] >= 0, ref["phys_link_ok"][[l.Iso == 1 for l in case["lines"]]])
    (P0, T0), (P1, T1) = ref["cells_PT"]
    got = oracle.gcoeff_cell(tab, grid, T0, P0, case["iso1"].MM, 3)
    want = ref["cells_nonlte"]
    for s in range(3):
        for k in range(3):
            if not np.any(want[s, k]):
                assert not np.any(got[s, k])
            else:
                assert rel_err(got[s, k], want[s, k]) < 1e-12, (s, k)
    got = oracle.gcoeff_cell(tab, grid, T1, P1, case["iso1"].MM, 3)[:, :, ::8]
    assert rel_err(got, ref["cells_nonlte_b8"]) < 1e-12
    lines2, tab2 = _tab(case, 2)
    P, T = ref["cells_lte_PT"]
    got = oracle.gcoeff_cell(tab2, grid, T, P, case["iso2"].MM, 1)[0]
    assert rel_err(got, ref["cells_lte"]) < 1e-12


def test_oracle_line_window_and_shape(ref, case, oracle):
    """A1/A3: closest_grid index, first abscissa of the 13010-point window and the normalised
    shape (K / (dw sqrt(pi/ln2)), spcl:1990-2008) of single lines."""
    grid = ref["grid"]
    lin_grid = oracle.line_window_offsets(grid)
    P, T = ref["cells_PT"][0]
    lines = [l for l in case["lines"] if l.Iso == 1 and l.LinkToMolec(case["iso1"])]
    assert [l.Freq for l in lines] == list(ref["shape_freq"])
    for i, l in enumerate(lines):
        ind, val = oracle.closest_grid(grid, l.Freq)
        assert ind == ref["shape_centre"][i]
        assert lin_grid[0] + val == ref["shape_first"][i]
    for j, i in enumerate(ref["shape_pick"]):
        l = lines[int(i)]
        lw, dw = oracle.widths_c(l.Freq, l.Air_broad, l.T_dep_broad, T, P, case["iso1"].MM)
        shp = oracle.line_shape(l.Freq, lw, dw, grid[ref["shape_centre"][int(i)]], lin_grid)
        assert rel_err(shp, ref["shape_spectra"][j]) < 1e-13


def test_lut_interpolation_rule(ref, oracle):
    """A10: LutSet.calculate + SpectralGcoeff.interpolate - NumPy restatement, C restatement of the
    node choice, and the library's host helper sr_lut_weights."""
    from spectrobot_b200 import engine
    g = ref["interp_g32"].astype(np.float64)
    PT = ref["interp_PT"]
    for i, (p, t) in enumerate(ref["interp_probes"]):
        cell, w = oracle.lut_weights(PT, p, t)
        cell2, w2 = engine.lut_weights(PT, p, t)
        assert np.array_equal(cell, cell2) and np.allclose(w, w2, rtol=1e-15, atol=0)
        for s in range(g.shape[1]):
            for k in range(3):
                want = ref["interp_result"][i, s, k]
                sets = [g[c, s, k] if np.any(g[:, s, k]) else None for c in range(len(PT))]
                got = oracle.LutSet_calculate(PT, sets, p, t)
                if got is None:
                    assert not np.any(want)
                    continue
                assert rel_err(got, want) < 1e-14
                blend = sum(w[j] * g[cell[j], s, k] for j in range(4))
                assert rel_err(blend, want) < 1e-12
    assert bool(ref["interp_extrap_raises"])
    with pytest.raises(ValueError):
        oracle.lut_weights(PT, 3.0, 150.0)


def test_abscoeff_assembly(ref, oracle):
    """A11: make_abscoeff_LUTS_fast - NumPy restatement and the C LOS routine (tau/J per unit
    column and unit isotopic ratio) against the reference, non-LTE and LTE populations."""
    lut = dict(g32=ref["interp_g32"], pt=ref["interp_PT"], level_energy=np.array(ENERGIES), mol=6,
               iso=1, iso_ratio=1.0, lte_unidentified=False)
    temps, press = ref["interp_probes"][:, 1], ref["interp_probes"][:, 0]
    for tv, want in ((ref["abscoeff_tvib"], ref["abscoeff_nonlte"]), (None, ref["abscoeff_lte"])):
        a, e = oracle.make_abscoeff_LUTS_fast(lut, temps, press, tv)
        assert rel_err(a, want[0], 1e-12) < 1e-12 and rel_err(e, want[1]) < 1e-12
        n = len(temps)
        tvib = None if tv is None else tv[None, :, None, :]
        rad, tau, src = oracle.los_rt([lut], [n], temps[None], press[None], np.ones((1, 1, n)),
                                      tvib, materialise=True)
        # orc_los_rt materialises tau = abs*column and S = J/tau
        assert rel_err(tau[0], want[0], 1e-9) < 1e-9
        assert rel_err(src[0] * tau[0], want[1], 1e-9) < 1e-9


def test_PT_couples_match_reference(ref, case):
    """A8: calc_PT_couples_atmosphere on the fixture atmosphere, three option sets."""
    from spectrobot_b200 import spect_main_module as smm
    sbm = case["sbm"]
    atm = sbm.AtmProfile(sbm.AtmGrid(['lat', 'alt'], [[-90, -30, 30, 90], ref["atm_z"]]),
                         ref["atm_temp"], 'temp', 'lin')
    atm.add_profile(ref["atm_pres"], 'pres', 'exp')
    opts = dict(a=dict(pres_step_log=1.0, temp_step=5.0, max_pres=2.0),
                b=dict(pres_step_log=0.4, temp_step=5.0, max_pres=0.1),
                c=dict(pres_step_log=1.0, temp_step=5.0, max_pres=2.5, add_lowpres=False))
    for tag, kw in opts.items():
        got = np.array(smm.calc_PT_couples_atmosphere(case["lines"], [case["iso1"], case["iso2"]],
                                                      atm, **kw))
        want = ref["ptc_" + tag]
        assert got.shape == want.shape, tag
        assert np.allclose(got, want, rtol=1e-14, atol=0), tag


def test_convolution_fov_and_masks(ref, oracle):
    """f1: Gaussian convolution (oracle restatement), FOV integration (closed form of the product
    and the oracle's spline + quad) and the retrieval masks."""
    from spectrobot_b200 import spect_main_module as smm
    got = oracle.convolve_to_grid_from_irregular(ref["conv_grid"], ref["conv_spec"],
                                                 ref["conv_cm_centres"], ref["conv_cm_widths"])
    assert rel_err(got, ref["conv_cm_result"]) < 1e-13
    # nm observation: grid -> 1e7/grid reversed, spectrum * grid^2 * 1e-7 reversed (spcl:771-778),
    # ergscm2 -> Wm2 = 1e-3 (spcl:1216-1221)
    x = ref["conv_grid"]
    got = oracle.convolve_to_grid_from_irregular((1.e7 / x)[::-1], (ref["conv_spec"] * x ** 2 * 1e-7)[::-1],
                                                 ref["conv_nm_centres"], ref["conv_nm_widths"]) * 1e-3
    assert rel_err(got, ref["conv_nm_result"]) < 1e-12
    assert list(ref["conv_nm_units"]) == ['Wm2', 'nm']
    for r, want in zip(ref["fov_rot"], ref["fov_out"]):
        assert rel_err(smm.fov_integrate(ref["fov_in"], r), want) < 1e-7     # quad's own accuracy
        assert rel_err(oracle.FOV_integr_1D(ref["fov_in"], ref["conv_cm_centres"][:5], r), want) < 1e-12
    z = ref["tri_z"]
    assert np.array_equal(smm.alt_triangle(z, 550.0, node_lo=450.0, node_up=700.0).mask, ref["tri_mid"])
    assert np.array_equal(smm.alt_triangle(z, 350.0, node_up=450.0, first=True).mask, ref["tri_first"])
    assert np.array_equal(smm.alt_triangle(z, 950.0, node_lo=850.0, last=True).mask, ref["tri_last"])
    assert np.array_equal(smm.alt_triangle(z, 600.0, step=100.0).mask, ref["tri_step"])
    for la, want in zip(ref["latbox_probe"], ref["latbox"]):
        assert np.array_equal(smm.lat_box(ref["latbox_limits"], la).mask, want)


def test_reference_written_lut_files_load(ref, case):
    """f3: a per-level LUT stream and a split/compressed LUT file, both WRITTEN BY THE REFERENCE
    (LutSet.add_PT, smm:1161; split_and_compress_LUTS, smm:1614-1728), read by the product."""
    from spectrobot_b200 import spect_main_module as smm
    pts, sets = smm.read_lutset_stream(os.path.join(GOLD, "ref_LUT_mol06_iso1_nonLTE_lev_01.pic"))
    assert np.array_equal(np.array(pts), ref["cells_PT"])
    for k, ct in enumerate(CTYPES):
        assert np.array_equal(sets[0][ct].spectrum, ref["cells_nonlte"][1, k])
        assert np.array_equal(sets[1][ct].spectrum[::8], ref["cells_nonlte_b8"][1, k])
        assert sets[0][ct].ctype == ct and sets[0][ct].spectral_grid is None
    split = smm.read_split_file(os.path.join(GOLD, "ref_LUT_csplit01_mol06_iso1_nonLTE.pic"))
    assert sorted(split) == ['lev_00', 'lev_01', 'lev_02']
    st = split['lev_01']
    assert len(st.spectral_grid.grid) == ref["split_lens"][1]
    assert st.sets[0]['sp_emission'].spectrum.dtype == np.float32
    assert np.array_equal(st.sets[0]['sp_emission'].spectrum.astype(float), ref["split1_lev01_cell0_sp"])
    for c in range(2):
        for s, lev in enumerate(sorted(split)):
            for k, ct in enumerate(CTYPES):
                assert (split[lev].sets[c][ct] is None) == bool(ref["split1_none"][c, s, k])


@pytest.mark.skipif(not os.path.exists('/root/reference/spect_classes.py'),
                    reason="reference tree not present (GPU box)")
def test_fixtures_are_what_the_reference_computes_now(ref, tmp_path):
    """Re-executes a slice of the reference live (this container only) and compares with the
    committed fixture: the fixture is the reference's output, not a hand edit."""
    sys.path.insert(0, GOLD)
    import ref_exec as R
    import make_ref_golden as M
    spcl, smm, sbm = R.load()
    lines = spcl.read_line_database(os.path.join(GOLD, "ref_lines.par"))
    dec = (lambda v: v.decode() if isinstance(v, bytes) else v)
    for l in lines:
        for k in ('Up_lev_str', 'Lo_lev_str'):
            setattr(l, k, dec(getattr(l, k)))
    iso1, _ = M.case_isomolecs(sbm)
    T, P = ref["phys_PT"][1]
    with R.quiet():
        g = np.array([[lin.Calc_Gcoeffs(T, isomolec=iso1 if lin.Iso == 1 else None)[c]
                       for c in CTYPES] for lin in lines])
    assert np.array_equal(g, ref["phys_gcoeff"][1])
    w = np.array([lin.CheckWidths(T, P, iso1.MM) for lin in lines if lin.Iso == 1])
    assert np.array_equal(w, ref["phys_widths"][1][[l.Iso == 1 for l in lines]])
    assert spcl.CalcPartitionSum(6, 1, temp=92.3) == ref["q_values"][0, 4]
