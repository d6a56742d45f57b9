"""GPU parity of the cross-section path (K1/K2 + Tier-1 drop-ins) against the CPU oracle.
All calls go through the C ABI (libspectrobot.so); tolerance 1e-6 relative (north_star)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_XS = 1e-6   # cross sections, BASELINE.json north_star


@pytest.fixture(scope="module")
def sb():
    from spectrobot_b200 import engine, lineshape, synthetic, _lib
    assert _lib.cuda_available(), "libspectrobot.so sees no CUDA device"
    return engine, lineshape, synthetic


def _window(oracle, S, w0=2995.0, w1=3005.0):
    g = S.spectral_grid(w0, w1)
    return g, oracle.line_window_offsets(g)


@pytest.mark.parametrize("P", [1e-6, 1e-3, 0.1, 2.5, 1000.0])
@pytest.mark.parametrize("T", [70.0, 150.0, 296.0])
def test_humliv_bb_inside_branch(sb, oracle, P, T):
    engine, lineshape, S = sb
    g, L = _window(oracle, S, 2990.0, 3010.0)
    nu0 = 3000.1234567
    ind, c = oracle.closest_grid(g, nu0)
    lw, dw = oracle.widths_c(nu0, 0.06, 0.7, T, P, 16.0313)
    x = L + c
    dwp = dw / np.sqrt(np.log(2.0))
    ref = oracle.humliv_bb(x, 1, 13010, nu0, lw, dwp)
    got = lineshape.humliv_bb(x, 1, 13010, nu0, lw, dwp)
    assert rel_err(got, ref) < TOL_XS
    # away from the float32-rounded core the two must agree to rounding
    reg = oracle.humliv_regions(x, 1, 13010, nu0, lw, dwp)
    wing = np.r_[0:reg[0] - 1, reg[1]:13010]
    assert rel_err(got[wing], ref[wing]) < 1e-9   # closed-form vs running xrun (lineshape.f:467)


@pytest.mark.parametrize("case", ["forward", "backward", "sub_forward", "sub_inside"])
def test_humliv_bb_other_branches(sb, oracle, case):
    """x0 left / right of the window (lineshape.f:272-442) and i1,i2 sub-ranges."""
    engine, lineshape, S = sb
    g, L = _window(oracle, S)
    lw, dw = 2.0e-3, 4.0e-3
    if case == "forward":
        x = L + 3000.0; x0 = x[0] - 0.0107; i1, i2 = 1, 13010
    elif case == "backward":
        x = L + 3000.0; x0 = x[-1] + 0.0031; i1, i2 = 1, 13010
    elif case == "sub_forward":
        x = L + 3000.0; x0 = 3000.00013; i1, i2 = 6600, 12000
    else:
        x = L + 3000.0; x0 = 3000.00013; i1, i2 = 3000, 9000
    ref = oracle.humliv_bb(x, i1, i2, x0, lw, dw)
    got = lineshape.humliv_bb(x, i1, i2, x0, lw, dw)
    assert np.all(got[:i1 - 1] == 0) and np.all(got[i2:] == 0)
    assert rel_err(got, ref) < TOL_XS


def test_humliv_bb_error_convention(sb):
    engine, lineshape, S = sb
    from spectrobot_b200._lib import SpectrobotError, SR_ERR_ARG, SR_ERR_DW
    x = np.linspace(0, 1, 13010)
    with pytest.raises(SpectrobotError) as e:
        lineshape.humliv_bb(x, 10, 5, 0.5, 1e-3, 1e-3)       # lineshape.f:253-256
    assert e.value.code == SR_ERR_ARG
    with pytest.raises(SpectrobotError) as e:
        lineshape.humliv_bb(x, 1, 13010, 0.5, 1e-3, 0.0)      # lineshape.f:260-264
    assert e.value.code == SR_ERR_DW
    with pytest.raises(ValueError):
        lineshape.humliv_bb(x[:100], 1, 100, 0.5, 1e-3, 1e-3)  # f2py fixed shape


def test_sum_all_lines(sb, oracle):
    engine, lineshape, S = sb
    rng = np.random.default_rng(1)
    n_lines, n_win, ld, n_spe = 37, 301, 50, 5000
    m = np.zeros((ld, n_win), order="F")
    m[:n_lines] = rng.uniform(size=(n_lines, n_win))
    init = rng.integers(1, n_spe - n_win, n_lines).astype(np.int32)
    fin = (init + rng.integers(0, n_win, n_lines)).astype(np.int32)
    spe = rng.uniform(size=n_spe)
    ref = oracle.sum_all_lines(spe, m, init, fin, n_lines)
    got = lineshape.sum_all_lines(spe, m, init, fin, n_lines, n_spe)
    assert rel_err(got, ref) < 1e-13


def test_curgod(sb, oracle):
    from spectrobot_b200 import curgods
    rng = np.random.default_rng(2)
    n_p = 120
    x = np.cumsum(rng.uniform(1.0, 10.0, n_p))
    nd = 1e12 * np.exp(-x / 80.0) * rng.uniform(0.9, 1.1, n_p)
    vmr = 1e-2 * (1 + 0.3 * np.sin(x / 50.0))
    f = 150.0 + 20 * np.cos(x / 70.0)
    got = [curgods.curgod_fort_1(nd, x, n_p), curgods.curgod_fort_2(nd, vmr, x, n_p),
           curgods.curgod_fort_3(nd, vmr, f, x, n_p), curgods.curgod_fort_4(nd, vmr, f, x, n_p)]
    ref = [oracle.curgod(1, nd, x), oracle.curgod(2, nd, vmr, x), oracle.curgod(3, nd, vmr, f, x),
           oracle.curgod(4, nd, vmr, f, x)]
    # the closed forms cancel heavily (curgods.f:66-69), so compare against the column scale
    for k in range(4):
        assert abs(got[k] - ref[k]) <= 1e-9 * abs(ref[k]), (k, got[k], ref[k])


def _cell_case(S, n_lines, w0, w1, n_levels, seed=20067, frac_unlinked=0.0):
    g = S.spectral_grid(w0, w1)
    lines = S.line_table(n_lines, w0 - 3.0, w1 + 3.0, n_levels=n_levels, seed=seed,
                         frac_unlinked=frac_unlinked)
    return g, lines


@pytest.mark.parametrize("P,T", [(1e-6, 150.0), (1e-3, 175.0), (0.1, 70.0), (2.5, 180.0),
                                 (1000.0, 94.0)])
def test_gcoeff_cell_nonlte_parity(sb, oracle, P, T):
    """12-level non-LTE cell; lines also sit closer than half a window to both grid edges
    (clipping of spect_classes.py:1113-1140) and 10% are unlinked (dropped, :1384-1388)."""
    engine, lineshape, S = sb
    g, lines = _cell_case(S, 600, 2996.0, 3004.0, 12, frac_unlinked=0.1)
    ref = oracle.gcoeff_cell(lines, g, T, P, S.CH4_MM, 12)
    ls = engine.LineSet(lines, g, S.CH4_MM, 12)
    got = ls.gcoeff_cells([[P, T]]).cpu().numpy()[0]
    assert got.shape == ref.shape
    for s in range(12):
        for ct in range(3):
            if np.max(np.abs(ref[s, ct])) == 0.0:
                assert np.all(got[s, ct] == 0.0)
            else:
                assert rel_err(got[s, ct], ref[s, ct]) < TOL_XS, (s, ct)
    # host-buffer entry point gives the same numbers
    got_h = ls.gcoeff_cells_host([[P, T]])[0]
    # rows fed by several group chunks are combined with FP64 atomics: order-of-addition noise
    assert rel_err(got_h, got) < 1e-13


def test_gcoeff_cell_lte_single_set(sb, oracle):
    """LTE isotopologue: one set 'all', E_vib = 0 (spect_main_module.py:742-748)."""
    engine, lineshape, S = sb
    g, lines = _cell_case(S, 400, 2050.0, 2058.0, 1, seed=5)
    ref = oracle.gcoeff_cell(lines, g, 150.0, 0.5, 27.994915, 1)
    ls = engine.LineSet(lines, g, 27.994915, 1)
    got = ls.gcoeff_cells([[0.5, 150.0]]).cpu().numpy()[0]
    for ct in range(3):
        assert rel_err(got[0, ct], ref[0, ct]) < TOL_XS


def test_gcoeff_multi_cell_batch_and_f32(sb, oracle):
    engine, lineshape, S = sb
    g, lines = _cell_case(S, 300, 2998.0, 3003.0, 6, seed=9)
    cells = [[1e-4, 145.0], [1e-2, 150.0], [1.0, 155.0]]
    ls = engine.LineSet(lines, g, S.CH4_MM, 6)
    got = ls.gcoeff_cells(cells).cpu().numpy()
    for i, (P, T) in enumerate(cells):
        ref = oracle.gcoeff_cell(lines, g, T, P, S.CH4_MM, 6)
        assert rel_err(got[i], ref) < TOL_XS
    g32 = ls.gcoeff_cells_f32(cells).cpu().numpy()
    assert g32.dtype == np.float32
    same = g32 == got.astype(np.float32)                   # spect_classes.py:732
    assert same.mean() > 0.999 and rel_err(g32.astype(float), got) < 1e-7


def test_gcoeff_empty_and_edge_inputs(sb, oracle):
    engine, lineshape, S = sb
    g = S.spectral_grid(2999.0, 3001.0)
    # no lines at all -> zeros
    lines = S.line_table(5, 2990.0, 3010.0, n_levels=3)
    lines["up_set"][:] = -1
    ls = engine.LineSet(lines, g, S.CH4_MM, 3)
    assert ls.n_active == 0
    assert float(ls.gcoeff_cells([[0.1, 150.0]]).abs().max()) == 0.0
    # A_coeff == 0 or g == 0 -> zero G coefficients (spect_classes.py:326-337)
    lines = S.line_table(50, 2995.8, 3004.2, n_levels=3, seed=3)   # within the grid
    lines["a_coeff"][::2] = 0.0
    lines["g_lo"][1::4] = 0.0
    ref = oracle.gcoeff_cell(lines, g, 160.0, 0.01, S.CH4_MM, 3)
    got = engine.LineSet(lines, g, S.CH4_MM, 3).gcoeff_cells([[0.01, 160.0]]).cpu().numpy()[0]
    assert rel_err(got, ref) < TOL_XS


def test_line_shapes_and_centres(sb, oracle):
    engine, lineshape, S = sb
    g, lines = _cell_case(S, 40, 2999.0, 3002.0, 4, seed=11)
    ls = engine.LineSet(lines, g, S.CH4_MM, 4)
    cen = ls.centres()
    for i in range(40):
        assert cen[i] == oracle.closest_grid(g, lines["freq"][i])[0]
    shapes, gco = ls.line_shapes(0.3, 165.0)
    shapes = shapes.cpu().numpy(); gco = gco.cpu().numpy()
    order = ls.order()
    L = oracle.line_window_offsets(g)
    for pos in range(0, len(order), 7):
        i = order[pos]
        lw, dw = oracle.widths_c(lines["freq"][i], lines["air_broad"][i], lines["t_dep"][i],
                                 165.0, 0.3, S.CH4_MM)
        ref = oracle.line_shape(lines["freq"][i], lw, dw, g[cen[i]], L)
        assert rel_err(shapes[pos], ref) < TOL_XS
        gref = oracle.gcoeffs_c(lines["freq"][i], lines["a_coeff"][i], lines["e_lower"][i],
                                lines["g_up"][i], lines["g_lo"][i], lines["e_vib_up"][i],
                                lines["e_vib_lo"][i], 165.0)
        assert rel_err(gco[pos], gref) < 1e-12
    # normalisation: integral(shape) = 1 (MakeShape docstring, spect_classes.py:1994-1995)
    assert abs(shapes[0].sum() * (g[1] - g[0]) - 1.0) < 1e-3


def test_linearity_and_line_sharding_property(sb):
    """Size-independent property used at full size: the cell spectrum is linear in the line list
    (sum over two halves == whole), which is what line-sharding + all_reduce relies on."""
    engine, lineshape, S = sb
    g, lines = _cell_case(S, 500, 2990.0, 3010.0, 12, seed=21)
    whole = engine.LineSet(lines, g, S.CH4_MM, 12).gcoeff_cells([[0.05, 160.0]])
    parts = []
    for sl in (slice(0, 250), slice(250, 500)):
        sub = {k: (v[sl] if isinstance(v, np.ndarray) and v.shape[:1] == (500,) else v)
               for k, v in lines.items()}
        parts.append(engine.LineSet(sub, g, S.CH4_MM, 12).gcoeff_cells([[0.05, 160.0]]))
    tot = parts[0] + parts[1]
    assert rel_err(tot.cpu().numpy(), whole.cpu().numpy()) < 1e-12


def test_cell_batches_are_invisible(sb, monkeypatch):
    """LUT cells are processed in batches that reuse the per-(line, cell) tables back to back on one
    stream (no host synchronisation in between): the batch size must not change a bit."""
    engine, lineshape, S = sb
    g, lines = _cell_case(S, 300, 2990.0, 3010.0, 6, seed=33)
    cells = [[10.0 ** (-4 + 0.5 * j), 140.0 + 3.0 * j] for j in range(7)]
    base = engine.LineSet(lines, g, S.CH4_MM, 6).gcoeff_cells(cells)
    base32 = engine.LineSet(lines, g, S.CH4_MM, 6).gcoeff_cells_f32(cells)
    monkeypatch.setenv("SR_K2_BATCH", "2")
    ls = engine.LineSet(lines, g, S.CH4_MM, 6)
    import torch
    assert torch.equal(ls.gcoeff_cells(cells), base)
    assert torch.equal(ls.gcoeff_cells_f32(cells), base32)
    assert np.array_equal(ls.gcoeff_cells_host(cells), base.cpu().numpy())


@pytest.mark.parametrize("n_lines", [2500, 7000])     # 7000: > 1024 candidates per tile, long-run variant
def test_far_field_of_the_far_wings_is_invisible(monkeypatch, n_lines):
    """k_far_nodes (distant full far wings evaluated at 12 Chebyshev nodes per tile and interpolated)
    against the point-by-point evaluation (SR_K1_FAR=0): <= 1e-9 of the largest value of a row, over
    Titan-like and high pressures (ry from 1e-3 to ~50) and both storage types."""
    import torch
    from spectrobot_b200 import engine, synthetic as S
    g = S.spectral_grid(2990.0, 3010.0)                       # 40 001 points: lines up to 3 windows away
    lines = S.line_table(n_lines, 2986.0, 3014.0, n_levels=6, seed=21)
    ls = engine.LineSet(lines, g, S.CH4_MM, 6)
    cells = [[1e-4, 120.0], [0.05, 160.0], [2.5, 175.0], [150.0, 200.0], [1500.0, 94.0]]
    monkeypatch.setenv("SR_K1_FAR", "0")
    exact = ls.gcoeff_cells(cells)
    exact32 = ls.gcoeff_cells_f32(cells)
    monkeypatch.setenv("SR_K1_FAR", "2")          # forced: by default small launches stay point by point
    far = ls.gcoeff_cells(cells)
    far32 = ls.gcoeff_cells_f32(cells)
    assert not torch.equal(far, exact)                        # the path is really taken
    scale = exact.abs().amax(dim=3, keepdim=True).clamp_min(1e-300)
    assert float(((far - exact).abs() / scale).max()) < 1e-9
    # point-wise, wherever a row is not dominated by cancellation-free tiny far wings
    rel = ((far - exact).abs() / exact.abs().clamp_min(1e-300))[exact.abs() > 1e-12 * scale]
    assert float(rel.max()) < 1e-7
    assert float(((far32.double() - exact32.double()).abs() / scale).max()) < 2e-7
    # a slab of the grid with its own lineset is still bit-identical to the full build
    from spectrobot_b200 import parallel
    tp = ls.tile_points()
    p0, n = parallel.shard_slab(len(g), 1, 3, align=tp)
    ls_r = engine.LineSet(parallel.slab_lines(lines, g, p0, n, align=tp), g, S.CH4_MM, 6)
    assert torch.equal(ls_r.gcoeff_cells_window(cells, p0, n, f32=False), far[..., p0:p0 + n])
