"""Wavenumber-slab partition of the hot path: a rank that owns the grid points
[pt0, pt0+n) builds every LUT cell on them from the lines that reach the slab and integrates every
LOS on them; the per-slab channel integrals add up to those of the whole grid."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case():
    import torch
    from spectrobot_b200 import engine, parallel, synthetic as S
    g = S.spectral_grid(2995.0, 3005.0)                       # 20 001 points
    n_lev = 5
    lines = S.line_table(1500, 2992.0, 3008.0, n_levels=n_lev, seed=11)
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    return dict(torch=torch, engine=engine, parallel=parallel, S=S, grid=g, lines=lines, ls=ls,
                n_lev=n_lev)


def test_window_build_is_bit_identical_to_the_full_build(case, monkeypatch):
    """Bit-identical when both builds take the same evaluation path (the far-field heuristics look at
    the launch size and at the line density of the lineset; this case sits right at
    the density threshold, so the path is pinned here); with the defaults the two builds agree to
    rounding."""
    eng, par, ls, g = case["engine"], case["parallel"], case["ls"], case["grid"]
    cells0 = [[0.02, 150.0], [0.4, 165.0], [2.5, 175.0]]
    p0_, n_ = par.shard_slab(len(g), 1, 2, align=ls.tile_points())
    ls_d = eng.LineSet(par.slab_lines(case["lines"], g, p0_, n_, align=ls.tile_points()), g,
                       case["S"].CH4_MM, case["n_lev"])
    wd, fd = ls_d.gcoeff_cells_window(cells0, p0_, n_, f32=False), ls.gcoeff_cells(cells0)[..., p0_:p0_ + n_]
    scale = fd.abs().amax(dim=3, keepdim=True).clamp_min(1e-300)
    assert float(((wd - fd).abs() / scale).max()) < 1e-12
    ls_d.close()
    monkeypatch.setenv("SR_K1_FAR", "0")
    cells = [[0.02, 150.0], [0.4, 165.0], [2.5, 175.0]]
    full64 = ls.gcoeff_cells(cells)
    full32 = ls.gcoeff_cells_f32(cells)
    tp = ls.tile_points()
    assert tp > 0
    for world in (2, 3):
        for rank in range(world):
            p0, n = par.shard_slab(len(g), rank, world, align=tp)
            # the rank's own lineset: only the lines whose window reaches the slab, whole grid
            sub = par.slab_lines(case["lines"], g, p0, n, align=tp)
            assert len(sub["freq"]) <= len(case["lines"]["freq"])
            ls_r = eng.LineSet(sub, g, case["S"].CH4_MM, case["n_lev"])
            w64 = ls_r.gcoeff_cells_window(cells, p0, n, f32=False)
            w32 = ls_r.gcoeff_cells_window(cells, p0, n, f32=True)
            assert case["torch"].equal(w64, full64[..., p0:p0 + n])
            assert case["torch"].equal(w32, full32[..., p0:p0 + n])
            ls_r.close()
    from spectrobot_b200._lib import SpectrobotError
    with pytest.raises(SpectrobotError):
        ls.gcoeff_cells_window(cells, 7, 100)                 # unaligned start
    with pytest.raises(SpectrobotError):
        ls.gcoeff_cells_window(cells, 0, len(g) + 1)          # beyond the grid


def test_slab_lowres_partial_sums_add_up(case):
    torch, eng, par, S, g = case["torch"], case["engine"], case["parallel"], case["S"], case["grid"]
    lines, n_lev, ls = case["lines"], case["n_lev"], case["ls"]
    atm = S.titan_atmosphere()
    st = S.limb_los_steps([420.0, 610.0, 800.0, 990.0], [3, 1, 5, 2], [35., 50., 65., 80.], atm,
                          lines["level_energies"])
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1,
                         st["temp"].min(), st["temp"].max())
    steps = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    centres = np.array([2996.0, 2998.7, 3000.05, 3002.5, 3004.4])
    widths = np.array([0.30, 0.45, 0.25, 0.50, 0.20])
    gdev = torch.as_tensor(g, device="cuda")
    lut = eng.Lut(ls.gcoeff_cells_f32(cells), cells, 6, 1, S.CH4_RATIO,
                  level_energies=lines["level_energies"])
    full = eng.los_rt_lut_lowres([lut], steps, gdev, centres, widths).cpu().numpy()
    assert np.all(full > 0)
    tp = ls.tile_points()
    for world in (2, 4):
        acc = np.zeros_like(full)
        for rank in range(world):
            p0, n = par.shard_slab(len(g), rank, world, align=tp)
            ls_r = eng.LineSet(par.slab_lines(lines, g, p0, n, align=tp), g, S.CH4_MM, n_lev)
            g32 = ls_r.gcoeff_cells_window(cells, p0, n)
            lut_r = eng.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
            part = eng.los_rt_lut_lowres([lut_r], steps, gdev[p0:p0 + n].contiguous(), centres, widths)
            acc += part.cpu().numpy()
            lut_r.close()
            ls_r.close()
        assert np.max(np.abs(acc - full) / full) < 1e-12
