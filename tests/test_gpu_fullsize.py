"""GPU: BASELINE.json's full sizes (CH4 3.3 um: grid [2850,3450] cm-1 = 1 200 001 points, 3e4
lines, 12 levels) checked through size-independent properties plus oracle parity on samples the
oracle finishes in seconds (the reference ships no golden vectors)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL_XS, TOL_RAD = 1e-6, 1e-5
W0, W1, N_LEV, N_LINES = 2850.0, 3450.0, 12, 30000


@pytest.fixture(scope="module")
def full():
    import torch
    from spectrobot_b200 import engine, synthetic as S
    g = S.spectral_grid(W0, W1)
    assert len(g) == 1200001
    lines = S.line_table(N_LINES, W0, W1, n_levels=N_LEV)
    return dict(torch=torch, engine=engine, S=S, grid=g, lines=lines)


def _subset(lines, sl, n=N_LINES):
    return {k: (v[sl] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v)
            for k, v in lines.items()}


def test_cross_sections_full_size(full, oracle):
    """(i) parity with the oracle for a 240-line sample on the full grid; (ii) linearity in the
    line list at full size (two halves add up to the whole cell), which carries (i) to 3e4 lines;
    (iii) the float32 LUT store is the float64 result rounded like numpy astype."""
    eng, S, g, lines = full["engine"], full["S"], full["grid"], full["lines"]
    cell = [[0.05, 160.0]]
    idx = np.arange(0, N_LINES, 125)
    sample = _subset(lines, idx)
    ref = oracle.gcoeff_cell(sample, g, 160.0, 0.05, S.CH4_MM, N_LEV, n_threads=8)
    got = eng.LineSet(sample, g, S.CH4_MM, N_LEV).gcoeff_cells(cell)[0].cpu().numpy()
    for s in range(N_LEV):
        for ct in range(3):
            assert rel_err(got[s, ct], ref[s, ct]) < TOL_XS, (s, ct)
    del ref, got
    ls = eng.LineSet(lines, g, S.CH4_MM, N_LEV)
    whole = ls.gcoeff_cells(cell)
    half = N_LINES // 2
    a = eng.LineSet(_subset(lines, slice(0, half)), g, S.CH4_MM, N_LEV).gcoeff_cells(cell)
    b = eng.LineSet(_subset(lines, slice(half, N_LINES)), g, S.CH4_MM, N_LEV).gcoeff_cells(cell)
    a += b
    scale = whole.abs().amax(dim=3, keepdim=True).clamp_min(1e-300)
    assert float(((a - whole).abs() / scale).max()) < 1e-12
    assert bool((whole >= 0).all()) and bool(full["torch"].isfinite(whole).all())
    f32 = ls.gcoeff_cells_f32(cell)
    assert bool(full["torch"].equal(f32, whole.to(full["torch"].float32)))


def test_los_full_size(full, oracle, monkeypatch):
    """36 LOS on the full grid: oracle parity on three 4096-point windows (the oracle reads the
    same float32 LUT slice), invariance under LOS blocks / wavenumber chunks, and the Jacobian
    identity sum_p dI/dp = dI/d(ln column) when the parameter weights of every step add up to 1."""
    eng, S, g, lines, torch = full["engine"], full["S"], full["grid"], full["lines"], full["torch"]
    atm = S.titan_atmosphere()
    n_los = 36
    rng = np.random.default_rng(3)
    st = S.limb_los_steps(rng.uniform(350.0, 1050.0, n_los), rng.integers(0, 7, n_los),
                          rng.uniform(30.0, 80.0, n_los), atm, lines["level_energies"])
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1,
                         st["temp"].min(), st["temp"].max())
    ls = eng.LineSet(lines, g, S.CH4_MM, N_LEV)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = eng.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    steps = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    rad = eng.los_rt_lut([lut], steps)
    assert bool(torch.isfinite(rad).all()) and float(rad.max()) > 0
    for pt0 in (0, 599000, 1200001 - 4096):
        sl = slice(pt0, pt0 + 4096)
        olut = dict(g32=np.ascontiguousarray(g32[..., sl].cpu().numpy()), pt=np.array(cells),
                    level_energy=lines["level_energies"], mol=6, iso=1, iso_ratio=S.CH4_RATIO,
                    lte_unidentified=False)
        ref = oracle.los_rt([olut], st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"],
                            n_threads=8)
        assert rel_err(rad[:, sl].cpu().numpy(), ref) < TOL_RAD, pt0
    monkeypatch.setenv("SR_LOS_BLOCK", "7")
    monkeypatch.setenv("SR_LOS_CHUNK", "300000")
    assert bool(torch.equal(eng.los_rt_lut([lut], steps), rad))
    monkeypatch.delenv("SR_LOS_BLOCK")
    monkeypatch.delenv("SR_LOS_CHUNK")
    # Jacobian: 3 parameters whose weights add up to 1 on every step
    w = rng.uniform(0.0, 1.0, (n_los, steps.n_steps_max, 3))
    w /= w.sum(axis=2, keepdims=True)
    sub = steps.subset(slice(0, 6))
    r6, jac = eng.los_rt_lut_jac([lut], sub, w[:6])
    assert bool(torch.equal(r6, rad[:6]))
    h = 1e-5
    up = eng.LosSteps(sub.n_steps, sub.temp, sub.pres, sub.column * (1 + h), sub.tvib)
    dn = eng.LosSteps(sub.n_steps, sub.temp, sub.pres, sub.column * (1 - h), sub.tvib)
    fd = (eng.los_rt_lut([lut], up) - eng.los_rt_lut([lut], dn)) / (2 * h)
    tot = jac.sum(dim=1)
    scale = float(fd.abs().max())
    assert float((tot - fd).abs().max()) < 1e-6 * scale
