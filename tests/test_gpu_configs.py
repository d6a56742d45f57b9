"""GPU parity cases shaped like the other BASELINE.json configs (the bench measures configs[1]):
configs[0] CO 1-0 band + one nadir LOS through the Titan profile (radtran_test_CO.py),
configs[3] HCN LUT + limb scan, tangent heights x SZA 30-80 deg, 3-D vs 2-D atmosphere
           (run_0607_lut_HCN.py / radtran_3Dvs2D_radtrans_new.py),
configs[4] a million-line list sharded by line (inversion_sequences_20067.py shape)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL_XS, TOL_RAD = 1e-6, 1e-5


def _ray_to(alt_km, lat_deg, obs):
    la = np.radians(lat_deg)
    tg = (2575.0 + alt_km) * np.array([np.cos(la), 0.0, np.sin(la)])
    d = tg - obs
    return d / np.linalg.norm(d)


def test_config0_co_band_nadir_los(oracle):
    """CO 1-0 band, grid [2050,2250] (400 001 points), LTE isotopologue without level assignment
    (single set 'all', pop = 1/Q, smm:742-748, 2214-2218), one NADIR line of sight: the ray ends at
    the surface, the first (far) step sits at ~1.4 bar where the Lorentz width dominates."""
    from spectrobot_b200 import engine, synthetic as S
    g = S.spectral_grid(2050.0, 2250.0)
    assert len(g) == 400001
    lines = S.line_table(1500, 2050.0, 2250.0, n_levels=1, seed=2, q296=107.42, iso_ratio=0.986544)
    atm = S.titan_atmosphere(n_bands=1)
    A = engine.Atmosphere(atm["z"], atm["temp"], atm["pres"], np.full((1, 1, len(atm["z"])), 4.5e-5))
    obs = np.array([1.0e5, 0.0, 0.0])
    steps, _ = engine.los_steps_build(A, obs[None], -obs[None] / 1.0e5)
    ref_st = oracle.los_steps_build(atm["z"], atm["temp"], atm["pres"], np.full((1, 1, len(atm["z"])), 4.5e-5),
                                    obs[None], -obs[None] / 1.0e5)[0]
    n = int(steps.n_steps[0])
    assert n == ref_st["n_steps"] and n > 10 and steps.pres[0, 0] > 1000.0
    assert np.allclose(steps.column[0, 0, :n], ref_st["column"][0], rtol=1e-10)
    cells = S.rect_cells(steps.pres[0, :n].min() * 0.9, steps.pres[0, :n].max() * 1.1,
                         steps.temp[0, :n].min(), steps.temp[0, :n].max())
    ls = engine.LineSet(lines, g, 27.994915, 1)
    # cross sections of the deepest cell against the oracle (ry ~ 50: every window is region 1/2)
    deep = max(cells)
    got = ls.gcoeff_cells([deep])[0].cpu().numpy()
    ref = oracle.gcoeff_cell(lines, g, deep[1], deep[0], 27.994915, 1, n_threads=8)
    for ct in range(3):
        assert rel_err(got[0, ct], ref[0, ct]) < TOL_XS, ct
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 5, 1, 0.986544, level_energies=None)
    rad = engine.los_rt_lut([lut], steps).cpu().numpy()
    olut = dict(g32=g32.cpu().numpy(), pt=np.array(cells), level_energy=None, mol=5, iso=1,
                iso_ratio=0.986544, lte_unidentified=True)
    rref = oracle.los_rt([olut], steps.n_steps, steps.temp, steps.pres, steps.column, None, n_threads=8)
    assert rref.max() > 0 and rel_err(rad, rref) < TOL_RAD
    # surface emission as initial intensity, absorption only
    import torch
    i0 = np.full((1, len(g)), 2.0e-7)
    got_a = engine.los_rt_lut([lut], steps, i0=torch.tensor(i0, device="cuda"), solo_absorption=True)
    ref_a = oracle.los_rt([olut], steps.n_steps, steps.temp, steps.pres, steps.column, None, i0=i0,
                          solo_absorption=True, n_threads=8)
    assert rel_err(got_a.cpu().numpy(), ref_a) < TOL_RAD


def test_config3_hcn_limb_scan_3d_vs_2d(oracle):
    """HCN 3 um band, grid [3200,3400], 6 levels: limb scan of 15 tangent heights for SZA 30..80
    deg (SZA-dependent vibrational temperatures), once through the 7-band (3-D) atmosphere and once
    through its equatorial band alone (2-D); oracle parity on a window, and 3-D == 2-D for rays
    that never leave the equatorial band."""
    from spectrobot_b200 import engine, synthetic as S
    g = S.spectral_grid(3200.0, 3400.0)
    n_lev = 6
    lines = S.line_table(4000, 3200.0, 3400.0, n_levels=n_lev, seed=9, q296=892.2, iso_ratio=0.985114)
    energies = lines["level_energies"]
    atm = S.titan_atmosphere()
    z, nb = atm["z"], len(atm["temp"])
    heights = np.linspace(350.0, 1050.0, 15)
    obs = np.array([1.0e5, 0.0, 0.0])
    lat = 10.0                                               # inside the equatorial band (-30, 30)
    org = np.tile(obs, (15, 1))
    drc = np.array([_ray_to(h, lat, obs) for h in heights])
    vmr = np.full((1, nb, len(z)), 2.0e-6)
    all_steps = []
    for sza in (30.0, 55.0, 80.0):
        tv = np.stack([S.vib_temperatures(z, atm["temp"][b], energies, sza) for b in range(nb)], axis=1)
        A3 = engine.Atmosphere(z, atm["temp"], atm["pres"], vmr, tvib=tv[None], lat_edges=atm["lat_edges"])
        A2 = engine.Atmosphere(z, atm["temp"][3], atm["pres"][3], vmr[:, 3], tvib=tv[None, :, 3])
        s3, _ = engine.los_steps_build(A3, org, drc)
        s2, _ = engine.los_steps_build(A2, org, drc)
        assert np.array_equal(s3.n_steps, s2.n_steps)
        w = s2.n_steps_max
        assert np.array_equal(s3.temp[:, :w], s2.temp) and np.array_equal(s3.tvib[..., :w], s2.tvib)
        all_steps.append(s3)
    pres = np.concatenate([s.pres[s.pres > 1e-6] for s in all_steps])
    temp = np.concatenate([s.temp[s.pres > 1e-6] for s in all_steps])
    cells = S.rect_cells(pres.min() * 0.9, pres.max() * 1.1, temp.min(), temp.max())
    ls = engine.LineSet(lines, g, 27.010899, n_lev)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 23, 1, 0.985114, level_energies=energies)
    pt0, npt = 200000, 2048
    olut = dict(g32=np.ascontiguousarray(g32[..., pt0:pt0 + npt].cpu().numpy()), pt=np.array(cells),
                level_energy=energies, mol=23, iso=1, iso_ratio=0.985114, lte_unidentified=False)
    rads = []
    for s3 in all_steps:
        rad = engine.los_rt_lut([lut], s3)
        ref = oracle.los_rt([olut], s3.n_steps, s3.temp, s3.pres, s3.column, s3.tvib, n_threads=8)
        assert rel_err(rad[:, pt0:pt0 + npt].cpu().numpy(), ref) < TOL_RAD
        rads.append(rad)
    # non-LTE: the limb radiance above 600 km grows with the solar illumination (smaller SZA)
    hi = heights > 600.0
    assert float(rads[0][hi].sum()) > float(rads[2][hi].sum()) > 0.0


def test_config4_million_lines_line_sharded():
    """1e6 synthetic lines on the [2850,3450] grid, one (P,T) cell: the cell spectrum of the whole
    list equals the sum over four line shards (what the 8-GPU run all-reduces over NVLink), every
    row is finite and non-negative, and lines the level filter drops contribute nothing."""
    import torch
    from spectrobot_b200 import engine, synthetic as S
    n = 1000000
    g = S.spectral_grid(2850.0, 3450.0)
    lines = S.line_table(n, 2850.0, 3450.0, n_levels=12, seed=20067, frac_unlinked=0.02)
    cell = [[0.02, 155.0]]
    ls = engine.LineSet(lines, g, S.CH4_MM, 12)
    assert 0.97 * n < ls.n_active < 0.99 * n
    whole = ls.gcoeff_cells(cell)
    del ls
    acc = torch.zeros_like(whole)
    for q in range(4):
        sl = slice(q * n // 4, (q + 1) * n // 4)
        sub = {k: (v[sl] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v)
               for k, v in lines.items()}
        part = engine.LineSet(sub, g, S.CH4_MM, 12)
        acc += part.gcoeff_cells(cell)
        del part
    scale = whole.abs().amax(dim=3, keepdim=True).clamp_min(1e-300)
    assert float(((acc - whole).abs() / scale).max()) < 1e-11
    assert bool(torch.isfinite(whole).all()) and bool((whole >= 0).all())
