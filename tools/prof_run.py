#!/usr/bin/env python
"""Small driver for ncu captures of single kernels (see profiles/README.md for the commands).

    python tools/prof_run.py k1      # K1/K2 tile kernel: one CH4 LUT cell, 3e4 lines, 800 001 points
    python tools/prof_run.py k3      # K3 recursion over materialised layers
    python tools/prof_run.py fused   # K3a+K3 fused from the LUT
Prints CUDA-event timings so the same command is meaningful without ncu.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from spectrobot_b200 import engine, synthetic as S  # noqa: E402


def timed(fn, n=3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out


def k1(n_lines=30000, n_lev=12, w0=2825.0, w1=3225.0):
    g = S.spectral_grid(w0, w1)
    lines = S.line_table(n_lines, w0, w1, n_levels=n_lev)
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    out = torch.empty((1, n_lev, 3, len(g)), dtype=torch.float64, device="cuda")
    for P, T in ((1e-4, 150.0), (0.05, 160.0), (2.5, 175.0)):
        ms = timed(lambda: ls.gcoeff_cells([[P, T]], out=out, check_status=False))
        print("k1 P=%g T=%g: %s ms -> %.3e evals/s" % (P, T, ["%.3f" % m for m in ms],
                                                       ls.n_active * 13010 / (min(ms) * 1e-3)))


def k1b(n_c=8):
    """the bench's K1 measurement: 8 cells per launch on the [2850,3450] grid"""
    w0, w1, n_lev = 2850.0, 3450.0, 12
    g = S.spectral_grid(w0, w1)
    lines = S.line_table(30000, w0, w1, n_levels=n_lev)
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    out = torch.empty((n_c, n_lev, 3, len(g)), dtype=torch.float64, device="cuda")
    pts = [[0.05 * (3 + j), 150.0 + 2.0 * j] for j in range(n_c)]
    ms = timed(lambda: ls.gcoeff_cells(pts, out=out, check_status=False), 5)
    print("k1b %d cells: %s ms -> %.3e evals/s" % (n_c, ["%.3f" % m for m in ms],
                                                   n_c * ls.n_active * 13010 / (min(ms) * 1e-3)))


def los(mode, n_los=int(os.environ.get("SR_PROF_NLOS", "8"))):
    w0, w1 = 2850.0, 3450.0
    g = S.spectral_grid(w0, w1)
    n_lev = 12
    lines = S.line_table(30000, w0, w1, n_levels=n_lev)
    atm = S.titan_atmosphere()
    tg = np.linspace(360.0, 1040.0, n_los)
    st = S.limb_los_steps(tg, [3] * n_los, [50.0] * n_los, atm, lines["level_energies"])
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1,
                         st["temp"].min(), st["temp"].max())
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    steps = engine.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    sp = float(st["n_steps"].sum()) * len(g)
    if mode == "k3":
        tau, src = engine.los_tau_src([lut], steps)
        nst = torch.tensor(st["n_steps"], dtype=torch.int32, device="cuda")
        rad = torch.empty((n_los, len(g)), dtype=torch.float64, device="cuda")
        ms = timed(lambda: engine.los_rt_layers(tau, src, nst, out=rad), 5)
        print("k3: %s ms -> %.1f GB/s" % (["%.3f" % m for m in ms],
                                          (16 * sp + 8 * n_los * len(g)) / (min(ms) * 1e-3) / 1e9))
    else:
        rad = torch.empty((n_los, len(g)), dtype=torch.float64, device="cuda")
        ms = timed(lambda: engine.los_rt_lut([lut], steps, out=rad, check_status=False), 3)
        print("fused: %s ms -> %.3e step-points/s" % (["%.3f" % m for m in ms], sp / (min(ms) * 1e-3)))


def batch(n_pix=int(os.environ.get("SR_PROF_NPIX", "1500")), jac=False):
    """the bench's batch leg on fewer pixels: geometry -> device step builder -> low-res radiances
    (and, with jac, 12 derivative spectra per LOS)"""
    import time
    w0, w1, n_lev = 2850.0, 3450.0, 12
    g = S.spectral_grid(w0, w1)
    lines = S.line_table(30000, w0, w1, n_levels=n_lev)
    atm = S.titan_atmosphere()
    env = S.limb_los_steps([338.0, 1062.0] * 7, list(range(7)) * 2, [55.0] * 14, atm, lines["level_energies"])
    cells = S.rect_cells(env["pres"][env["pres"] > 1e-6].min() * 0.9, env["pres"].max() * 1.1,
                         env["temp"].min(), env["temp"].max())
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    rng = np.random.default_rng(7)
    n_b = 3 * n_pix
    tg_alt = np.repeat(rng.uniform(350.0, 1050.0, n_pix), 3) + np.tile([-12.0, 0.0, 12.0], n_pix)
    tg_lat = np.radians(np.repeat(rng.uniform(-90.0, 90.0, n_pix), 3))
    rt = S.R_TITAN_KM + tg_alt
    tgp = rt[:, None] * np.stack([np.cos(tg_lat), np.zeros(n_b), np.sin(tg_lat)], axis=1)
    east = np.tile([0.0, 1.0, 0.0], (n_b, 1))
    org = tgp + east * np.sqrt(1.0e5 ** 2 - rt ** 2)[:, None]
    tv = np.stack([S.vib_temperatures(atm["z"], atm["temp"][b], lines["level_energies"], 60.0)
                   for b in range(len(atm["temp"]))], axis=1)
    A = engine.Atmosphere(atm["z"], atm["temp"], atm["pres"], np.full((1,) + atm["temp"].shape, 0.015),
                          tvib=tv[None], lat_edges=atm["lat_edges"])
    masks = np.stack([np.clip(1 - np.abs(atm["z"] - c) / 100.0, 0, 1) for c in np.arange(300., 1401., 100.)])
    gdev = torch.as_tensor(g, device="cuda")
    c = torch.as_tensor(np.linspace(g[0] + 10.0, g[-1] - 10.0, 36), device="cuda")
    w = torch.as_tensor(np.full(36, 6.2), device="cuda")
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        steps, dfrac = engine.los_steps_build(A, org, -east, masks=masks if jac else None, jac_gas=0)
        t1 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if jac:
            low, jl = engine.los_rt_lut_jac_lowres([lut], steps, dfrac, gdev, c, w)
        else:
            low = engine.los_rt_lut_lowres([lut], steps, gdev, c, w)
        e1.record()
        t15 = time.perf_counter()
        low.cpu()
        t2 = time.perf_counter()
        print("batch%s %d LOS (%.1f steps/LOS, n_steps_max %d): steps %.3f s, radiances %.3f s (host returns "
              "after %.3f s, device %.3f s) -> %.1f LOS/s"
              % (" +jac" if jac else "", n_b, steps.n_steps.mean(), steps.n_steps_max, t1 - t0, t2 - t1,
                 t15 - t1, e0.elapsed_time(e1) * 1e-3, n_b / (t2 - t0)))


def conv(n_spec=512):
    g = S.spectral_grid(2850.0, 3450.0)
    gdev = torch.as_tensor(g, device="cuda")
    torch.manual_seed(1)
    spec = torch.rand((n_spec, len(g)), dtype=torch.float64, device="cuda")
    c = torch.as_tensor(np.linspace(g[0] + 10.0, g[-1] - 10.0, 36), device="cuda")
    w = torch.as_tensor(np.full(36, 6.2), device="cuda")
    out = engine.convolve_lowres(gdev, spec, c, w)
    ms = timed(lambda: engine.convolve_lowres(gdev, spec, c, w, out=out), 5)
    print("conv %d spectra: %s ms -> %.2f TB/s (checksum %.12e)"
          % (n_spec, ["%.3f" % m for m in ms], n_spec * len(g) * 8 / (min(ms) * 1e-3) / 1e12, float(out.sum())))


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "k1"
    if mode == "k1":
        k1()
    elif mode == "k1b":
        k1b(int(sys.argv[2]) if len(sys.argv) > 2 else 8)
    elif mode == "conv":
        conv()
    elif mode == "batch":
        batch()
    elif mode == "batchjac":
        batch(jac=True)
    else:
        los(mode)
