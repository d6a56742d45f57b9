#!/usr/bin/env python
"""Tuning sweep over the kernel configuration switches (SR_K1_CFG, SR_K3_CFG)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spectrobot_b200 import engine, synthetic as S
from tools.prof_run import timed

def k1():
    w0, w1, n_lev = 2825.0, 3225.0, 12
    g = S.spectral_grid(w0, w1)
    lines = S.line_table(30000, w0, w1, n_levels=n_lev)
    cells = [[1e-4 * (1 + j), 150.0 + j] for j in range(8)]
    ref = None
    for cfg in os.environ.get("SR_TUNE_CFGS", "0,5,6,7").split(","):
        os.environ["SR_K1_CFG"] = cfg       # read at LineSet creation (tile geometry)
        ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
        out = torch.empty((8, n_lev, 3, len(g)), dtype=torch.float64, device="cuda")
        try:
            ms = timed(lambda: ls.gcoeff_cells(cells, out=out, check_status=False), 3)
        except Exception as e:
            print("k1 cfg", cfg, "failed", e); continue
        if ref is None: ref = out.clone()
        d = float(((out - ref).abs() / ref.abs().clamp_min(1e-300)).max())
        print("k1 cfg %s: %.3f ms / 8 cells -> %.3e evals/s (max rel diff %.1e)" % (cfg, min(ms), 8 * ls.n_active * 13010 / (min(ms) * 1e-3), d))
        del ls, out
    lines1 = S.line_table(30000, w0, w1, n_levels=1)
    for cfg in ("0", "1"):
        os.environ["SR_K1_CFG"] = cfg
        ls1 = engine.LineSet(lines1, g, 27.99, 1)
        out1 = torch.empty((4, 1, 3, len(g)), dtype=torch.float64, device="cuda")
        ms = timed(lambda: ls1.gcoeff_cells(cells[:4], out=out1, check_status=False), 3)
        print("k1 LTE cfg %s: %.3f ms / 4 cells -> %.3e evals/s" % (cfg, min(ms), 4 * ls1.n_active * 13010 / (min(ms) * 1e-3)))
    os.environ.pop("SR_K1_CFG", None)

def k3(n_los=8):
    w0, w1 = 2850.0, 3450.0
    g = S.spectral_grid(w0, w1); n_lev = 12
    lines = S.line_table(3000, w0, w1, n_levels=n_lev)
    atm = S.titan_atmosphere()
    st = S.limb_los_steps(np.linspace(360.0, 1040.0, n_los), [3] * n_los, [50.0] * n_los, atm, lines["level_energies"])
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1, st["temp"].min(), st["temp"].max())
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    steps = engine.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    sp = float(st["n_steps"].sum()) * len(g)
    tau, src = engine.los_tau_src([lut], steps)
    nst = torch.tensor(st["n_steps"], dtype=torch.int32, device="cuda")
    rad = torch.empty((n_los, len(g)), dtype=torch.float64, device="cuda")
    for cfg in ("0", "1", "2", "3", "4", "5"):
        os.environ["SR_K3_CFG"] = cfg
        ms = timed(lambda: engine.los_rt_layers(tau, src, nst, out=rad), 5)
        print("k3 cfg %s: %.3f ms -> %.1f GB/s" % (cfg, min(ms), (16 * sp + 8 * n_los * len(g)) / (min(ms) * 1e-3) / 1e9))

def fused(n_los=int(os.environ.get("TUNE_NLOS", "12"))):
    w0, w1 = 2850.0, 3450.0
    g = S.spectral_grid(w0, w1); n_lev = 12
    lines = S.line_table(3000, w0, w1, n_levels=n_lev)
    atm = S.titan_atmosphere()
    tg = np.repeat(np.linspace(400.0, 1000.0, n_los // 3), 3) + np.tile([-12.0, 0.0, 12.0], n_los // 3)
    st = S.limb_los_steps(tg, [3] * n_los, [50.0] * n_los, atm, lines["level_energies"])
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1, st["temp"].min(), st["temp"].max())
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    steps = engine.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    sp = float(st["n_steps"].sum()) * len(g)
    rad = torch.empty((n_los, len(g)), dtype=torch.float64, device="cuda")
    ref = None
    for G in ("1", "2", "3", "4"):
        for ppt in ("1", "2", "4"):
            os.environ["SR_LOS_G"] = G; os.environ["SR_LOS_PPT"] = ppt
            ms = timed(lambda: engine.los_rt_lut([lut], steps, out=rad, check_status=False), 3)
            if ref is None: ref = rad.clone()
            print("fused G %s ppt %s: %.3f ms -> %.3e step-points/s (same %s)" % (G, ppt, min(ms), sp / (min(ms) * 1e-3), bool((rad == ref).all())))
    import time
    host = torch.empty((n_los, len(g)), dtype=torch.float64).pin_memory().numpy()
    for _ in range(3):
        t0 = time.perf_counter(); engine.los_rt_lut_host([lut], steps, out=host); t1 = time.perf_counter()
        print("host call %d LOS: %.1f ms" % (n_los, 1e3 * (t1 - t0)))
    pag = np.empty((n_los, len(g)))
    t0 = time.perf_counter(); engine.los_rt_lut_host([lut], steps, out=pag); print("host call pageable out: %.1f ms" % (1e3 * (time.perf_counter() - t0)))


if __name__ == "__main__":
    what = sys.argv[1:] or ["k1", "k3", "fused"]
    if "k1" in what: k1()
    if "k3" in what: k3()
    if "fused" in what: fused()
