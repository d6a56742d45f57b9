"""Distribution of the layer optical depths of the bench workload (which share of the layer updates
could take a short series instead of the full exponential)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from spectrobot_b200 import engine, parallel


class A: small = False; pixels = 200; lines = 30000
P = bench.make_problem(A)
S, grid, lines, cells, atm = P["S"], P["grid"], P["lines"], P["cells"], P["atm"]
ls = engine.LineSet(lines, grid, S.CH4_MM, 12)
g32 = ls.gcoeff_cells_f32(cells)
lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
org, dirs, sun = bench.batch_geometry(S, 16, S.SEED + 7)
tv3 = S.vib_temperatures_3d(atm["z"], atm["temp"], lines["level_energies"])
At = engine.Atmosphere(atm["z"], atm["temp"], atm["pres"], np.full((1,) + atm["temp"].shape, 0.015),
                       tvib=tv3[None], lat_edges=atm["lat_edges"], radius_km=S.R_TITAN_KM, top_km=1500.0,
                       sza_nodes=S.SZA_NODES)
st, _ = engine.los_steps_build(At, org, dirs, sun=sun)
tau, src = engine.los_tau_src([lut], st)
nst = torch.as_tensor(st.n_steps, device="cuda")
mask = (torch.arange(tau.shape[1], device="cuda")[None, :] < nst[:, None])[:, :, None].expand_as(tau)
t = tau[mask].abs()
n = t.numel()
for thr in (1e-12, 1e-10, 1e-9, 7.5e-9, 1e-7, 1e-6, 4e-6, 1e-5, 1e-4, 1e-3, 1e-2, 0.1, 0.3466, 1.0):
    print("|tau| < %-8g : %6.2f %% of (step, point) pairs" % (thr, 100.0 * float((t < thr).sum()) / n))
# warp granularity of k_los_layers_f32: 64 consecutive points of one (LOS, step)
tt = tau[:, :, : (tau.shape[2] // 64) * 64].abs().reshape(tau.shape[0], tau.shape[1], -1, 64).amax(dim=3)
m2 = (torch.arange(tau.shape[1], device="cuda")[None, :] < nst[:, None])[:, :, None].expand_as(tt)
w = tt[m2]
for thr in (7.5e-9, 1e-7, 4e-6, 1e-4, 1e-3, 1e-2, 0.3466):
    print("warp max |tau| < %-8g : %6.2f %% of warps" % (thr, 100.0 * float((w < thr).sum()) / w.numel()))
