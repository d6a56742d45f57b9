"""One rank's share of the million-line K1 leg at 8 GPUs, on one GPU: which tile geometry is best
when a launch is only 293 tiles (half a wave)?  usage: SR_K1_CFG=<n> python tools/k1_slab_probe.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from spectrobot_b200 import engine, parallel, synthetic as S
g = S.spectral_grid(2850.0, 3450.0)
big = S.line_table(1000000, g[0], g[-1], n_levels=12, seed=20067)
probe = engine.LineSet(parallel.subset_lines(big, 0, 1), g, S.CH4_MM, 12)
al = probe.tile_points(); probe.close()
p0, n = parallel.shard_slab(len(g), 3, 8, align=512)
ls = engine.LineSet(parallel.slab_lines(big, g, p0, n, align=512), g, S.CH4_MM, 12)
out = torch.empty((1, 12, 3, n), dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ms = []
for i in range(6):
    torch.cuda.synchronize(); e0.record()
    ls.gcoeff_cells_window([[0.02, 155.0]], p0, n, f32=False, out=out)
    e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
print("cfg %s tile %d window %d pts, %d lines: %s ms  checksum %.10e" % (os.environ.get("SR_K1_CFG", "default"), ls.tile_points(), n, ls.n_active, ["%.3f" % m for m in ms[1:]], float(out.sum())))
