#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on N >= 2 GPUs:
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_check.py
Checks that (1) the cell-sharded LUT build + gather, (2) the LOS-sharded radiance batch + gather and
(3) the line-sharded cross-sections + NCCL all_reduce reproduce the single-GPU results."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from spectrobot_b200 import engine, parallel, synthetic as S  # noqa: E402


def main():
    rank, n, local = parallel.init("nccl")
    g = S.spectral_grid(2990.0, 3010.0)
    n_lev = 6
    lines = S.line_table(3000, 2987.0, 3013.0, n_levels=n_lev, seed=5)
    atm = S.titan_atmosphere()
    st = S.limb_los_steps(np.linspace(400.0, 1000.0, 10), [3] * 10, [50.0] * 10, atm,
                          lines["level_energies"])
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1,
                         st["temp"].min(), st["temp"].max())
    n_cells = len(cells)
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    full = ls.gcoeff_cells_f32(cells)                       # single-GPU reference on every rank
    # (1) cell-sharded build + gather
    mine = parallel.shard_cells(n_cells, rank, n)
    lut = torch.zeros_like(full)
    lut[mine] = ls.gcoeff_cells_f32([cells[c] for c in mine])
    parallel.gather_lut(lut, n_cells, rank, n)
    ok1 = bool(torch.equal(lut, full))
    # (2) LOS-sharded batch + gather
    L = engine.Lut(full, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    steps = engine.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    rad_full = engine.los_rt_lut([L], steps)
    b, e = parallel.shard_los(steps.n_los, rank, n)
    rad_mine = engine.los_rt_lut([L], steps.subset(slice(b, e)))
    rad = parallel.gather_rows(rad_mine, steps.n_los, rank, n)
    ok2 = bool(torch.equal(rad, rad_full))
    # (3) line-sharded cross-sections + all_reduce
    pts = [cells[0], cells[n_cells // 2]]
    xs_full = ls.gcoeff_cells(pts)
    xs = parallel.gcoeff_cells_line_sharded(lines, g, S.CH4_MM, n_lev, pts)
    err = float(((xs - xs_full).abs() / xs_full.abs().clamp_min(1e-300 + 1e-30 * xs_full.abs().max())).max())
    ok3 = err < 1e-12
    flags = torch.tensor([ok1, ok2, ok3], dtype=torch.int32, device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("multi_gpu_check world=%d: lut_gather=%s los_gather=%s line_allreduce=%s (rel err %.1e)"
              % (n, bool(flags[0]), bool(flags[1]), bool(flags[2]), err))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flags.min()) == 1 else 1)


if __name__ == "__main__":
    main()
