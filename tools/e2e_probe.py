"""Per-rank time of the host-buffer LOS call (sr_los_rt_lut_host) alone and with all ranks at once."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from spectrobot_b200 import engine, synthetic as S
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); engine.lib().sr_set_device(local)
if world > 1: dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w0, w1, n_lev = 2850.0, 3450.0, 12
g = S.spectral_grid(w0, w1)
lines = S.line_table(3000, w0, w1, n_levels=n_lev)
atm = S.titan_atmosphere()
n_los = 360
rng = np.random.default_rng(rank)
st = S.limb_los_steps(rng.uniform(350, 1050, n_los), rng.integers(0, 7, n_los), rng.uniform(30, 80, n_los), atm, lines["level_energies"])
env = S.limb_los_steps([338.0, 1062.0] * 7, list(range(7)) * 2, [55.0] * 14, atm, lines["level_energies"])
cells = S.rect_cells(env["pres"][env["pres"] > 1e-6].min() * 0.9, env["pres"].max() * 1.1, env["temp"].min(), env["temp"].max())
ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
lut = engine.Lut(ls.gcoeff_cells_f32(cells), cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
steps = engine.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
out = torch.empty((n_los, len(g)), dtype=torch.float64, pin_memory=True).numpy()
dev = torch.empty((n_los, len(g)), dtype=torch.float64, device="cuda")
def call_host(): engine.los_rt_lut_host([lut], steps, out=out)
def call_dev(): engine.los_rt_lut([lut], steps, out=dev); torch.cuda.synchronize()
def call_dev_copy():
    engine.los_rt_lut([lut], steps, out=dev); torch.from_numpy(out).copy_(dev); torch.cuda.synchronize()
for name, fn in (("host-call", call_host), ("device-only", call_dev), ("device+1copy", call_dev_copy)):
    fn()
    for mode in ("alone", "concurrent"):
        for r in range(world if mode == "alone" else 1):
            if world > 1: dist.barrier()
            if mode == "concurrent" or r == rank:
                ts = []
                for _ in range(3):
                    t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
                print("%-13s rank %d %-10s %.3f s/call -> %.0f LOS/s" % (name, rank, mode, np.median(ts), n_los / np.median(ts)), flush=True)
            if world > 1: dist.barrier()
if world > 1: dist.destroy_process_group()
