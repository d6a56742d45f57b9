#!/usr/bin/env python
"""bench.py -- headline benchmark of the SpectRobot hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement on host cores

Workload (BASELINE.json configs[1]): CH4 3.3 um non-LTE limb radiances, spectral grid
[2850,3450] cm-1 at 5e-4 (1 200 001 points), 12 vibrational levels, 3e4 synthetic lines, float32
(P,T) LUT built by the K2 kernel, a block of synthetic VIMS limb lines of sight (tangent heights
350-1050 km, SZA 30-80 deg, 3 LOS per pixel).

A "step" is one pass of the LOS integral over the resident LOS block:
  value     LOS radiances/s of K3 (recursion over layer optical depths and source functions that are
            already resident in HBM), CUDA-event timed, max over ranks
  e2e       the same metric through the reference-facing host call (radtran_fast's contract: host
            step tables in, host hi-res radiances out; sr_los_rt_lut_host), copies inside
  roofline  K3 kernel against the measured HBM copy bandwidth (16 B per LOS*step*point + 8 B per
            LOS*point)
Extra objects report the other two parts of BASELINE.json's metric: "voigt" (K1 line*gridpoint
evals/s, FP64 roofline) and "lut_build" (K2 wall time of the whole LUT), and "fused" (K3a+K3 from
the LUT with device-resident inputs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_GRID_FULL = 1200001
W0, W1 = 2850.0, 3450.0
N_LEVELS = 12
N_LINES = 30000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--los-block", type=int, default=36, help="LOS per rank in the K3 block")
    ap.add_argument("--e2e-los", type=int, default=360, help="LOS per host call of the e2e leg")
    ap.add_argument("--fused-los", type=int, default=360,
                    help="LOS per rank of the device-resident K3a+K3 leg")
    ap.add_argument("--batch-pixels", type=int, default=10000,
                    help="pixels (3 LOS each) of the low-res batch leg, whole job; 0 disables it")
    ap.add_argument("--lines", type=int, default=N_LINES)
    ap.add_argument("--small", action="store_true", help="tiny sizes (CI / debugging only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def make_workload(args, rank, n_los, want_lines=True):
    from spectrobot_b200 import synthetic as S
    w0, w1 = (2990.0, 3010.0) if args.small else (W0, W1)
    grid = S.spectral_grid(w0, w1)
    n_lines = 600 if args.small else args.lines
    lines = S.line_table(n_lines if want_lines else 10, w0, w1, n_levels=N_LEVELS)
    atm = S.titan_atmosphere()
    rng = np.random.default_rng(S.SEED + 1000 * rank)
    n_pix = (n_los + 2) // 3
    tg = rng.uniform(350.0, 1050.0, n_pix)
    band = rng.integers(0, 7, n_pix)
    sza = rng.uniform(30.0, 80.0, n_pix)
    # 3 LOS per pixel (low, centre, up: spect_main_module.py:3091-3096), +-12 km about the centre
    tgs = np.repeat(tg, 3)[:n_los] + np.tile([-12.0, 0.0, 12.0], n_pix)[:n_los]
    st = S.limb_los_steps(tgs, np.repeat(band, 3)[:n_los], np.repeat(sza, 3)[:n_los], atm,
                          lines["level_energies"])
    # LUT cells must cover every rank's LOS block: use the envelope of the whole tangent range
    env = S.limb_los_steps([338.0, 1062.0] * 7, list(range(7)) * 2, [55.0] * 14, atm,
                           lines["level_energies"])
    pmin = min(st["pres"][st["pres"] > 1e-6].min(), env["pres"][env["pres"] > 1e-6].min())
    pmax = max(st["pres"].max(), env["pres"].max())
    tmin = min(st["temp"].min(), env["temp"].min())
    tmax = max(st["temp"].max(), env["temp"].max())
    cells = S.rect_cells(pmin * 0.9, pmax * 1.1, tmin, tmax)
    return dict(grid=grid, lines=lines, st=st, cells=cells, S=S)


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons sampled during the timed region: NVML in-process (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints;
    spawning nvidia-smi every 0.2 s from every rank perturbs the host-side legs), nvidia-smi as
    the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes the visible devices: map through CUDA_VISIBLE_DEVICES if set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        act = lambda bit: "Active" if r & bit else "Not Active"   # noqa: E731
        # order of Q: hw_slowdown 0x8, hw_thermal_slowdown 0x40, sw_thermal_slowdown 0x20, sw_power_cap 0x4
        return [str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        period = float(os.environ.get("SR_BENCH_SAMPLE_S", "0.5"))   # 10 Hz measurably slows the host legs
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.samples.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits"], capture_output=True,
                                         text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(period)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4)
                          if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_mhz_min": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# CPU arm (oracle): used as cpu_baseline of our arm and as --impl reference
# --------------------------------------------------------------------------------------------
def cpu_los_sample(args, wl, seconds, threads):
    """Times the CPU restatement of the LOS path (LUT interpolation + populations + layer
    recursion, oracle/sr_oracle.c orc_los_rt) on a bounded sample of the workload: a few LOS of
    the block on a contiguous range of grid points, LUT rows for that range built by the CPU
    restatement of the cell builder from the lines centred inside the range.
    Returns (LOS-equivalents per second, description)."""
    from oracle import cpu_oracle as O
    S = wl["S"]
    st, grid, lines = wl["st"], wl["grid"], wl["lines"]
    n_pts = 4096 if args.small else 16384
    pt0 = (len(grid) // 2 // 1024) * 1024
    sub = grid[pt0:pt0 + n_pts]
    inside = (lines["freq"] > sub[0]) & (lines["freq"] < sub[-1])
    sub_lines = {k: (v[inside] if isinstance(v, np.ndarray) and v.shape[:1] == inside.shape else v)
                 for k, v in lines.items()}
    n_los = min(2, st["temp"].shape[0])
    need = set()
    for l in range(n_los):
        for k in range(st["n_steps"][l]):
            c, _ = O.lut_weights(wl["cells"], st["pres"][l, k], st["temp"][l, k])
            need.update(int(x) for x in c if x >= 0)
    lin = O.line_window_offsets(grid)
    g32 = np.zeros((len(wl["cells"]), N_LEVELS, 3, n_pts), dtype=np.float32)
    for c in sorted(need):
        P, T = wl["cells"][c]
        g32[c] = O.gcoeff_cell(sub_lines, sub, T, P, S.CH4_MM, N_LEVELS, n_threads=threads,
                               lin_grid=lin).astype(np.float32)
    lut = dict(g32=g32, pt=np.array(wl["cells"]), level_energy=lines["level_energies"], mol=6,
               iso=1, iso_ratio=S.CH4_RATIO, lte_unidentified=False)
    sl = slice(0, n_los)
    kw = dict(n_steps=st["n_steps"][sl], temp=st["temp"][sl], pres=st["pres"][sl],
              column=st["column"][:, sl], tvib=st["tvib"][:, :, sl], n_threads=threads)
    O.los_rt([lut], **kw)                                   # warm-up
    reps, t_used = 0, 0.0
    while t_used < seconds and reps < 50:
        t0 = time.perf_counter()
        O.los_rt([lut], **kw)
        t_used += time.perf_counter() - t0
        reps += 1
    per = t_used / reps
    los_equiv = n_los * n_pts / float(len(grid))
    desc = ("%d LOS x %d of %d grid points, %d steps/LOS avg, %d threads, %d reps; "
            "CPU restatement (oracle) of LUT interpolation + level populations + layer recursion"
            % (n_los, n_pts, len(grid), int(st["n_steps"][sl].mean()), threads, reps))
    return los_equiv / per, per, desc


def cpu_voigt_sample(args, wl, threads, n_sample=None):
    """CPU restatement of one LUT cell (humliv_bb + G coefficients + line sum) on a line sample."""
    from oracle import cpu_oracle as O
    S = wl["S"]
    grid, lines = wl["grid"], wl["lines"]
    n = min(len(lines["freq"]), n_sample or (200 if args.small else 30000))   # whole list: the fixed
    # cost of the per-thread output spectra (n_threads x 346 MB) is amortised as in a real LUT cell
    sub_lines = {k: (v[:n] if isinstance(v, np.ndarray) and v.shape[:1] == lines["freq"].shape else v)
                 for k, v in lines.items()}
    t0 = time.perf_counter()
    O.gcoeff_cell(sub_lines, grid, 160.0, 0.05, S.CH4_MM, N_LEVELS, n_threads=threads)
    dt = time.perf_counter() - t0
    return n * 13010 / dt, "%d lines x 13010 points, one (P,T) cell, %d threads" % (n, threads)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    wl = make_workload(args, 0, max(2, min(args.los_block, 6)))
    vals, per = [], []
    for i in range(args.warmup + args.steps):
        v, p, desc = cpu_los_sample(args, wl, seconds=2.0, threads=threads)
        if i >= args.warmup:
            vals.append(v)
            per.append(p)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "LOS radiances/s", "value": value, "unit": "LOS/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(per)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, wl),
        "cpu_baseline": {"value": value, "unit": "LOS/s", "cores": threads, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "LOS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference (Fortran 77 + Python 2, spect_base_module missing) cannot be built "
                "or run here; this arm times the C restatement in oracle/ on all host threads",
    }
    print(json.dumps(line))


def workload_config(args, wl):
    return {"workload": "CH4 3.3um non-LTE limb LOS radiances (BASELINE configs[1], "
                        "radtran_3D_ch4.py shape)",
            "grid": "[%g,%g] cm-1 step 5e-4 (%d points)" % (wl["grid"][0], wl["grid"][-1],
                                                           len(wl["grid"])),
            "n_levels": N_LEVELS, "n_lines": int(len(wl["lines"]["freq"])),
            "lut_cells": len(wl["cells"]), "los_per_rank": int(wl["st"]["temp"].shape[0]),
            "e2e_los_per_call": int(args.e2e_los), "fused_los_per_rank": int(args.fused_los),
            "steps_per_los_mean": float(wl["st"]["n_steps"].mean()),
            "l2": "inputs larger than L2 (tau/S block streamed once per step)",
            "small": bool(args.small)}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    # only the JSON line may go to stdout (NCCL prints its version banner there)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from spectrobot_b200 import engine
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    engine.lib().sr_set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = engine.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_los = 6 if args.small else args.los_block
    n_fused = 12 if args.small else max(args.fused_los, n_los)
    n_e = 9 if args.small else max(1, args.e2e_los)
    n_big = max(n_los, n_fused, n_e)
    wl = make_workload(args, rank, n_big)     # the K3 block is the first n_los LOS of the batch
    S, grid, lines, st_all, cells = wl["S"], wl["grid"], wl["lines"], wl["st"], wl["cells"]
    st = {k: (v[..., :n_los, :] if k in ("temp", "pres", "column", "tvib") else v[:n_los])
          for k, v in st_all.items()}
    wl["st"] = st
    n_grid, n_cells = len(grid), len(cells)
    ls = engine.LineSet(lines, grid, S.CH4_MM, N_LEVELS)

    # ---- K1: evals/s per (P,T) cell; cells are launched in batches like the LUT builder does ----
    n_k1 = 2 if args.small else 8
    cell_buf = torch.empty((n_k1, N_LEVELS, 3, n_grid), dtype=torch.float64, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def k1_time(n_c, reps=5):
        ms = []
        for i in range(reps + 2):
            pts = [[0.05 * (1 + i + j), 150.0 + 2.0 * j] for j in range(n_c)]
            torch.cuda.synchronize()
            ev0.record()
            ls.gcoeff_cells(pts, out=cell_buf[:n_c], check_status=(i < 2))
            ev1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(ev0.elapsed_time(ev1))
        return float(np.median(ms)) * 1e-3

    k1_single = k1_time(1)
    k1_t = k1_time(n_k1) / n_k1
    evals = ls.n_active * 13010.0
    fp64_peak = engine.fp64_peak(40000)
    voigt = {"metric": "Voigt line*gridpoint evals/s", "value": evals / k1_t, "unit": "evals/s",
             "ms_per_cell": 1e3 * k1_t, "cells_per_launch": n_k1,
             "ms_single_cell_launch": 1e3 * k1_single, "lines": int(ls.n_active),
             "roofline": {"bound": "fp64", "achieved": 15.0 * evals / k1_t / 1e12,
                          "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                          "frac": 15.0 * evals / k1_t / fp64_peak, "traffic": None,
                          "note": "15 FP64 flop per eval (SURVEY 8d) over the DFMA rate measured "
                                  "live by sr_fp64_peak; time includes k_line_cell_params, "
                                  "k_core_eval and k_voigt_tile"}}
    del cell_buf

    # ---- K1 on a million-line list, sharded by line with an all_reduce of the partial spectra
    # (BASELINE configs[4]); the line table and the per-rank line set are set-up
    n_big = 20000 if args.small else 1000000
    big = S.line_table(n_big, grid[0], grid[-1], n_levels=N_LEVELS, seed=20067)
    from spectrobot_b200 import parallel as _par
    b0_, b1_ = _par.shard_lines(n_big, rank, world)
    ls_big = engine.LineSet(_par.subset_lines(big, b0_, b1_), grid, S.CH4_MM, N_LEVELS)
    xs_big = torch.empty((1, N_LEVELS, 3, n_grid), dtype=torch.float64, device="cuda")
    rows_big = _par.fed_rows(big, N_LEVELS)
    big_each = []
    for i in range(4):
        barrier()
        ev0.record()
        ls_big.gcoeff_cells([[0.02, 155.0]], out=xs_big, check_status=(i == 0))
        _par.allreduce_spectra(xs_big, rows=rows_big)
        ev1.record()
        barrier()
        if i > 0:
            big_each.append(ev0.elapsed_time(ev1))
    big_ms = max_over_ranks(float(np.median(big_each)))
    voigt_sharded = {"metric": "Voigt line*gridpoint evals/s, 1e6-line list sharded by line",
                     "value": n_big * 13010.0 / (big_ms * 1e-3), "unit": "evals/s", "lines": n_big,
                     "lines_per_rank": int(b1_ - b0_), "ms": big_ms, "scaling": "strong",
                     "allreduce_bytes": int(len(rows_big) * n_grid * 8) if world > 1 else 0,
                     "roofline_frac_fp64": 15.0 * n_big * 13010.0 / (big_ms * 1e-3) / (world * fp64_peak),
                     "path": "k_voigt_tile on the rank's lines, then NCCL all_reduce(SUM, fp64) of the "
                             "[12][3][n_grid] partial spectra"}
    ls_big.close()
    del ls_big, xs_big, big

    # ---- K2: LUT build (cells sharded over ranks, all_gather) --------------------------------
    from spectrobot_b200 import parallel
    my_cells = parallel.shard_cells(n_cells, rank, world)
    g32 = engine.lut_tensor(n_cells, N_LEVELS, n_grid)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    if my_cells:   # contiguous block per rank, written in place
        ls.gcoeff_cells_f32([cells[c] for c in my_cells],
                            out=g32[my_cells[0]:my_cells[0] + len(my_cells)])
    ev1.record()
    torch.cuda.synchronize()
    build_dev_s = ev0.elapsed_time(ev1) * 1e-3
    build_wall_s = max_over_ranks(time.perf_counter() - t0)
    # cells are independent (no collective on the build itself); the LOS leg below wants the whole
    # table on every rank, so the cell blocks are exchanged once over NVLink (NCCL broadcasts)
    barrier()
    t1 = time.perf_counter()
    parallel.gather_lut(g32, n_cells, rank, world)
    barrier()
    gather_s = max_over_ranks(time.perf_counter() - t1)
    lut_build = {"metric": "CH4 LUT build time", "value": build_wall_s, "unit": "s",
                 "cells": n_cells, "cells_per_rank": len(my_cells),
                 "device_s_per_rank": max_over_ranks(build_dev_s),
                 "evals_per_s": n_cells * evals / build_wall_s,
                 "includes": "per-cell parameters, core evaluation, tile kernel with float32 store",
                 "replicate_gather_s": gather_s if world > 1 else 0.0,
                 "lut_bytes": int(n_cells) * N_LEVELS * 3 * n_grid * 4}
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    steps_all = engine.LosSteps(st_all["n_steps"], st_all["temp"], st_all["pres"],
                                st_all["column"], st_all["tvib"])
    steps = steps_all.subset(slice(0, n_los))

    # ---- K3a: materialise tau/S of the block (resident inputs of K3) -------------------------
    tau, src = engine.los_tau_src([lut], steps)
    nst = torch.tensor(st["n_steps"], dtype=torch.int32, device="cuda")
    rad = torch.empty((n_los, n_grid), dtype=torch.float64, device="cuda")
    step_pts = float(st["n_steps"].sum()) * n_grid
    k3_bytes = 16.0 * step_pts + 8.0 * n_los * n_grid

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        engine.los_rt_layers(tau, src, nst, out=rad)
    barrier()
    l0 = lib.sr_kernel_launch_count()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        engine.los_rt_layers(tau, src, nst, out=rad)
    ev1.record()
    barrier()
    k3_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = lib.sr_kernel_launch_count() - l0
    value = world * n_los / (k3_ms * 1e-3)
    rad_k3 = rad.clone()
    del tau, src

    # ---- K3a+K3 from the LUT, device-resident (grouped tensor-path product + recursion) -------
    steps_f = steps_all.subset(slice(0, n_fused))
    rad_f = torch.empty((n_fused, n_grid), dtype=torch.float64, device="cuda")
    for _ in range(2):
        engine.los_rt_lut([lut], steps_f, out=rad_f)
    barrier()
    n_f = max(1, min(args.steps, 3))
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_f + 1)]
    evs[0].record()
    for i in range(n_f):
        engine.los_rt_lut([lut], steps_f, out=rad_f, check_status=False)
        evs[i + 1].record()
    barrier()
    fused_each = [evs[i].elapsed_time(evs[i + 1]) for i in range(n_f)]
    fused_ms = max_over_ranks(float(np.median(fused_each)))   # median: one call in ~10 catches a host hiccup
    agree = float(((rad_f[:n_los] - rad_k3).abs() / rad_k3.abs().clamp_min(1e-300)).max().item())
    fused_step_pts = float(st_all["n_steps"][:n_fused].sum()) * n_grid
    del rad_f
    # algorithmic FP64 work of the grouped product: every (step, point) needs one FMA per non-zero
    # LUT row of the 4 interpolation cells
    rows_cell = int((g32[:2] != 0).any(dim=3).any(dim=0).sum().item())
    fused_flop = 2.0 * 4 * rows_cell * fused_step_pts

    # ---- forward + analytic Jacobians (BASELINE configs[4] shape: every LOS also returns its
    # derivative spectra for n_par VMR nodes), reduced to the instrument channels on the device ----
    n_par = 12
    ns_f = st_all["n_steps"][:n_fused]
    kk = np.arange(steps_f.n_steps_max)[None, :]
    node = np.minimum(kk * n_par // np.maximum(ns_f[:, None], 1), n_par - 1)     # triangle masks:
    wgt = 0.25 + 0.5 * ((kk * 7919) % 97) / 97.0                                 # two nodes per step
    dfrac = np.zeros((n_fused, steps_f.n_steps_max, n_par))
    ii = np.arange(n_fused)[:, None].repeat(steps_f.n_steps_max, 1)
    np.add.at(dfrac, (ii, kk.repeat(n_fused, 0), node), np.broadcast_to(wgt, node.shape))
    np.add.at(dfrac, (ii, kk.repeat(n_fused, 0), np.minimum(node + 1, n_par - 1)),
              np.broadcast_to(1.0 - wgt, node.shape))
    dfrac *= (kk < ns_f[:, None])[:, :, None]
    j_centres = np.linspace(grid[0] + 10.0, grid[-1] - 10.0, 36)
    j_widths = np.full(36, 6.2)
    gdev_j = torch.as_tensor(grid, device="cuda")
    cj, wj = torch.as_tensor(j_centres, device="cuda"), torch.as_tensor(j_widths, device="cuda")
    engine.los_rt_lut_lowres([lut], steps_f, gdev_j, cj, wj)
    low_j, jac_j = engine.los_rt_lut_jac_lowres([lut], steps_f, dfrac, gdev_j, cj, wj)
    barrier()
    fwd_each, jac_each = [], []
    for _ in range(3):
        ev0.record()
        low_f = engine.los_rt_lut_lowres([lut], steps_f, gdev_j, cj, wj, check_status=False)
        ev1.record()
        torch.cuda.synchronize()
        fwd_each.append(ev0.elapsed_time(ev1))
        ev0.record()
        low_j, jac_j = engine.los_rt_lut_jac_lowres([lut], steps_f, dfrac, gdev_j, cj, wj,
                                                    check_status=False)
        ev1.record()
        torch.cuda.synchronize()
        jac_each.append(ev0.elapsed_time(ev1))
    barrier()
    fwd_low_ms = max_over_ranks(float(np.median(fwd_each)))
    jac_ms = max_over_ranks(float(np.median(jac_each)))
    jacobian = {"value": world * n_fused / (jac_ms * 1e-3), "unit": "LOS/s (radiance + %d derivative "
                "spectra each)" % n_par, "n_par": n_par, "los_per_rank": n_fused, "ms": jac_ms,
                "forward_only_ms": fwd_low_ms, "channels": 36, "ms_each_rank0": jac_each,
                "derivative_spectra_per_s": world * n_fused * n_par / (jac_ms * 1e-3),
                "radiance_same_as_forward": bool(torch.equal(low_j, low_f)),
                "finite": bool(torch.isfinite(jac_j).all().item()),
                "path": "sr_los_rt_lut_jac_lowres_dev (k_los_mma -> k_los_layers_jac -> "
                        "k_convolve_lowres per LOS block)"}
    del low_j, jac_j, low_f, dfrac

    # ---- e2e: reference-facing host call (host step tables in, host radiances out) ------------
    sub = steps_all.subset(slice(0, n_e))
    host_out = torch.empty((n_e, n_grid), dtype=torch.float64, pin_memory=True).numpy()
    engine.los_rt_lut_host([lut], sub, out=host_out)
    barrier()
    n_h = max(1, min(args.steps, 3))
    e2e_each = []
    for _ in range(n_h):
        t0 = time.perf_counter()
        engine.los_rt_lut_host([lut], sub, out=host_out)
        e2e_each.append(time.perf_counter() - t0)
    barrier()
    e2e_s = max_over_ranks(float(np.median(e2e_each)))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = int(sub.n_steps.nbytes + sub.temp.nbytes + sub.pres.nbytes + sub.column.nbytes +
              sub.tvib.nbytes)
    d2h = int(host_out.nbytes)

    # ---- batch: the north_star batch (10^4 pixels x 3 LOS), reduced to VIMS-like channels -------
    batch = None
    n_pix = 40 if args.small else args.batch_pixels
    if n_pix > 0:
        del host_out
        p0, p1 = parallel.shard_range(n_pix, rank, world)      # pixels are split over the ranks
        n_b = 3 * (p1 - p0)
        # every pixel has its own geometry: tangent height U(350,1050) km, tangent latitude
        # U(-90,90), observer at 1e5 km; low / centre / up LOS +-12 km.  Geometry ->
        # radtran steps for the whole batch on the device (sr_los_steps_build).
        rng_b = np.random.default_rng(S.SEED + 7 + 1000 * rank)
        npx = p1 - p0
        tg_alt = np.repeat(rng_b.uniform(350.0, 1050.0, npx), 3) + np.tile([-12.0, 0.0, 12.0], npx)
        tg_lat = np.radians(np.repeat(rng_b.uniform(-90.0, 90.0, npx), 3))
        rt_ = S.R_TITAN_KM + tg_alt
        tgp = rt_[:, None] * np.stack([np.cos(tg_lat), np.zeros(n_b), np.sin(tg_lat)], axis=1)
        east = np.tile([0.0, 1.0, 0.0], (n_b, 1))
        org_b = tgp + east * np.sqrt(1.0e5 ** 2 - rt_ ** 2)[:, None]
        atm_b = S.titan_atmosphere()
        tv_b = np.stack([S.vib_temperatures(atm_b["z"], atm_b["temp"][b_], lines["level_energies"], 60.0)
                         for b_ in range(len(atm_b["temp"]))], axis=1)          # [set][band][z]
        A_b = engine.Atmosphere(atm_b["z"], atm_b["temp"], atm_b["pres"],
                                np.full((1,) + atm_b["temp"].shape, 0.015), tvib=tv_b[None],
                                lat_edges=atm_b["lat_edges"], radius_km=S.R_TITAN_KM, top_km=1500.0)
        engine.los_steps_build(A_b, org_b, -east)       # warm-up: scratch pool growth, first-touch of host pages
        barrier()
        t0 = time.perf_counter()
        steps_b, _ = engine.los_steps_build(A_b, org_b, -east, delta_x=5.0, max_T_variation=5.0,
                                            max_Plog_variation=1.0)
        barrier()
        steps_build_s = max_over_ranks(time.perf_counter() - t0)
        centres = np.linspace(grid[0] + 10.0, grid[-1] - 10.0, 36)   # ~16 nm sampling at 3.3 um
        widths = np.full(36, 6.2)                                    # sigma of a 14.6 cm-1 FWHM
        gdev = torch.as_tensor(grid, device="cuda")
        cdev, wdev = torch.as_tensor(centres, device="cuda"), torch.as_tensor(widths, device="cuda")
        warm = steps_b.subset(slice(0, min(n_b, 256)))
        engine.los_rt_lut_lowres([lut], warm, gdev, cdev, wdev)      # workspaces, first-call costs
        barrier()
        t0 = time.perf_counter()
        low = engine.los_rt_lut_lowres([lut], steps_b, gdev, cdev, wdev)
        low_host = low.cpu().numpy()                                  # the call's result leaves the device
        barrier()
        batch_s = max_over_ranks(time.perf_counter() - t0)
        batch = {"pixels": n_pix, "los": 3 * n_pix, "seconds": batch_s,
                 "steps_build_seconds": steps_build_s,
                 "value": 3 * n_pix / (batch_s + steps_build_s), "unit": "LOS/s", "channels": 36,
                 "d2h_bytes": int(low_host.nbytes), "distinct_geometries_per_rank": int(n_b),
                 "steps_per_los_mean": float(steps_b.n_steps.mean()),
                 "finite": bool(np.isfinite(low_host).all()),
                 "path": "sr_los_steps_build (LOS geometry -> radtran steps, device) + "
                         "sr_los_rt_lut_lowres_dev: observer/direction vectors in, low-res channel "
                         "radiances out (wall clock incl. host planning and copies); hi-res radiances "
                         "exist per LOS block on the device only"}
        del low, steps_b

    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0
    traffic = None     # dram__bytes_read + write of this kernel on this workload, one ncu capture
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_k3_bench_traffic.json")))
        if not args.small and n_los == 36 and n_grid == 1200001:
            traffic = float(tr["traffic"])
    except Exception:
        pass
    achieved = k3_bytes / (k3_ms * 1e-3) / 1e9
    line = {
        "metric": "LOS radiances/s", "value": value, "unit": "LOS/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": k3_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, wl),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": traffic, "kernel": "k_los_layers",
                     "traffic_source": "profiles/r1_k3_bench_traffic.json (ncu --set full, per launch)"
                     if traffic else None,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks
                     else "fallback 6650 GB/s (of fallback)",
                     "algorithmic_bytes_per_launch": k3_bytes},
        "e2e": {"value": world * n_e / e2e_s, "unit": "LOS/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "los_per_call": n_e, "s_each_rank0": e2e_each,
                "timing": "median of the calls (wall clock around the synchronous host call)",
                "path": "sr_los_rt_lut_host: LUT interpolation + populations + layer recursion, "
                        "host step tables -> host hi-res radiances"},
        "fused": {"value": world * n_fused / (fused_ms * 1e-3), "unit": "LOS/s",
                  "los_per_rank": n_fused, "ms_per_step": fused_ms,
                  "step_points_per_s": world * fused_step_pts / (fused_ms * 1e-3),
                  "max_rel_diff_vs_k3": agree, "ms_each_rank0": fused_each,
                  "roofline": {"bound": "fp64", "achieved": fused_flop / (fused_ms * 1e-3) / 1e12,
                               "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                               "frac": fused_flop / (fused_ms * 1e-3) / fp64_peak, "traffic": None,
                               "kernel": "k_los_mma (+ k_los_layers in the same timed region)",
                               "note": "2 flop x %d non-zero LUT rows x 4 cells per (step, point); "
                                       "DFMA rate measured live by sr_fp64_peak" % rows_cell}},
        "jacobian": jacobian,
        "batch": batch, "voigt": voigt, "voigt_sharded": voigt_sharded, "lut_build": lut_build,
        "gpu_launches": int(launches), "clocks": sampler.summary(),
    }
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, per, desc = cpu_los_sample(args, wl, seconds=args.cpu_seconds, threads=threads)
        line["cpu_baseline"] = {"value": v, "unit": "LOS/s", "cores": threads, "kind": "port",
                                "sample": desc}
        ve, vdesc = cpu_voigt_sample(args, wl, threads)
        line["voigt"]["cpu_baseline"] = {"value": ve, "unit": "evals/s", "cores": threads,
                                         "kind": "port", "sample": vdesc}
        v1, v1desc = cpu_voigt_sample(args, wl, 1, n_sample=100 if args.small else 2000)
        line["voigt"]["cpu_baseline_1thread"] = {"value": v1, "unit": "evals/s", "cores": 1,
                                                 "kind": "port", "sample": v1desc}
        line["voigt"]["context"] = ("the reference author's own estimate of the Python + f2py path is "
                                    "~2.2e6 evals/s (spect_main_module.py:791-801; BASELINE.md section 1)")
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
