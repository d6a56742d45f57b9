#!/usr/bin/env python
"""bench.py -- headline benchmark of the SpectRobot hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement on host cores

Workload (BASELINE.json configs[1], the north_star batch): CH4 3.3 um non-LTE 3-D limb radiances,
spectral grid [2850,3450] cm-1 at 5e-4 (1 200 001 points), 12 vibrational levels, 3e4 synthetic
lines, float32 (P,T) LUT built by the K2 kernel in the same run, 10^4 synthetic VIMS pixels x 3 lines
of sight (tangent height U(350,1050) km, tangent latitude U(-90,90), solar zenith angle U(30,80)
deg per pixel with T_vib(lat, SZA, z) tables), reduced to 36 instrument channels.

A "step" is one pass of the LOS path over the whole 10^4-pixel batch: observer/direction/Sun
vectors -> radtran steps (sr_los_steps_build_rays) -> LUT interpolation + level populations + layer
recursion (k_los_mma, k_los_layers) -> instrument convolution (k_convolve_lowres).
  value     LOS radiances/s of that step with the LUT resident in HBM and the result left on the
            device; CUDA events around the K timed steps, max over ranks.  With N GPUs every rank
            owns one wavenumber slab of the grid for the LUT and for all LOS (strong scaling); the
            only collectives are the all_gather of the step tables and the all_reduce of the
            [n_LOS][n_chan] channel partial sums.
  e2e       the same steps by wall clock with pageable NumPy buffers on both sides (geometry in,
            low-res spectra out), host<->device copies inside.
  roofline  the dominant kernel (k_los_mma, FP64 tensor sub-pipe) timed live with CUDA events on its
            own stream inside the library (sr_prof_*): algorithmic flop / summed launch durations.
Extras carry the other parts of BASELINE.json's metric: "voigt" (K1 line*gridpoint evals/s),
"voigt_1e6" (million-line list), "lut_build" (K2 wall time of the whole LUT), "k3_layers" (recursion
on resident tau/S against the HBM roofline), "jacobian", "hires_host".
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

W0, W1 = 2850.0, 3450.0
N_LEVELS = 12
N_LINES = 30000
N_CHAN = 36


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pixels", type=int, default=10000, help="pixels (3 LOS each) of the batch, whole job")
    ap.add_argument("--lines", type=int, default=N_LINES)
    ap.add_argument("--small", action="store_true", help="tiny sizes (CI / debugging only)")
    ap.add_argument("--no-extras", action="store_true", help="headline legs only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# workload (identical in both arms)
# --------------------------------------------------------------------------------------------
def make_problem(args, want_lines=True):
    from spectrobot_b200 import synthetic as S
    w0, w1 = (2990.0, 3010.0) if args.small else (W0, W1)
    grid = S.spectral_grid(w0, w1)
    n_lines = 600 if args.small else args.lines
    lines = S.line_table(n_lines if want_lines else 10, w0, w1, n_levels=N_LEVELS)
    atm = S.titan_atmosphere()
    # LUT cells: the envelope of the whole tangent-height range in every latitude band
    env = S.limb_los_steps([338.0, 1062.0] * 7, list(range(7)) * 2, [55.0] * 14, atm,
                           lines["level_energies"])
    cells = S.rect_cells(env["pres"][env["pres"] > 1e-6].min() * 0.9, env["pres"].max() * 1.1,
                         env["temp"].min(), env["temp"].max())
    n_pix = 40 if args.small else args.pixels
    centres = np.linspace(grid[0] + (2.0 if args.small else 10.0), grid[-1] - (2.0 if args.small else 10.0), N_CHAN)
    widths = np.full(N_CHAN, 0.4 if args.small else 6.2)            # sigma of a 14.6 cm-1 FWHM
    return dict(S=S, grid=grid, lines=lines, atm=atm, cells=cells, n_pix=n_pix, centres=centres,
                widths=widths)


def batch_geometry(S, n_pix, seed):
    """10^4 pixels x (low, centre, up) LOS: observer positions at 1e5 km, unit directions, and the
    unit vector towards the Sun such that the SZA at the pixel's tangent point is U(30,80) deg."""
    rng = np.random.default_rng(seed)
    n = 3 * n_pix
    tg_alt = np.repeat(rng.uniform(350.0, 1050.0, n_pix), 3) + np.tile([-12.0, 0.0, 12.0], n_pix)
    lat = np.radians(np.repeat(rng.uniform(-90.0, 90.0, n_pix), 3))
    sza = np.radians(np.repeat(rng.uniform(30.0, 80.0, n_pix), 3))
    az = np.repeat(rng.uniform(0.0, 2 * np.pi, n_pix), 3)
    rt = S.R_TITAN_KM + tg_alt
    that = np.stack([np.cos(lat), np.zeros(n), np.sin(lat)], axis=1)          # tangent point direction
    north = np.stack([-np.sin(lat), np.zeros(n), np.cos(lat)], axis=1)
    east = np.tile([0.0, 1.0, 0.0], (n, 1))
    org = rt[:, None] * that + east * np.sqrt(1.0e5 ** 2 - rt ** 2)[:, None]
    perp = np.cos(az)[:, None] * north + np.sin(az)[:, None] * east
    sun = np.cos(sza)[:, None] * that + np.sin(sza)[:, None] * perp
    return org, -east, sun


def workload_config(args, P):
    return {"workload": "CH4 3.3um non-LTE 3D limb LOS radiance batch (BASELINE configs[1], "
                        "radtran_3D_ch4.py shape): geometry -> radtran steps -> LUT LOS integral -> "
                        "%d channels" % N_CHAN,
            "grid": "[%g,%g] cm-1 step 5e-4 (%d points)" % (P["grid"][0], P["grid"][-1], len(P["grid"])),
            "n_levels": N_LEVELS, "n_lines": int(len(P["lines"]["freq"])), "lut_cells": len(P["cells"]),
            "pixels": int(P["n_pix"]), "los": 3 * int(P["n_pix"]), "channels": N_CHAN,
            "sza": "U(30,80) deg per pixel, T_vib(lat band, SZA, z)",
            "l2": "inputs larger than L2 (float32 LUT of tens of GB streamed per step)",
            "small": bool(args.small)}


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons sampled during the timed region: NVML in-process (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints),
    nvidia-smi as the fallback when pynvml is missing."""
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
        return [str(sm), str(mx), act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        period = float(os.environ.get("SR_BENCH_SAMPLE_S", "0.5"))
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


def profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None


# --------------------------------------------------------------------------------------------
# CPU arm (oracle): cpu_baseline of our arm and --impl reference
# --------------------------------------------------------------------------------------------
class CpuSample(object):
    """Bounded sample of the workload for the CPU restatement of the LOS path (LUT interpolation +
    level populations + layer recursion, oracle/sr_oracle.c orc_los_rt, OpenMP over LOS): 2 LOS per
    host thread drawn like the batch's pixels, on a contiguous window of the grid, LUT rows of that
    window built by the CPU restatement of the cell builder.  LOS-equivalents = LOS x window/grid."""

    def __init__(self, args, P, threads):
        from oracle import cpu_oracle as O
        S = P["S"]
        self.O, self.threads = O, threads
        grid, lines = P["grid"], P["lines"]
        self.n_grid = len(grid)
        self.n_pts = 2048 if args.small else 8192
        pt0 = (len(grid) // 2 // 1024) * 1024
        sub = grid[pt0:pt0 + self.n_pts]
        inside = (lines["freq"] > sub[0]) & (lines["freq"] < sub[-1])
        sub_lines = {k: (v[inside] if isinstance(v, np.ndarray) and v.shape[:1] == inside.shape else v)
                     for k, v in lines.items()}
        self.n_los = max(4, 2 * threads)
        rng = np.random.default_rng(S.SEED + 99)
        n_pix = (self.n_los + 2) // 3
        tg = np.repeat(rng.uniform(350.0, 1050.0, n_pix), 3)[:self.n_los] + \
            np.tile([-12.0, 0.0, 12.0], n_pix)[:self.n_los]
        band = np.repeat(rng.integers(0, 7, n_pix), 3)[:self.n_los]
        sza = np.repeat(rng.uniform(30.0, 80.0, n_pix), 3)[:self.n_los]
        st = S.limb_los_steps(tg, band, sza, P["atm"], lines["level_energies"])
        need = set()
        for l in range(self.n_los):
            for k in range(st["n_steps"][l]):
                c, _ = O.lut_weights(P["cells"], st["pres"][l, k], st["temp"][l, k])
                need.update(int(x) for x in c if x >= 0)
        lin = O.line_window_offsets(grid)
        g32 = np.zeros((len(P["cells"]), N_LEVELS, 3, self.n_pts), dtype=np.float32)
        for c in sorted(need):
            Pc, Tc = P["cells"][c]
            g32[c] = O.gcoeff_cell(sub_lines, sub, Tc, Pc, S.CH4_MM, N_LEVELS, n_threads=threads,
                                   lin_grid=lin).astype(np.float32)
        self.lut = dict(g32=g32, pt=np.array(P["cells"]), level_energy=lines["level_energies"], mol=6,
                        iso=1, iso_ratio=S.CH4_RATIO, lte_unidentified=False)
        self.kw = dict(n_steps=st["n_steps"], temp=st["temp"], pres=st["pres"], column=st["column"],
                       tvib=st["tvib"], n_threads=threads)
        self.steps_mean = float(st["n_steps"].mean())
        O.los_rt([self.lut], **self.kw)                        # warm-up

    def run(self, seconds):
        reps, t_used = 0, 0.0
        while reps == 0 or (t_used < seconds and reps < 400):
            t0 = time.perf_counter()
            self.O.los_rt([self.lut], **self.kw)
            t_used += time.perf_counter() - t0
            reps += 1
        per = t_used / reps
        los_equiv = self.n_los * self.n_pts / float(self.n_grid)
        return los_equiv / per, per, reps

    def describe(self, reps):
        return ("%d LOS (2 per thread, all %d threads busy) x %d of %d grid points, %.1f steps/LOS, "
                "%d calls; C restatement (oracle) of LUT interpolation + level populations + layer "
                "recursion; geometry->steps and the convolution are NOT in the CPU time"
                % (self.n_los, self.threads, self.n_pts, self.n_grid, self.steps_mean, reps))


def cpu_voigt_sample(args, P, threads, n_sample=None):
    """CPU restatement of one LUT cell (humliv_bb + G coefficients + line sum) on a line sample."""
    from oracle import cpu_oracle as O
    S = P["S"]
    grid, lines = P["grid"], P["lines"]
    n = min(len(lines["freq"]), n_sample or (200 if args.small else 30000))
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
    P = make_problem(args)
    cs = CpuSample(args, P, threads)
    vals, per, reps = [], [], 0
    for i in range(args.warmup + args.steps):
        v, p, r = cs.run(seconds=1.5)
        if i >= args.warmup:
            vals.append(v)
            per.append(p)
            reps += r
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "LOS radiances/s", "value": value, "unit": "LOS/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * 3 * P["n_pix"] / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, P),
        "note": "the reference (Fortran 77 + Python 2, spect_base_module missing) cannot be built or "
                "run here; this arm times the C restatement in oracle/ on all host threads; "
                "ms_per_step = the whole batch at the sampled rate",
        "cpu_baseline": {"value": value, "unit": "LOS/s", "cores": threads, "kind": "port",
                         "sample": cs.describe(reps)},
        "e2e": {"value": value, "unit": "LOS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    # only the JSON line may go to stdout (NCCL prints its version banner there)
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from spectrobot_b200 import engine, parallel
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
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    P = make_problem(args)
    S, grid, lines, cells, atm = P["S"], P["grid"], P["lines"], P["cells"], P["atm"]
    n_grid, n_cells, n_pix = len(grid), len(cells), P["n_pix"]
    n_los = 3 * n_pix
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fp64_peak = engine.fp64_peak(40000)
    peaks = measured_peaks()
    hbm_peak = peaks["hbm_gbs"] if peaks else 6650.0

    # ---- this rank's wavenumber slab: all cells of the LUT and all LOS on grid[p0:p0+n_slab] ----
    ls_probe = engine.LineSet(parallel.subset_lines(lines, 0, 1), grid, S.CH4_MM, N_LEVELS)
    align = ls_probe.tile_points()
    ls_probe.close()
    p0, n_slab = parallel.shard_slab(n_grid, rank, world, align=align)
    my_lines = parallel.slab_lines(lines, grid, p0, n_slab, align=align)
    ls = engine.LineSet(my_lines, grid, S.CH4_MM, N_LEVELS)
    all_evals = float(len(lines["freq"])) * 13010.0      # whole list, whole grid (per cell)

    # ---- K2: LUT build, every rank all cells of its slab; nothing to gather --------------------
    g32 = engine.lut_tensor(n_cells, N_LEVELS, n_slab)
    barrier()
    t0 = time.perf_counter()
    ls.gcoeff_cells_window(cells, p0, n_slab, out=g32)                  # first build: incl. allocations
    torch.cuda.synchronize()
    first_build_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    ev0.record()
    ls.gcoeff_cells_window(cells, p0, n_slab, out=g32)
    ev1.record()
    torch.cuda.synchronize()
    build_dev_s = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
    build_wall_s = max_over_ranks(time.perf_counter() - t0)
    lut_build = {"metric": "CH4 LUT build time", "value": build_wall_s, "unit": "s", "cells": n_cells,
                 "device_s": build_dev_s, "first_call_s": first_build_s, "evals_per_s": n_cells * all_evals / build_wall_s,
                 "roofline_frac_fp64": 15.0 * n_cells * all_evals / build_wall_s / (world * fp64_peak),
                 "lut_bytes_per_rank": int(n_cells) * N_LEVELS * 3 * int(n_slab) * 4,
                 "partition": "wavenumber slabs: every rank builds all cells on its %d points; no "
                              "gather, no replication" % n_slab}
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=lines["level_energies"])
    grid_slab = torch.as_tensor(np.ascontiguousarray(grid[p0:p0 + n_slab]), device="cuda")
    cdev = torch.as_tensor(P["centres"], device="cuda")
    wdev = torch.as_tensor(P["widths"], device="cuda")

    # ---- the batch: geometry of the whole job (every rank sees all pixels) ---------------------
    org, dirs, sun = batch_geometry(S, n_pix, S.SEED + 7)
    tv3 = S.vib_temperatures_3d(atm["z"], atm["temp"], lines["level_energies"])   # [lev][band][sza][z]
    A = engine.Atmosphere(atm["z"], atm["temp"], atm["pres"], np.full((1,) + atm["temp"].shape, 0.015),
                          tvib=tv3[None], lat_edges=atm["lat_edges"], radius_km=S.R_TITAN_KM,
                          top_km=1500.0, sza_nodes=S.SZA_NODES)
    b_los, e_los = parallel.shard_range(n_los, rank, world)      # the step BUILDER is sharded by LOS
    info = {}

    def step(to_host):
        t_s = time.perf_counter()
        st_loc, _ = engine.los_steps_build(A, org[b_los:e_los], dirs[b_los:e_los], sun=sun[b_los:e_los],
                                           delta_x=5.0, max_T_variation=5.0, max_Plog_variation=1.0,
                                           n_steps_max=info.get("width", 64))
        info["width"] = st_loc.n_steps_max      # the caller sizes the tables: no trimming copy next time
        st_all = parallel.allgather_steps(st_loc, n_los, rank, world)
        info["steps_s"] = time.perf_counter() - t_s
        low = engine.los_rt_lut_lowres([lut], st_all, grid_slab, cdev, wdev, check_status=False)
        parallel.allreduce_lowres(low)
        info["steps"] = st_all
        info["h2d"] = int(org[b_los:e_los].nbytes * 3 + st_all.n_steps.nbytes + st_all.temp.nbytes +
                          st_all.pres.nbytes + st_all.column.nbytes + st_all.tvib.nbytes)
        info["d2h"] = int(st_loc.n_steps.nbytes + st_loc.temp.nbytes + st_loc.pres.nbytes +
                          st_loc.column.nbytes + st_loc.tvib.nbytes)
        if to_host:
            out = np.empty(tuple(low.shape))                     # pageable
            out[...] = low.cpu().numpy()
            info["d2h"] += int(out.nbytes)
            return out
        return low

    sampler = ClockSampler(local)
    for _ in range(args.warmup):
        low = step(False)
    engine.check(lib.sr_los_check(engine._lut_array([lut]), engine._stream_ptr()))
    barrier()
    engine.prof_enable(True)
    l0 = lib.sr_kernel_launch_count()
    sampler.start()
    ev0.record()
    for _ in range(args.steps):
        low = step(False)
    ev1.record()
    barrier()
    step_ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
    launches = lib.sr_kernel_launch_count() - l0
    prof = engine.prof_summary()
    engine.prof_enable(False)
    value = n_los / (step_ms * 1e-3)
    low_ref = low.clone()
    finite = bool(torch.isfinite(low_ref).all().item() and (low_ref > 0).any().item())
    steps_mean = float(info["steps"].n_steps.mean())

    # e2e: the same step by wall clock, pageable NumPy on both sides
    e2e_each = []
    for _ in range(max(1, min(args.steps, 3))):
        barrier()
        t0 = time.perf_counter()
        out_host = step(True)
        e2e_each.append(time.perf_counter() - t0)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_s = max_over_ranks(float(np.median(e2e_each)))
    same = bool(np.array_equal(out_host, low_ref.cpu().numpy()))

    # roofline of the dominant kernel from the library's own events (summed over ranks)
    kern = {}
    for name in ("los_mma", "los_fused", "los_layers", "conv"):      # sums over ranks
        n_l, ms, work = prof.get(name, (0, 0.0, 0.0))
        tot = (sum_over_ranks(n_l), sum_over_ranks(ms), sum_over_ranks(work))
        if tot[0] > 0:
            kern[name] = {"launches": tot[0], "ms": tot[1], "work": tot[2]}
    dom = "los_fused" if "los_fused" in kern else "los_mma"
    roof = {"bound": "tensor", "unit": "TFLOP/s", "kernel": "k_%s (FP64 DMMA.8x8x4)" % dom,
            "achieved": None, "peak": fp64_peak / 1e12, "frac": None, "traffic": None}
    if dom in kern and kern[dom]["ms"] > 0:
        k = kern[dom]
        ach = k["work"] / (k["ms"] * 1e-3) / 1e12               # per GPU: flop / GPU-seconds in the kernel
        roof.update({"achieved": ach, "frac": ach / (fp64_peak / 1e12),
                     "launches": int(k["launches"]), "ms_per_launch": k["ms"] / k["launches"],
                     "flop_per_launch": k["work"] / k["launches"],
                     "share_of_step": k["ms"] / world / (args.steps * step_ms),
                     "peak_source": "live sr_fp64_peak (no FP64 entry in MEASURED_PEAKS.json; nominal 37.2)",
                     "algorithmic": "2 flop x non-zero LUT rows of the 4 cells per (LOS, step, point)"})
        tr = profile_json("r2_los_traffic.json")
        if tr and not args.small:
            roof["traffic"] = tr.get("traffic_per_launch")
            roof["traffic_source"] = tr.get("source")
    kernels = {n: {"ms_per_step_per_gpu": k["ms"] / world / args.steps,
                   "launches_per_step_per_gpu": k["launches"] / args.steps / world}
               for n, k in kern.items()}
    if "los_layers" in kern and kern["los_layers"]["ms"] > 0:
        kernels["los_layers"]["hbm_frac"] = (kern["los_layers"]["work"] /
                                             (kern["los_layers"]["ms"] * 1e-3) / 1e9 / hbm_peak)

    extras = {}
    if not args.no_extras:
        extras = run_extras(args, P, engine, parallel, torch, dist, ls, lut, info["steps"], grid_slab,
                            cdev, wdev, p0, n_slab, rank, world, barrier, max_over_ranks, fp64_peak,
                            hbm_peak, all_evals)

    # key order: long descriptive objects first, the headline numbers last (a reader of the tail of
    # the line sees value / e2e / roofline.frac)
    line = {"lut_build": lut_build}
    line.update(extras)
    line.update({
        "kernels": kernels,
        "batch": {"steps_per_los_mean": steps_mean, "finite_positive": finite,
                  "checksum": float(low_ref.sum().item()),
                  "steps_build_and_gather_s": max_over_ranks(info["steps_s"]),
                  "e2e_equals_device_result": same, "slab_points_per_rank": int(n_slab),
                  "collectives": "all_gather(step tables), all_reduce([n_los][%d] partial sums)" % N_CHAN
                  if world > 1 else "none"},
        "config": workload_config(args, P),
        "cpu_baseline": None,
        "roofline": roof,
        "clocks": sampler.summary(),
        "e2e": {"value": n_los / e2e_s, "unit": "LOS/s", "h2d_bytes_per_step": info["h2d"],
                "d2h_bytes_per_step": info["d2h"], "s_per_step": e2e_s,
                "path": "sr_los_steps_build_rays + sr_los_rt_lut_channels_dev, pageable NumPy in and out"},
        "metric": "LOS radiances/s", "value": value, "unit": "LOS/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "gpu_launches": int(launches),
    })
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cs = CpuSample(args, P, threads)
        v, per, reps = cs.run(args.cpu_seconds)
        line["cpu_baseline"] = {"sample": cs.describe(reps), "kind": "port", "cores": threads,
                                "unit": "LOS/s", "value": v}
        if "voigt" in line:
            ve, vdesc = cpu_voigt_sample(args, P, threads)
            line["voigt"]["cpu_baseline"] = {"value": ve, "unit": "evals/s", "cores": threads,
                                             "kind": "port", "sample": vdesc}
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_extras(args, P, engine, parallel, torch, dist, ls, lut, steps_all, grid_slab, cdev, wdev, p0,
               n_slab, rank, world, barrier, max_over_ranks, fp64_peak, hbm_peak, all_evals):
    """The other parts of BASELINE.json's metric and the secondary paths, each a few seconds."""
    S, grid, lines = P["S"], P["grid"], P["lines"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}

    def timed(fn, reps, skip=1):
        ms = []
        for i in range(reps + skip):
            barrier()
            ev0.record()
            fn(i)
            ev1.record()
            barrier()
            if i >= skip:
                ms.append(ev0.elapsed_time(ev1))
        return max_over_ranks(float(np.median(ms))) * 1e-3

    # ---- K1: evals/s on FP64 cells (16 cells per call, this rank's slab) ----------------------
    n_k1 = 2 if args.small else 16
    buf = torch.empty((n_k1, N_LEVELS, 3, n_slab), dtype=torch.float64, device="cuda")
    pts = lambda i: [[0.05 * (1 + i + j), 150.0 + 2.0 * j] for j in range(n_k1)]   # noqa: E731
    k1_s = timed(lambda i: ls.gcoeff_cells_window(pts(i), p0, n_slab, f32=False, out=buf), 4, 2) / n_k1
    out["voigt"] = {"metric": "Voigt line*gridpoint evals/s", "value": all_evals / k1_s, "unit": "evals/s",
                    "ms_per_cell": 1e3 * k1_s, "cells_per_call": n_k1, "lines": int(len(lines["freq"])),
                    "roofline": {"bound": "fp64", "achieved": 15.0 * all_evals / k1_s / 1e12 / world,
                                 "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
                                 "frac": 15.0 * all_evals / k1_s / fp64_peak / world,
                                 "note": "ALGORITHMIC: 15 FP64 flop per line*gridpoint (SURVEY 8d) x the "
                                         "evals of the problem / time of k_line_cell_params + k_core_eval "
                                         "+ k_far_nodes + k_voigt_tile, per GPU.  Distant full far wings are "
                                         "evaluated at 12 Chebyshev nodes per 512-point tile and interpolated "
                                         "(<= 4e-11 of a line's own value), so fewer flops are executed than "
                                         "counted; SR_K1_FAR=0 evaluates every point (frac 0.41)"}}
    del buf

    # ---- K1 on a million-line list: wavenumber slabs, result stays sharded, no collective ------
    n_big = 20000 if args.small else 1000000
    big = S.line_table(n_big, grid[0], grid[-1], n_levels=N_LEVELS, seed=20067)
    ls_big = engine.LineSet(parallel.slab_lines(big, grid, p0, n_slab, align=ls.tile_points()), grid, S.CH4_MM, N_LEVELS)
    xs_big = torch.empty((1, N_LEVELS, 3, n_slab), dtype=torch.float64, device="cuda")
    big_s = timed(lambda i: ls_big.gcoeff_cells_window([[0.02, 155.0]], p0, n_slab, f32=False, out=xs_big), 3)
    out["voigt_1e6"] = {"metric": "Voigt evals/s, 1e6-line list (BASELINE configs[4])",
                        "value": n_big * 13010.0 / big_s, "unit": "evals/s", "ms": 1e3 * big_s,
                        "roofline_frac_fp64": 15.0 * n_big * 13010.0 / big_s / (world * fp64_peak),
                        "partition": "wavenumber slabs (each rank: the lines whose window reaches its "
                                     "slab), no collective; result stays sharded"}
    ls_big.close()
    del ls_big, xs_big, big

    # ---- K3 alone on resident tau/S (north_star's LOS integral proper): HBM roofline -----------
    n_k3 = 6 if args.small else 36
    st36 = steps_all.subset(slice(0, n_k3))
    tau, src = engine.los_tau_src([lut], st36)
    nst = torch.tensor(st36.n_steps, dtype=torch.int32, device="cuda")
    rad = torch.empty((n_k3, n_slab), dtype=torch.float64, device="cuda")
    k3_s = timed(lambda i: engine.los_rt_layers(tau, src, nst, out=rad), 5, 3)
    k3_bytes = (16.0 * float(st36.n_steps.sum()) + 8.0 * n_k3) * n_slab
    out["k3_layers"] = {"value": n_k3 / k3_s, "unit": "LOS/s", "los": n_k3, "ms": 1e3 * k3_s,
                        "roofline": {"bound": "hbm", "achieved": k3_bytes / k3_s / 1e9, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": k3_bytes / k3_s / 1e9 / hbm_peak,
                                     "kernel": "k_los_layers", "algorithmic_bytes_per_launch": k3_bytes,
                                     "peak_source": "MEASURED_PEAKS.json hbm_gbs"}}
    rad_k3 = rad.clone()
    del tau, src
    rad_f = engine.los_rt_lut([lut], st36)
    out["k3_layers"]["max_rel_diff_vs_fused_path"] = float(
        ((rad_f - rad_k3).abs() / rad_k3.abs().clamp_min(1e-300)).max().item())
    del rad_f, rad_k3, rad

    # ---- forward + analytic Jacobians (BASELINE configs[4] shape), low-res on the device --------
    n_j = 12 if args.small else 360
    n_par = 12
    st_j = steps_all.subset(slice(0, n_j))
    ns_j = st_j.n_steps
    kk = np.arange(st_j.n_steps_max)[None, :]
    node = np.minimum(kk * n_par // np.maximum(ns_j[:, None], 1), n_par - 1)
    wgt = 0.25 + 0.5 * ((kk * 7919) % 97) / 97.0
    dfrac = np.zeros((n_j, st_j.n_steps_max, n_par))
    ii = np.arange(n_j)[:, None].repeat(st_j.n_steps_max, 1)
    np.add.at(dfrac, (ii, kk.repeat(n_j, 0), node), np.broadcast_to(wgt, node.shape))
    np.add.at(dfrac, (ii, kk.repeat(n_j, 0), np.minimum(node + 1, n_par - 1)),
              np.broadcast_to(1.0 - wgt, node.shape))
    dfrac *= (kk < ns_j[:, None])[:, :, None]

    def jac_call(i):
        lo, jl = engine.los_rt_lut_jac_lowres([lut], st_j, dfrac, grid_slab, cdev, wdev, check_status=False)
        parallel.allreduce_lowres(lo)
        parallel.allreduce_lowres(jl)
        jac_call.res = (lo, jl)
    jac_s = timed(jac_call, 3)
    out["jacobian"] = {"value": n_j / jac_s, "unit": "LOS/s (radiance + %d derivative spectra each)" % n_par,
                       "los": n_j, "ms": 1e3 * jac_s, "finite": bool(torch.isfinite(jac_call.res[1]).all().item())}
    del dfrac

    # ---- hi-res radiances to host memory (radtran_fast's own contract), pageable buffer ---------
    n_h = 9 if args.small else 120
    st_h = steps_all.subset(slice(0, n_h))
    host_out = np.empty((n_h, n_slab))
    engine.los_rt_lut_host([lut], st_h, out=host_out)
    hs = []
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        engine.los_rt_lut_host([lut], st_h, out=host_out)
        hs.append(time.perf_counter() - t0)
    h_s = max_over_ranks(min(hs))
    out["hires_host"] = {"value": n_h / h_s, "unit": "LOS/s", "los": n_h, "d2h_bytes": int(host_out.nbytes),
                         "path": "sr_los_rt_lut_host -> pageable NumPy [n_los][slab points]"}
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
