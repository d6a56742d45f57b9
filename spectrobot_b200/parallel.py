"""Multi-GPU partitioning of the hot path: one process per GPU, `torch.distributed` for the
plumbing (NCCL on the GPUs; the same code runs on gloo/CPU tensors in the tests).

The path shards where the reference itself fans out over processes:

  LUT cells    independent (`for [Pres,Temp] in PTcouples`, spect_main_module.py:753) -> contiguous
               blocks over ranks, no collective on the data path, one gather of the float32 LUT at the end
               if every rank needs the whole table (gather_lut);
  LOS batch    independent (one forked process per LOS in the reference, spect_main_module.py:
               3202-3221) -> contiguous blocks per rank, gather of the (low-res) spectra (gather_rows);
  one huge line list (config 5)  the sum over lines is linear (lineshape.f:17-23) -> lines split by
               index, all_reduce(SUM, fp64) of the partial [set][3][n_grid] spectra (allreduce_spectra).

Wavenumber slabs (the partition the whole job scales on): every grid point of a
cross-section, of a LUT row and of a hi-res radiance is independent of every other one, so rank r
owns ONE contiguous slab of the spectral grid for everything - all (P,T) cells of the LUT on its
slab (slab_lines: only the lines whose Voigt window reaches it), all lines of sight on its slab -
and the only exchange of the LOS batch is the all_reduce of the [n_LOS][n_chan] low-res partial
sums (allreduce_lowres), because a channel's Gaussian window may straddle slabs.  Neighbouring
slabs share their boundary point so that every trapezoid of the convolution is counted once.
"""
import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process: 0, 1, 0)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend=None):
    """Initialise torch.distributed from the environment (no-op for a single process)."""
    import torch
    import torch.distributed as dist
    rank, n, local = world()
    if n == 1 or dist.is_initialized():
        return rank, n, local
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        from . import engine
        engine.lib().sr_set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, n, local


def shard_cells(n_cells, rank, n_ranks):
    """Indices of the (P,T) cells rank builds: one contiguous block per rank (the per-cell cost is
    uniform to ~2 % across the pressure ladder), so that a rank writes its cells straight into its
    slice of the LUT tensor and the gather moves contiguous memory."""
    b, e = shard_range(n_cells, rank, n_ranks)
    return list(range(b, e))


def shard_range(n, rank, n_ranks):
    """Contiguous [begin, end) block of n items for rank (sizes differ by at most one)."""
    base, extra = divmod(n, n_ranks)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_los(n_los, rank, n_ranks):
    return shard_range(n_los, rank, n_ranks)


def shard_lines(n_lines, rank, n_ranks):
    """Lines are split by index; with a frequency-sorted list every rank gets a contiguous
    wavenumber interval, but the result does not depend on that (the sum is linear)."""
    return shard_range(n_lines, rank, n_ranks)


def shard_slab(n_grid, rank, n_ranks, align=512):
    """(pt0, n_pts) of rank's wavenumber slab: grid points [pt0, pt0+n_pts).  Slab starts are
    multiples of `align` (the tile size of the cross-section kernel, LineSet.tile_points());
    consecutive slabs SHARE one boundary point (the last point of slab r is the first of slab r+1),
    so that the trapezoid segments of the instrument convolution partition exactly."""
    n_grid, align = int(n_grid), int(align)
    n_tiles = (n_grid + align - 1) // align
    used = max(1, min(int(n_ranks), n_tiles))          # ranks that get a slab
    if rank >= used:                                   # more ranks than tiles: nothing to do
        return n_grid - 1, 0
    b, e = shard_range(n_tiles, rank, used)
    pt0 = b * align
    end = n_grid if rank == used - 1 else min(e * align + 1, n_grid)   # +1: the shared point
    return pt0, end - pt0


def slab_lines(tab, grid, pt0, n_pts, align=512, half_window=None):
    """The lines whose 13010-point Voigt window can reach a TILE of the slab [pt0, pt0+n_pts): line
    centre within half a window + one tile (`align` grid points, the cross-section kernel's tile)
    + closest_grid rounding of the slab.  Returns the sub-table.  A LineSet built from it on the
    WHOLE grid gives the slab bit-identically to a full build: every tile that overlaps the slab
    sees exactly the candidate lines (and therefore the summation order) of the full line list."""
    from ._lib import IMXSIG
    freq = np.asarray(tab["freq"])
    step = float(grid[1] - grid[0])
    hw = (IMXSIG // 2 + 2 + int(align)) * step if half_window is None else float(half_window)
    lo, hi = grid[pt0] - hw, grid[pt0 + n_pts - 1] + hw
    keep = (freq >= lo) & (freq <= hi)
    n = len(freq)
    return {k: (v[keep] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v)
            for k, v in tab.items()}


def allreduce_lowres(low):
    """Sum of the per-slab partial channel integrals [n_los, n_chan] (or [n_los, n_par, n_chan])
    over the ranks: the one collective of the wavenumber-sharded LOS batch."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(low, op=dist.ReduceOp.SUM)
    return low


def allgather_steps(local, n_total, rank, n_ranks, device=None):
    """Step tables (engine.LosSteps) built by every rank for its shard_range block of the LOS ->
    the tables of all n_total LOS on every rank (the step builder is sharded by LOS, the LOS
    integral by wavenumber).  One all_gather per table on the device (NVLink), tables padded to the
    widest rank and the longest block; the result comes back through pinned host memory."""
    import torch
    import torch.distributed as dist
    from . import engine
    if n_ranks == 1:
        return local
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    w = torch.tensor([local.n_steps_max], dtype=torch.int64, device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    w = int(w.item())
    sizes = [shard_range(n_total, r, n_ranks) for r in range(n_ranks)]
    nmax = max(e - b for b, e in sizes)

    def gather(a, los_axis, fill):
        """numpy [..., n_los_local, w_local] -> numpy [..., n_total, w] (LOS axis at los_axis)"""
        t = torch.as_tensor(a).to(dev)
        if los_axis != 0:
            t = t.movedim(los_axis, 0)
        shape = (nmax,) + tuple(t.shape[1:-1]) + ((w,) if a.ndim > 1 else ())
        pad = torch.full(shape, fill, dtype=t.dtype, device=dev)
        if a.ndim > 1:
            pad[:t.shape[0], ..., :t.shape[-1]] = t
        else:
            pad[:t.shape[0]] = t
        out = torch.empty((n_ranks * shape[0],) + shape[1:], dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, pad)
        out = out.view((n_ranks,) + shape)
        full = torch.cat([out[r, :e - b] for r, (b, e) in enumerate(sizes)], dim=0)
        if los_axis != 0:
            full = full.movedim(0, los_axis)
        full = full.contiguous()
        if full.is_cuda:
            host = torch.empty(full.shape, dtype=full.dtype, pin_memory=True)
            host.copy_(full)
            return host.numpy()
        return full.numpy()

    n_steps = gather(local.n_steps, 0, 0)
    temp = gather(local.temp, 0, 100.0)
    pres = gather(local.pres, 0, 1e-6)
    col = gather(local.column, 1, 0.0)
    tvib = None if local.tvib is None else gather(local.tvib, 2, 100.0)
    return engine.LosSteps(n_steps, temp, pres, col, tvib)


def gather_lut(g32, n_cells, rank, n_ranks):
    """Complete a LUT tensor [n_cells, ...] in which every rank has filled shard_cells(rank):
    the owner of each block broadcasts it (every cell crosses NVLink once per receiver)."""
    import torch.distributed as dist
    if n_ranks == 1:
        return g32
    buf = g32
    if g32.is_cuda and not g32.is_contiguous():   # row-padded LUT view: send the padded rows
        from . import engine
        buf = engine.lut_padded(g32)
    work = []
    for src in range(n_ranks):
        b, e = shard_range(n_cells, src, n_ranks)
        if e > b:   # contiguous slice, received in place; all owners send at the same time
            work.append(dist.broadcast(buf[b:e], src=src, async_op=True))
    for w in work:
        w.wait()
    return g32


def gather_rows(local, n_total, rank, n_ranks):
    """All ranks' row blocks (shard_range partition of n_total rows) -> [n_total, ...] on every
    rank."""
    import torch
    import torch.distributed as dist
    if n_ranks == 1:
        return local
    sizes = [shard_range(n_total, r, n_ranks) for r in range(n_ranks)]
    nmax = max(e - b for b, e in sizes)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(n_ranks)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:e - b] for p, (b, e) in zip(parts, sizes)], dim=0)


def fed_rows(tab, n_sets):
    """(set, ctype) rows of the [n_sets][3] cross-section block that the line list can feed: a
    line adds to sp_emission / ind_emission of its upper set and to absorption of its lower set
    (spect_classes.py:1304-1313); every other row is identically zero on every rank."""
    up, lo = np.asarray(tab["up_set"]), np.asarray(tab["lo_set"])
    ok = (up >= 0) & (lo >= 0)
    rows = set()
    for u in np.unique(up[ok]):
        rows.update((int(u) * 3 + 0, int(u) * 3 + 1))
    for l in np.unique(lo[ok]):
        rows.add(int(l) * 3 + 2)
    return sorted(r for r in rows if r < 3 * n_sets)


def allreduce_spectra(t, rows=None):
    """In-place sum of partial cross-section spectra [n_cells][n_sets][3][n_grid] over the ranks
    (line-sharded K1).  rows: the (set*3 + ctype) rows that can be non-zero (fed_rows); only those
    cross NVLink (packed into one contiguous buffer), the others stay zero."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return t
    if rows is None or len(rows) >= t.shape[1] * t.shape[2]:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t
    flat = t.view(t.shape[0], t.shape[1] * t.shape[2], t.shape[3])
    idx = torch.as_tensor(rows, dtype=torch.long, device=t.device)
    packed = flat.index_select(1, idx)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    flat.index_copy_(1, idx, packed)
    return t


def subset_lines(tab, begin, end):
    """Rows [begin, end) of a line-table dict (see spect_classes.line_table / synthetic)."""
    n = len(tab["freq"])
    return {k: (v[begin:end] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v)
            for k, v in tab.items()}


def gcoeff_cells_line_sharded(tab, grid, MM, n_sets, PTcouples):
    """Cross-sections of a line list too large for one GPU's time budget: every rank evaluates its
    share of the lines for ALL cells, then one all_reduce sums the partial spectra."""
    from . import engine
    rank, n, _ = world()
    b, e = shard_lines(len(tab["freq"]), rank, n)
    ls = engine.LineSet(subset_lines(tab, b, e), grid, MM, n_sets)
    out = ls.gcoeff_cells(PTcouples)
    ls.close()
    return allreduce_spectra(out, rows=fed_rows(tab, n_sets))
