"""world_size-2 gloo tests (CPU) of the multi-GPU partitioning logic (spectrobot_b200/parallel.py):
cell sharding + LUT gather, LOS sharding + row gather, line sharding + all_reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spectrobot_b200 import parallel


def test_shard_arithmetic():
    for n in (0, 1, 7, 273, 30000):
        for w in (1, 2, 3, 8):
            cells = sorted(sum((parallel.shard_cells(n, r, w) for r in range(w)), []))
            assert cells == list(range(n))
            rng = [parallel.shard_range(n, r, w) for r in range(w)]
            assert rng[0][0] == 0 and rng[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rng[:-1], rng[1:]))
            sizes = [e - b for b, e in rng]
            assert max(sizes) - min(sizes) <= 1
    tab = dict(freq=np.arange(10.0), up_set=np.arange(10), n_sets=3, level_energies=np.zeros(3))
    sub = parallel.subset_lines(tab, 2, 5)
    assert list(sub["freq"]) == [2.0, 3.0, 4.0] and sub["n_sets"] == 3 and len(sub["level_energies"]) == 3


def test_fed_rows():
    tab = dict(up_set=np.array([1, 1, 2, -1, 3]), lo_set=np.array([0, 0, 0, 0, -1]))
    # lines 0-2 are linked: upper sets 1, 2 -> sp/ind rows; lower set 0 -> absorption row
    assert parallel.fed_rows(tab, 4) == [2, 3, 4, 6, 7]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, n, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(n), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=n)
    try:
        r, w, _ = parallel.world()
        assert (r, w) == (rank, n)
        # LUT cells: every rank fills its own cells, gather completes the table
        n_cells = 7
        full = torch.arange(n_cells * 6, dtype=torch.float32).reshape(n_cells, 2, 3)
        mine = torch.zeros_like(full)
        idx = parallel.shard_cells(n_cells, rank, n)
        mine[idx] = full[idx]
        got = parallel.gather_lut(mine, n_cells, rank, n)
        ok_lut = bool(torch.equal(got, full))
        # LOS rows: contiguous blocks, unequal sizes
        n_los = 5
        rows = torch.arange(n_los * 4, dtype=torch.float64).reshape(n_los, 4)
        b, e = parallel.shard_los(n_los, rank, n)
        ok_rows = bool(torch.equal(parallel.gather_rows(rows[b:e].clone(), n_los, rank, n), rows))
        # line-sharded partial spectra: the sum over ranks equals the full sum
        rng = np.random.default_rng(3)
        contrib = torch.as_tensor(rng.uniform(size=(11, 2, 3, 16)))
        b, e = parallel.shard_lines(11, rank, n)
        part = contrib[b:e].sum(dim=0)
        parallel.allreduce_spectra(part)
        ok_sum = bool(torch.allclose(part, contrib.sum(dim=0), rtol=1e-14, atol=0))
        # same with the rows the line list can feed (the others are zero on every rank)
        rows_fed = [0, 1, 5]
        mask = torch.zeros(6, dtype=torch.float64)
        mask[rows_fed] = 1.0
        sparse = contrib.reshape(11, 1, 6, 16) * mask[None, None, :, None]
        part4 = sparse[b:e].sum(dim=0).reshape(1, 2, 3, 16).clone()
        parallel.allreduce_spectra(part4, rows=rows_fed)
        ok_sum = ok_sum and bool(torch.allclose(part4, sparse.sum(dim=0).reshape(1, 2, 3, 16),
                                                rtol=1e-14, atol=0))
        q.put((rank, ok_lut, ok_rows, ok_sum))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1] and r[2] and r[3], r


def test_slab_arithmetic():
    """Wavenumber slabs: aligned starts, one shared boundary point, exact cover."""
    for n in (1200001, 800001, 1025, 1024, 513, 100):
        for w in (1, 2, 3, 8):
            sl = [parallel.shard_slab(n, r, w) for r in range(w)]
            used = [s for s in sl if s[1] > 0]
            assert used[0][0] == 0 and used[-1][0] + used[-1][1] == n
            for a, b in zip(used[:-1], used[1:]):
                assert b[0] % 512 == 0 and a[0] + a[1] - 1 == b[0]    # shared point
            # trapezoid segments [i, i+1] are partitioned: every segment belongs to one slab
            segs = sum(s[1] - 1 for s in used)
            assert segs == n - 1
    grid = 2850.0 + 5e-4 * np.arange(40000)
    tab = dict(freq=np.array([2850.1, 2853.0, 2856.6, 2860.0, 2869.9]), up_set=np.zeros(5, int),
               level_energies=np.zeros(2))
    sub = parallel.slab_lines(tab, grid, 10240, 2049)     # slab [2855.12, 2856.144]
    # half a window is 3.2525 cm-1 (+ one 512-point tile = 0.256 + rounding margin): 2853.0 and
    # 2856.6 reach it, 2850.1 and 2860.0 do not
    assert list(sub["freq"]) == [2853.0, 2856.6] and len(sub["level_energies"]) == 2


def _trapz_channels(grid, spec, centres, widths, n_sigma=5.0):
    """convolve_to_grid_from_irregular (spect_classes.py:883-918) on a slab of the grid."""
    out = np.zeros((spec.shape[0], len(centres)))
    for c, (f, s) in enumerate(zip(centres, widths)):
        m = (grid >= f - n_sigma * s) & (grid <= f + n_sigma * s)
        if m.sum() < 2:
            continue
        g = np.exp(-0.5 * ((grid[m] - f) / s) ** 2) / (s * np.sqrt(2 * np.pi))
        out[:, c] = np.trapezoid(spec[:, m] * g, grid[m], axis=1)
    return out


def _slab_worker(rank, n, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(n), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=n)
    try:
        from spectrobot_b200 import engine
        # (1) low-res partial sums of the slabs add up to the convolution of the whole spectrum
        n_grid = 4097
        grid = 3000.0 + 5e-4 * np.arange(n_grid)
        rng = np.random.default_rng(5)
        spec = rng.uniform(size=(3, n_grid))
        centres, widths = np.array([3000.3, 3001.0, 3001.024, 3001.9]), np.array([0.05, 0.08, 0.02, 0.1])
        p0, m = parallel.shard_slab(n_grid, rank, n)
        part = torch.as_tensor(_trapz_channels(grid[p0:p0 + m], spec[:, p0:p0 + m], centres, widths))
        parallel.allreduce_lowres(part)
        full = _trapz_channels(grid, spec, centres, widths)
        ok_low = bool(np.allclose(part.numpy(), full, rtol=1e-12, atol=0))
        # (2) step tables built per LOS shard (different widths per rank) -> all LOS on every rank
        n_los = 5
        b, e = parallel.shard_range(n_los, rank, n)
        w = 3 + rank                                        # rank 1's tables are wider
        mk = lambda *shape: rng.uniform(size=shape)         # noqa: E731
        full_t = np.arange(n_los * 4, dtype=float).reshape(n_los, 4)
        loc = engine.LosSteps(np.full(e - b, 3, np.int32), full_t[b:e, :w], full_t[b:e, :w] * 2,
                              np.stack([full_t[b:e, :w] * 3, full_t[b:e, :w] * 4]),
                              np.stack([full_t[b:e, :w] * 5, full_t[b:e, :w] * 6]).reshape(2, 1, e - b, w))
        allst = parallel.allgather_steps(loc, n_los, rank, n, device="cpu")
        ok_st = allst.n_los == n_los and allst.n_steps_max == 4 and allst.n_gas == 2
        ok_st = ok_st and bool(np.array_equal(allst.temp[:, :3], full_t[:, :3]))
        ok_st = ok_st and bool(np.array_equal(allst.pres[b:e, :w], 2 * full_t[b:e, :w]))
        ok_st = ok_st and bool(np.array_equal(allst.column[1][:, :3], 4 * full_t[:, :3]))
        ok_st = ok_st and bool(np.array_equal(allst.tvib[1, 0][:, :3], 6 * full_t[:, :3]))
        ok_st = ok_st and bool(np.array_equal(allst.n_steps, np.full(n_los, 3)))
        q.put((rank, ok_low, ok_st))
    finally:
        dist.destroy_process_group()


def test_gloo_slab_partition_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] and r[2], r
