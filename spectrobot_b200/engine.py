"""Host-side handles over the Tier-2 C ABI (include/spectrobot.h).

PyTorch is used for what it is good at here: owning device memory and streams.  All arithmetic
of the hot path happens inside libspectrobot.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import IMXSIG, as_f64, as_i32, check, dptr, iptr, lib

CTYPES = ('sp_emission', 'ind_emission', 'absorption')   # spect_classes.py:313

_LINE_FIELDS = ("freq", "a_coeff", "air_broad", "t_dep", "e_lower", "g_up", "g_lo", "e_vib_up",
                "e_vib_lo")


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("spectrobot_b200 needs a CUDA device (no CPU fallback)")
    return torch


def _stream_ptr(stream=None):
    torch = _torch()
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def line_window_offsets(grid):
    """Window offsets of PrepareCalcShapes (spect_classes.py:1445-1446): built from the ACTUAL
    grid step grid[1]-grid[0], and required to have exactly imxsig points (humliv_bb's f2py
    signature fixes the length, lineshape.f:236)."""
    sp_step = grid[1] - grid[0]
    lin = np.arange(-IMXSIG * sp_step / 2, IMXSIG * sp_step / 2, sp_step, dtype=float)
    if len(lin) != IMXSIG:
        raise ValueError("line window has %d points; humliv_bb needs exactly %d (imxsig)"
                         % (len(lin), IMXSIG))
    return lin


class LineSet(object):
    """Device-resident line table of one isotopologue bound to one spectral grid.

    lines: dict of equal-length arrays (freq, a_coeff, air_broad, t_dep, e_lower, g_up, g_lo,
    e_vib_up, e_vib_lo float64; up_set, lo_set int32), see sr_lines in spectrobot.h.
    """

    def __init__(self, lines, grid, MM, n_sets, consts=None, lin_grid=None):
        self._h = C.c_void_p()
        self.grid = as_f64(grid)
        self.n_grid = len(self.grid)
        self.n_sets = int(n_sets)
        self.MM = float(MM)
        self.lin_grid = as_f64(line_window_offsets(self.grid) if lin_grid is None else lin_grid)
        if len(self.lin_grid) != IMXSIG:
            raise ValueError("lin_grid must have %d points" % IMXSIG)
        keep = [as_f64(lines[k]) for k in _LINE_FIELDS]
        up = as_i32(lines["up_set"])
        lo = as_i32(lines["lo_set"])
        n = len(up)
        for a in keep:
            if len(a) != n:
                raise ValueError("line arrays have different lengths")
        st = _lib.sr_lines()
        st.n_lines = n
        for name, a in zip(_LINE_FIELDS, keep):
            setattr(st, name, dptr(a))
        st.up_set = iptr(up)
        st.lo_set = iptr(lo)
        self.consts = consts if consts is not None else _lib.python_consts()
        self.n_lines = n
        check(lib().sr_lineset_create(C.byref(st), dptr(self.grid), self.n_grid,
                                      dptr(self.lin_grid), self.n_sets, self.MM,
                                      C.byref(self.consts), C.byref(self._h)))
        self.n_active = int(lib().sr_lineset_n_active(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().sr_lineset_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def centres(self):
        """closest_grid index of every input line (-1 for lines dropped by the level filter)."""
        out = np.empty(self.n_lines, dtype=np.int32)
        check(lib().sr_lineset_centres(self._h, iptr(out)))
        return out

    def order(self):
        out = np.empty(self.n_active, dtype=np.int32)
        check(lib().sr_lineset_order(self._h, iptr(out)))
        return out

    def cells_elems(self, n_cells):
        return int(n_cells) * self.n_sets * 3 * self.n_grid

    def gcoeff_cells(self, PTcouples, out=None, stream=None, check_status=True):
        """[n_cells, n_sets, 3, n_grid] float64 CUDA tensor of G-coefficient spectra."""
        torch = _torch()
        pt = as_f64(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
        n_cells = pt.shape[0]
        if out is None:
            out = torch.empty((n_cells, self.n_sets, 3, self.n_grid), dtype=torch.float64,
                              device="cuda")
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous()
        assert out.numel() == self.cells_elems(n_cells)
        sp = _stream_ptr(stream)
        check(lib().sr_gcoeff_cells_dev(self._h, dptr(pt), n_cells, C.c_void_p(out.data_ptr()),
                                        sp))
        if check_status:
            check(lib().sr_lineset_check(self._h, sp))
        return out

    def gcoeff_cells_f32(self, PTcouples, out=None, stream=None):
        """Same as gcoeff_cells but stored as float32 (the compressed LUT,
        spect_main_module.py:1676); the tile kernel rounds in its store, no FP64 copy exists.
        The result is a [n_cells, n_sets, 3, n_grid] view of a table whose rows are padded to a
        multiple of 32 floats (lut_tensor): 128-byte aligned rows for the LOS kernels.  `out` may
        be such a view (or a dim-0 slice of one) or a plain contiguous tensor."""
        torch = _torch()
        pt = as_f64(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
        n_cells = pt.shape[0]
        if out is None:
            out = lut_tensor(n_cells, self.n_sets, self.n_grid)
        rs = lut_row_stride(out)
        assert tuple(out.shape) == (n_cells, self.n_sets, 3, self.n_grid)
        sp = _stream_ptr(stream)
        check(lib().sr_gcoeff_cells_dev_f32_ld(self._h, dptr(pt), n_cells,
                                               C.c_void_p(out.data_ptr()), rs, sp))
        check(lib().sr_lineset_check(self._h, sp))
        return out

    def tile_points(self):
        """Alignment (grid points) of the output windows of gcoeff_cells_window."""
        return int(lib().sr_lineset_tile_points(self._h))

    def gcoeff_cells_window(self, PTcouples, pt0, n_pts, f32=True, out=None, stream=None):
        """The cells on the grid points [pt0, pt0+n_pts) only (a wavenumber slab of the LUT; pt0 a
        multiple of tile_points()): [n_cells, n_sets, 3, n_pts] float32 (row-padded view, see
        lut_tensor) or float64.  Bit-identical to the same points of a build on the whole grid."""
        torch = _torch()
        pt = as_f64(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
        n_cells = pt.shape[0]
        if out is None:
            out = (lut_tensor(n_cells, self.n_sets, n_pts) if f32 else
                   torch.empty((n_cells, self.n_sets, 3, n_pts), dtype=torch.float64, device="cuda"))
        assert tuple(out.shape) == (n_cells, self.n_sets, 3, int(n_pts))
        rs = lut_row_stride(out) if f32 else int(n_pts)
        assert f32 or out.is_contiguous()
        sp = _stream_ptr(stream)
        check(lib().sr_gcoeff_cells_window_dev(self._h, dptr(pt), n_cells, C.c_void_p(out.data_ptr()),
                                               int(bool(f32)), rs, int(pt0), int(n_pts), sp))
        check(lib().sr_lineset_check(self._h, sp))
        return out

    def gcoeff_cells_host(self, PTcouples, out=None):
        """Host-buffer entry point (copies inside): numpy [n_cells, n_sets, 3, n_grid]."""
        pt = as_f64(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
        n_cells = pt.shape[0]
        if out is None:
            out = np.empty((n_cells, self.n_sets, 3, self.n_grid))
        check(lib().sr_gcoeff_cells_host(self._h, dptr(pt), n_cells, dptr(out)))
        return out

    def line_shapes(self, Pres, Temp):
        """Per-line normalised shapes [n_active, 13010] and G coefficients [n_active, 3] (CUDA
        tensors, internal sorted order; see order())."""
        torch = _torch()
        shapes = torch.empty((self.n_active, IMXSIG), dtype=torch.float64, device="cuda")
        g = torch.empty((self.n_active, 3), dtype=torch.float64, device="cuda")
        check(lib().sr_line_shapes_dev(self._h, float(Pres), float(Temp),
                                       C.c_void_p(shapes.data_ptr()), C.c_void_p(g.data_ptr()),
                                       _stream_ptr()))
        return shapes, g


def lut_tensor(n_cells, n_sets, n_grid, zero=False):
    """Float32 LUT storage [n_cells, n_sets, 3, n_grid] as a view of a row-padded CUDA tensor
    (row stride = n_grid rounded up to 32 floats; the padding is zero)."""
    torch = _torch()
    rs = (int(n_grid) + 31) // 32 * 32
    alloc = torch.zeros if zero else torch.empty
    base = alloc((int(n_cells), int(n_sets), 3, rs), dtype=torch.float32, device="cuda")
    if not zero and rs > n_grid:
        base[..., n_grid:].zero_()
    return base[..., :n_grid]


def lut_from_host(arr):
    """Row-padded device LUT (see lut_tensor) from a host array [n_cells, n_sets, 3, n_grid]."""
    torch = _torch()
    arr = np.asarray(arr, dtype=np.float32)
    out = lut_tensor(arr.shape[0], arr.shape[1], arr.shape[3])
    out.copy_(torch.from_numpy(np.ascontiguousarray(arr)))
    return out


def lut_cat(tables):
    """Concatenate LUT tensors along the cell axis, keeping the padded-row layout."""
    torch = _torch()
    n = sum(t.shape[0] for t in tables)
    out = lut_tensor(n, tables[0].shape[1], tables[0].shape[3])
    c = 0
    for t in tables:
        out[c:c + t.shape[0]].copy_(t)
        c += t.shape[0]
    return out


def lut_row_stride(g32):
    """Row stride (floats) of a LUT tensor: a contiguous [n_cells, n_sets, 3, n_grid] tensor or a
    view over padded rows as made by lut_tensor."""
    torch = _torch()
    assert g32.is_cuda and g32.dtype == torch.float32 and g32.dim() == 4 and g32.shape[2] == 3
    n_cells, n_sets, _, n_grid = g32.shape
    rs = g32.stride(2) if n_grid > 1 else n_grid
    ok = g32.stride(3) == 1 and rs >= n_grid and g32.stride(1) == 3 * rs and \
        (n_cells == 1 or g32.stride(0) == n_sets * 3 * rs)
    if not ok:
        raise ValueError("LUT tensor must be [n_cells, n_sets, 3, n_grid] with unit point stride "
                         "and uniformly padded rows (strides %s)" % (g32.stride(),))
    return int(rs)


def lut_padded(g32):
    """The contiguous row-padded tensor behind a LUT view (for collectives, which need
    contiguous memory)."""
    torch = _torch()
    rs = lut_row_stride(g32)
    n_cells, n_sets, _, _ = g32.shape
    return torch.as_strided(g32, (n_cells, n_sets, 3, rs), (n_sets * 3 * rs, 3 * rs, rs, 1),
                            g32.storage_offset())


class Lut(object):
    """Float32 LUT of one isotopologue resident on the device (sr_lut in spectrobot.h).

    g32: CUDA float32 tensor [n_cells, n_sets, 3, n_grid] (kept alive by this object);
    PTcouples: [[P_hPa, T_K], ...]; level_energies: per set (cm-1), None for an LTE isotopologue
    whose lines are not assigned to levels (single set 'all', pop = 1/Q).
    """

    def __init__(self, g32, PTcouples, mol, iso, iso_ratio, level_energies=None, consts=None):
        rs = lut_row_stride(g32)
        self.g32 = g32
        self.pt = as_f64(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
        n_cells, n_sets, three, n_grid = g32.shape
        assert three == 3 and n_cells == self.pt.shape[0]
        self.n_cells, self.n_sets, self.n_grid = n_cells, n_sets, n_grid
        self.mol, self.iso, self.iso_ratio = int(mol), int(iso), float(iso_ratio)
        self.lte_unidentified = level_energies is None
        self.level_energies = None if level_energies is None else as_f64(level_energies)
        if self.level_energies is not None and len(self.level_energies) != n_sets:
            raise ValueError("level_energies must have one entry per set")
        self.consts = consts if consts is not None else _lib.python_consts()
        self._h = C.c_void_p()
        check(lib().sr_lut_create_ld(
            C.c_void_p(g32.data_ptr()), rs, dptr(self.pt), n_cells, n_sets, n_grid,
            None if self.level_energies is None else dptr(self.level_energies),
            self.mol, self.iso, self.iso_ratio, int(self.lte_unidentified),
            C.byref(self.consts), C.byref(self._h)))

    def set_emission_mask(self, mask):
        """Bit s of `mask`: spontaneous emission of set s contributes to the source function of
        the following LOS calls (None: all sets, the default).  Absorption is not affected."""
        m = 0xFFFFFFFFFFFFFFFF if mask is None else int(mask) & 0xFFFFFFFFFFFFFFFF
        check(lib().sr_lut_set_emission_mask(self._h, C.c_ulonglong(m)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().sr_lut_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LosSteps(object):
    """Step tables of a LOS batch (sr_los_steps in spectrobot.h), steps ordered far end ->
    observer.  n_steps [n_los]; temp, pres [n_los, n_steps_max]; column [n_gas, n_los,
    n_steps_max]; tvib [n_gas, n_sets_max, n_los, n_steps_max] or None (LTE)."""

    def __init__(self, n_steps, temp, pres, column, tvib=None):
        self.n_steps = as_i32(n_steps)
        self.temp = as_f64(temp)
        self.pres = as_f64(pres)
        self.column = as_f64(column)
        self.tvib = None if tvib is None else as_f64(tvib)
        self.n_los, self.n_steps_max = self.temp.shape
        if self.column.ndim == 2:
            self.column = self.column[None]
        self.n_gas = self.column.shape[0]
        assert self.pres.shape == self.temp.shape and len(self.n_steps) == self.n_los
        assert self.column.shape == (self.n_gas, self.n_los, self.n_steps_max)
        self.n_sets_max = 0
        if self.tvib is not None:
            assert self.tvib.ndim == 4 and self.tvib.shape[0] == self.n_gas
            assert self.tvib.shape[2:] == (self.n_los, self.n_steps_max)
            self.n_sets_max = self.tvib.shape[1]

    def struct(self):
        st = _lib.sr_los_steps()
        st.n_los, st.n_steps_max, st.n_gas, st.n_sets_max = (self.n_los, self.n_steps_max,
                                                              self.n_gas, self.n_sets_max)
        st.n_steps = iptr(self.n_steps)
        st.temp = dptr(self.temp)
        st.pres = dptr(self.pres)
        st.column = dptr(self.column)
        st.tvib = None if self.tvib is None else dptr(self.tvib)
        return st

    def subset(self, sl):
        return LosSteps(self.n_steps[sl], self.temp[sl], self.pres[sl], self.column[:, sl],
                        None if self.tvib is None else self.tvib[:, :, sl])


def _lut_array(luts):
    arr = (C.c_void_p * len(luts))(*[l._h for l in luts])
    return arr


def los_rt_lut(luts, steps, pt0=0, n_pts=None, i0=None, solo_absorption=False, out=None,
               stream=None, check_status=True):
    """Fused K3a+K3: radiances [n_los, n_pts] (CUDA float64 tensor) from device-resident LUTs."""
    torch = _torch()
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    if out is None:
        out = torch.empty((steps.n_los, n_pts), dtype=torch.float64, device="cuda")
    arr = _lut_array(luts)
    st = steps.struct()
    sp = _stream_ptr(stream)
    check(lib().sr_los_rt_lut_dev(arr, C.byref(st), int(pt0), int(n_pts),
                                  None if i0 is None else C.c_void_p(i0.data_ptr()),
                                  int(bool(solo_absorption)), C.c_void_p(out.data_ptr()), sp))
    if check_status:
        check(lib().sr_los_check(arr, sp))
    return out


def _channels(centres, widths, n_sigma, units):
    """(sr_channels, tensors to keep alive).  units: 'same' (channels in the units of the grid) or
    'nm' (grid in cm-1, channels in nm: the reference's convert_grid_to, spect_classes.py:771-778)."""
    torch = _torch()
    c = centres if torch.is_tensor(centres) else torch.as_tensor(as_f64(centres), device="cuda")
    w = widths if torch.is_tensor(widths) else torch.as_tensor(as_f64(widths), device="cuda")
    assert c.numel() == w.numel() and c.dtype == torch.float64 and w.dtype == torch.float64
    ch = _lib.sr_channels()
    ch.n_chan, ch.n_sigma = c.numel(), float(n_sigma)
    ch.centre_dev, ch.width_dev = c.data_ptr(), w.data_ptr()
    ch.units = {'same': _lib.SR_CHAN_SAME_UNITS, 'nm': _lib.SR_CHAN_NM_FROM_CM1}[units]
    return ch, (c, w)


def los_rt_lut_lowres(luts, steps, grid, centres, widths, pt0=0, n_pts=None, n_sigma=5.0, i0=None,
                      solo_absorption=False, out=None, stream=None, check_status=True,
                      units='same'):
    """K3a+K3 + instrument convolution: low-res spectra [n_los, n_chan] (CUDA float64) of a batch
    of any size; the hi-res radiances only ever exist per LOS block on the device.  grid: CUDA
    float64 tensor with the WHOLE spectral grid of the LUTs; centres/widths: channel definition in
    the units of the grid (array-likes or CUDA tensors)."""
    torch = _torch()
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    assert grid.is_cuda and grid.dtype == torch.float64 and grid.numel() == luts[0].n_grid
    ch, keep = _channels(centres, widths, n_sigma, units)
    if out is None:
        out = torch.empty((steps.n_los, ch.n_chan), dtype=torch.float64, device="cuda")
    arr = _lut_array(luts)
    st = steps.struct()
    sp = _stream_ptr(stream)
    gwin = grid[pt0:pt0 + n_pts]
    check(lib().sr_los_rt_lut_channels_dev(arr, C.byref(st), int(pt0), int(n_pts),
                                           C.c_void_p(gwin.data_ptr()), C.byref(ch),
                                           None if i0 is None else C.c_void_p(i0.data_ptr()),
                                           int(bool(solo_absorption)), C.c_void_p(out.data_ptr()), sp))
    if check_status:
        check(lib().sr_los_check(arr, sp))
    return out


def los_rt_lut_host(luts, steps, pt0=0, n_pts=None, i0=None, solo_absorption=False, out=None):
    """Host-buffer entry point (copies inside): numpy [n_los, n_pts]."""
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    if out is None:
        out = np.empty((steps.n_los, n_pts))
    arr = _lut_array(luts)
    st = steps.struct()
    i0a = None if i0 is None else as_f64(i0)
    check(lib().sr_los_rt_lut_host(arr, C.byref(st), int(pt0), int(n_pts),
                                   None if i0a is None else dptr(i0a),
                                   int(bool(solo_absorption)), dptr(out)))
    return out


def los_tau_src(luts, steps, pt0=0, n_pts=None, stream=None):
    """K3a alone: materialised (tau, S) CUDA tensors [n_los, n_steps_max, n_pts]."""
    torch = _torch()
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    shape = (steps.n_los, steps.n_steps_max, n_pts)
    tau = torch.zeros(shape, dtype=torch.float64, device="cuda")
    src = torch.zeros(shape, dtype=torch.float64, device="cuda")
    arr = _lut_array(luts)
    st = steps.struct()
    sp = _stream_ptr(stream)
    check(lib().sr_los_tau_src_dev(arr, C.byref(st), int(pt0), int(n_pts),
                                   C.c_void_p(tau.data_ptr()), C.c_void_p(src.data_ptr()), sp))
    check(lib().sr_los_check(arr, sp))
    return tau, src


def los_abs_emi(luts, steps, pt0=0, n_pts=None, stream=None):
    """make_abscoeff_LUTS_fast on the device: (abs, emi) CUDA tensors [n_los, n_steps_max, n_pts]
    (coefficients x column x isotopic ratio)."""
    torch = _torch()
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    shape = (steps.n_los, steps.n_steps_max, n_pts)
    a = torch.zeros(shape, dtype=torch.float64, device="cuda")
    e = torch.zeros(shape, dtype=torch.float64, device="cuda")
    arr = _lut_array(luts)
    st = steps.struct()
    sp = _stream_ptr(stream)
    check(lib().sr_los_abs_emi_dev(arr, C.byref(st), int(pt0), int(n_pts),
                                   C.c_void_p(a.data_ptr()), C.c_void_p(e.data_ptr()), sp))
    check(lib().sr_los_check(arr, sp))
    return a, e


def _jac_args(steps, dfrac, gas_in_jac):
    dfrac = as_f64(dfrac)
    assert dfrac.ndim == 3 and dfrac.shape[:2] == (steps.n_los, steps.n_steps_max)
    gij = None if gas_in_jac is None else as_i32(gas_in_jac)
    assert gij is None or len(gij) == steps.n_gas
    return dfrac, gij


def los_rt_lut_jac(luts, steps, dfrac, gas_in_jac=None, pt0=0, n_pts=None, i0=None,
                   solo_absorption=False, stream=None, check_status=True):
    """Fused K3a+K3 with analytic Jacobians: (rad [n_los, n_pts], jac [n_los, n_par, n_pts]) CUDA
    float64 tensors.  dfrac [n_los, n_steps_max, n_par] = (d column / d parameter) / column of the
    retrieved gas per step; gas_in_jac [n_gas]: which LUTs belong to that gas (None = all)."""
    torch = _torch()
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    dfrac, gij = _jac_args(steps, dfrac, gas_in_jac)
    n_par = dfrac.shape[2]
    rad = torch.empty((steps.n_los, n_pts), dtype=torch.float64, device="cuda")
    jac = torch.empty((steps.n_los, n_par, n_pts), dtype=torch.float64, device="cuda")
    arr = _lut_array(luts)
    st = steps.struct()
    sp = _stream_ptr(stream)
    check(lib().sr_los_rt_lut_jac_dev(arr, C.byref(st), n_par, None if gij is None else iptr(gij),
                                      dptr(dfrac), int(pt0), int(n_pts),
                                      None if i0 is None else C.c_void_p(i0.data_ptr()),
                                      int(bool(solo_absorption)), C.c_void_p(rad.data_ptr()),
                                      C.c_void_p(jac.data_ptr()), sp))
    if check_status:
        check(lib().sr_los_check(arr, sp))
    return rad, jac


def los_rt_lut_jac_lowres(luts, steps, dfrac, grid, centres, widths, gas_in_jac=None, pt0=0,
                          n_pts=None, n_sigma=5.0, i0=None, solo_absorption=False, stream=None,
                          check_status=True, units='same'):
    """Same, reduced to the instrument channels on the device: (low [n_los, n_chan],
    jac_low [n_los, n_par, n_chan]); see los_rt_lut_lowres for grid / centres / widths."""
    torch = _torch()
    if n_pts is None:
        n_pts = luts[0].n_grid - pt0
    dfrac, gij = _jac_args(steps, dfrac, gas_in_jac)
    n_par = dfrac.shape[2]
    assert grid.is_cuda and grid.dtype == torch.float64 and grid.numel() == luts[0].n_grid
    ch, keep = _channels(centres, widths, n_sigma, units)
    low = torch.empty((steps.n_los, ch.n_chan), dtype=torch.float64, device="cuda")
    jlow = torch.empty((steps.n_los, n_par, ch.n_chan), dtype=torch.float64, device="cuda")
    arr = _lut_array(luts)
    st = steps.struct()
    sp = _stream_ptr(stream)
    gwin = grid[pt0:pt0 + n_pts]
    check(lib().sr_los_rt_lut_jac_channels_dev(
        arr, C.byref(st), n_par, None if gij is None else iptr(gij), dptr(dfrac), int(pt0),
        int(n_pts), C.c_void_p(gwin.data_ptr()), C.byref(ch),
        None if i0 is None else C.c_void_p(i0.data_ptr()), int(bool(solo_absorption)),
        C.c_void_p(low.data_ptr()), C.c_void_p(jlow.data_ptr()), sp))
    if check_status:
        check(lib().sr_los_check(arr, sp))
    return low, jlow


def los_rt_layers_jac(tau, emi, dfrac, n_steps, tau_g=None, emi_g=None, i0=None,
                      solo_absorption=False, stream=None):
    """K3 + Jacobians on materialised layers (tau, emi from los_abs_emi; tau_g, emi_g the retrieved
    gas alone or None): CUDA tensors -> (rad [n_los, n_pts], jac [n_los, n_par, n_pts])."""
    torch = _torch()
    n_los, n_steps_max, n_pts = tau.shape
    assert emi.shape == tau.shape and tau.is_contiguous() and emi.is_contiguous()
    assert dfrac.is_cuda and dfrac.dtype == torch.float64 and dfrac.is_contiguous()
    assert dfrac.shape[:2] == (n_los, n_steps_max)
    assert (tau_g is None) == (emi_g is None)
    n_par = dfrac.shape[2]
    rad = torch.empty((n_los, n_pts), dtype=torch.float64, device="cuda")
    jac = torch.empty((n_los, n_par, n_pts), dtype=torch.float64, device="cuda")
    vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())   # noqa: E731
    check(lib().sr_los_rt_layers_jac_dev(vp(tau), vp(emi), vp(tau_g), vp(emi_g), vp(dfrac), n_par,
                                         vp(n_steps), n_los, n_steps_max, n_pts, vp(i0),
                                         int(bool(solo_absorption)), vp(rad), vp(jac),
                                         _stream_ptr(stream)))
    return rad, jac


class Atmosphere(object):
    """Atmosphere tables for the device step builder (sr_atmosphere in spectrobot.h).

    z [n_z] km; temp, pres [n_band, n_z] (or [n_z]); vmr [n_gas, n_band, n_z], one entry per LUT of
    the later LOS call; tvib [n_gas, n_sets_max, n_band, n_z] or, with a solar-zenith-angle axis
    sza_nodes [n_sza] (degrees, ascending), [n_gas, n_sets_max, n_band, n_sza, n_z], or None;
    tvib_on [n_gas, n_sets_max] (1 own profile, 0 T_vib = step temperature, -1 no such level);
    lat_edges [n_band+1] degrees."""

    def __init__(self, z, temp, pres, vmr, tvib=None, tvib_on=None, lat_edges=None,
                 radius_km=2575.0, top_km=1500.0, sza_nodes=None):
        self.z = as_f64(z)
        nz = len(self.z)
        self.temp = as_f64(np.asarray(temp, dtype=float).reshape(-1, nz))
        self.pres = as_f64(np.asarray(pres, dtype=float).reshape(-1, nz))
        self.n_band = self.temp.shape[0]
        vmr = np.asarray(vmr, dtype=float)
        self.vmr = as_f64(vmr.reshape(-1, self.n_band, nz))
        self.n_gas = self.vmr.shape[0]
        self.tvib = None
        self.n_sets_max = 0
        self.tvib_on = None
        if tvib_on is not None:
            self.tvib_on = as_i32(np.asarray(tvib_on).reshape(self.n_gas, -1))
            self.n_sets_max = self.tvib_on.shape[1]
        self.sza_nodes = None if sza_nodes is None or len(sza_nodes) <= 1 else as_f64(sza_nodes)
        self.n_sza = 1 if self.sza_nodes is None else len(self.sza_nodes)
        if tvib is not None:
            self.tvib = as_f64(np.asarray(tvib, dtype=float).reshape(self.n_gas, -1, self.n_band,
                                                                     self.n_sza, nz))
            if self.tvib_on is None:
                self.n_sets_max = self.tvib.shape[1]
                self.tvib_on = as_i32(np.ones((self.n_gas, self.n_sets_max)))
            assert self.tvib.shape[1] == self.n_sets_max
        self.lat_edges = None if lat_edges is None or self.n_band == 1 else as_f64(lat_edges)
        assert self.n_band == 1 or len(self.lat_edges) == self.n_band + 1
        assert self.pres.shape == self.temp.shape
        self.radius_km, self.top_km = float(radius_km), float(top_km)

    def struct(self):
        st = _lib.sr_atmosphere()
        st.n_band, st.n_z, st.n_gas, st.n_sets_max = self.n_band, len(self.z), self.n_gas, self.n_sets_max
        st.lat_edges = None if self.lat_edges is None else dptr(self.lat_edges)
        st.z, st.temp, st.pres, st.vmr = dptr(self.z), dptr(self.temp), dptr(self.pres), dptr(self.vmr)
        st.tvib = None if self.tvib is None else dptr(self.tvib)
        st.tvib_on = None if self.tvib_on is None else iptr(self.tvib_on)
        st.radius_km, st.top_km = self.radius_km, self.top_km
        st.n_sza = self.n_sza
        st.sza_nodes = None if self.sza_nodes is None else dptr(self.sza_nodes)
        return st


def los_steps_build(atm, origins, directions, delta_x=5.0, max_T_variation=5.0,
                    max_Plog_variation=1.0, masks=None, jac_gas=-1, n_steps_max=64, sun=None,
                    sza_fixed=None, max_opt_depth=None, sigma_peak=None, photon_order=False):
    """LOS geometry + radtran steps of a whole batch on the device (sr_los_steps_build_rays): rays
    from origins [n_los, 3] (km, planetocentric Cartesian) along unit directions [n_los, 3].
    sun [n_los, 3] (or [3]): direction of the Sun per LOS, for SZA-dependent vibrational
    temperatures; sza_fixed [n_los]: one SZA per LOS instead (use_tangent_sza); max_opt_depth with
    sigma_peak [n_gas]: optical-depth limit of a step; photon_order: LOS_order='photon'.  Returns
    (LosSteps, dfrac) - dfrac [n_los, n_steps_max, n_par] for the parameter masks [n_par, n_z] or
    [n_par, n_band, n_z] of gas entry jac_gas, or None.  The table width grows until every LOS fits."""
    org, dr = as_f64(origins).reshape(-1, 3), as_f64(directions).reshape(-1, 3)
    n_los = org.shape[0]
    assert dr.shape == org.shape
    rays = _lib.sr_los_rays()
    rays.n_los, rays.origin, rays.direction = n_los, dptr(org), dptr(dr)
    if sun is not None:
        sun = np.asarray(sun, dtype=float)
        sun = as_f64(np.broadcast_to(sun / np.linalg.norm(sun, axis=-1, keepdims=True), (n_los, 3)))
        rays.sun = dptr(sun)
    if sza_fixed is not None:
        sza_fixed = as_f64(np.broadcast_to(np.asarray(sza_fixed, dtype=float), (n_los,)))
        rays.sza_fixed = dptr(sza_fixed)
    opt = _lib.sr_steps_opt()
    opt.delta_x_km, opt.max_T_variation = float(delta_x), float(max_T_variation)
    opt.max_Plog_variation = float(max_Plog_variation)
    opt.max_opt_depth = 0.0 if max_opt_depth is None else float(max_opt_depth)
    opt.photon_order = int(bool(photon_order))
    if sigma_peak is not None:
        sigma_peak = as_f64(sigma_peak)
        assert len(sigma_peak) == atm.n_gas
        opt.sigma_peak = dptr(sigma_peak)
    mk = None
    if masks is not None:   # [n_par, n_z] (same in every latitude box) or [n_par, n_band, n_z]
        mk = np.asarray(masks, dtype=float)
        if mk.ndim == 2:
            mk = np.repeat(mk[:, None, :], atm.n_band, axis=1)
        assert mk.shape[1:] == (atm.n_band, len(atm.z))
        mk = as_f64(mk)
    n_par = 0 if mk is None else mk.shape[0]
    st = atm.struct()
    while True:
        n_steps = np.zeros(n_los, dtype=np.int32)
        temp = np.empty((n_los, n_steps_max))
        pres = np.empty((n_los, n_steps_max))
        col = np.empty((atm.n_gas, n_los, n_steps_max))
        tvib = np.empty((atm.n_gas, atm.n_sets_max, n_los, n_steps_max)) if atm.n_sets_max else None
        dfrac = np.empty((n_los, n_steps_max, n_par)) if n_par else None
        need = C.c_int(0)
        rc = lib().sr_los_steps_build_rays(C.byref(st), C.byref(rays), C.byref(opt), n_par,
                                           None if mk is None else dptr(mk), int(jac_gas),
                                           int(n_steps_max), iptr(n_steps), dptr(temp), dptr(pres),
                                           dptr(col), None if tvib is None else dptr(tvib),
                                           None if dfrac is None else dptr(dfrac), C.byref(need))
        if rc == _lib.SR_ERR_LIMIT and need.value > n_steps_max:
            n_steps_max = need.value
            continue
        check(rc)
        break
    # trim the tables to the longest LOS: the LOS kernels size their layer scratch by the width
    w = max(int(n_steps.max()), 1)
    if w < n_steps_max:
        temp, pres, col = temp[:, :w], pres[:, :w], col[:, :, :w]
        tvib = None if tvib is None else tvib[:, :, :, :w]
        dfrac = None if dfrac is None else dfrac[:, :w]
    return LosSteps(n_steps, temp, pres, col, tvib), dfrac


def partition_sum(mol, iso, temp=296.0):
    """CalcPartitionSum (spect_classes.py:1692-1710) from the library's TIPS tables."""
    q = C.c_double()
    check(lib().sr_partition_sum(int(mol), int(iso), float(temp), C.byref(q)))
    return q.value


def los_rt_layers(tau, src, n_steps, i0=None, solo_absorption=False, out=None, stream=None):
    """K3 alone on materialised layers: tau, src CUDA float64 [n_los, n_steps_max, n_pts];
    n_steps CUDA int32 [n_los] -> radiances [n_los, n_pts]."""
    torch = _torch()
    n_los, n_steps_max, n_pts = tau.shape
    assert src.shape == tau.shape and tau.is_contiguous() and src.is_contiguous()
    assert n_steps.is_cuda and n_steps.dtype == torch.int32
    if out is None:
        out = torch.empty((n_los, n_pts), dtype=torch.float64, device="cuda")
    check(lib().sr_los_rt_layers_dev(C.c_void_p(tau.data_ptr()), C.c_void_p(src.data_ptr()),
                                     C.c_void_p(n_steps.data_ptr()), n_los, n_steps_max, n_pts,
                                     None if i0 is None else C.c_void_p(i0.data_ptr()),
                                     int(bool(solo_absorption)), C.c_void_p(out.data_ptr()),
                                     _stream_ptr(stream)))
    return out


def lut_weights(PTcouples, Pres, Temp):
    """LutSet.calculate's cell choice (spect_main_module.py:997-1066): (cells[4], weights[4])."""
    pt = as_f64(np.asarray(PTcouples, dtype=float).reshape(-1, 2))
    cell = np.empty(4, dtype=np.int32)
    w = np.empty(4)
    check(lib().sr_lut_weights(dptr(pt), pt.shape[0], float(Pres), float(Temp), iptr(cell),
                               dptr(w)))
    return cell, w


def convolve_lowres(grid, spec, centres, widths, n_sigma=5.0, out=None, stream=None, units='same'):
    """Instrument convolution on the device (spect_classes.py:883-918): grid [n_pts] and spec
    [n_spec, n_pts] CUDA float64 tensors, channel centres/widths (array-likes or CUDA tensors) ->
    CUDA float64 [n_spec, n_chan]."""
    torch = _torch()
    if spec.dim() == 1:
        spec = spec[None]
    assert spec.is_cuda and spec.dtype == torch.float64 and spec.is_contiguous()
    assert grid.is_cuda and grid.dtype == torch.float64 and grid.numel() == spec.shape[1]
    ch, keep = _channels(centres, widths, n_sigma, units)
    if out is None:
        out = torch.empty((spec.shape[0], ch.n_chan), dtype=torch.float64, device="cuda")
    check(lib().sr_convolve_channels_dev(C.c_void_p(grid.data_ptr()), spec.shape[1],
                                         C.c_void_p(spec.data_ptr()), spec.shape[0], C.byref(ch),
                                         C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
    return out


def convolve_lowres_host(grid, spec, centres, widths, n_sigma=5.0, units='same'):
    """Host-buffer form of convolve_lowres: numpy in, numpy [n_spec, n_chan] out."""
    grid = as_f64(grid)
    spec = np.atleast_2d(as_f64(spec))
    c, w = as_f64(centres), as_f64(widths)
    out = np.empty((spec.shape[0], len(c)))
    u = {'same': _lib.SR_CHAN_SAME_UNITS, 'nm': _lib.SR_CHAN_NM_FROM_CM1}[units]
    check(lib().sr_convolve_channels_host(dptr(grid), len(grid), dptr(spec), spec.shape[0], dptr(c),
                                          dptr(w), len(c), float(n_sigma), u, dptr(out)))
    return out


PROF_KINDS = {"los_mma": _lib.SR_PROF_LOS_MMA, "los_layers": _lib.SR_PROF_LOS_LAYERS,
              "conv": _lib.SR_PROF_CONV, "voigt_tile": _lib.SR_PROF_VOIGT_TILE,
              "voigt_core": _lib.SR_PROF_VOIGT_CORE, "los_fused": _lib.SR_PROF_LOS_FUSED}


def prof_enable(on=True):
    """Per-kernel CUDA-event timing inside the library (sr_prof_enable); clears earlier records."""
    check(lib().sr_prof_enable(int(bool(on))))


def prof_summary():
    """{kernel: (launches, total ms, total algorithmic work)} since prof_enable (synchronises)."""
    out = {}
    for name, kind in PROF_KINDS.items():
        n, ms, work = C.c_longlong(), C.c_double(), C.c_double()
        check(lib().sr_prof_summary(kind, C.byref(n), C.byref(ms), C.byref(work)))
        if n.value:
            out[name] = (int(n.value), float(ms.value), float(work.value))
    return out


def fp64_peak(iters=20000):
    v = C.c_double()
    check(lib().sr_fp64_peak(int(iters), C.byref(v)))
    return v.value
