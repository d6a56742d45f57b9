"""`spect_main_module` for the hot path, Python 3, backed by libspectrobot.so.

Keeps the reference's names and call signatures for the forward-model chain:

    prepare_spe_grid, calc_PT_couples_atmosphere, LutSet, LookUpTable, makeLUT_nonLTE_Gcoeffs,
    check_and_build_allluts, make_abscoeff_LUTS_fast, radtrans

Differences in HOW, not in WHAT: LookUpTable.make builds every (P,T) cell x level x ctype with one
batched GPU call and keeps the LUT resident on the device as float32 (the reference's compressed
LUT, smm:1676); radtrans runs all lines of sight of a wavenumber chunk in one launch instead of one
forked process per LOS, and reduces each LOS block to the instrument channels on the device;
FOV_integr_1D integrates the three LOS of a pixel in closed form.  LUT files can be written and
read in the reference's per-level pickle stream (LookUpTable.export_levels / import_levels).  The
retrieval algebra runs on the host under the reference's names (tiny dense matrices); observation
readers whose formats live in the missing spect_base_module are not provided.
"""
import copy
import math as mt
import os
import pickle
import time
import types

import numpy as np

from . import engine
from . import spect_base_module as sbm
from . import spect_classes as spcl

n_threads = 8
CTYPES = spcl.CTYPES


def date_stamp():
    t = time.localtime()
    return '_{:d}-{}-{:d}'.format(t.tm_mday, time.strftime('%b', t), t.tm_year)


def find_free_name(filename, maxnum=1000, split_at='.'):
    """`filename`, or the first `name_NNN.ext` that does not exist yet (smm:34-51); the counter is
    inserted in front of the first `split_at` and has 2, 3 or 5 digits for maxnum <= 100, <= 1000,
    larger."""
    form = '_{:02d}' if maxnum <= 100 else ('_{:03d}' if maxnum <= 1000 else '_{:05d}')
    orig, ind, num = filename, filename.index(split_at), 1
    while os.path.isfile(filename):
        filename = orig[:ind] + form.format(num) + orig[ind:]
        num += 1
        if num > maxnum:
            raise ValueError('Check filenames! More than {} with the same name'.format(maxnum))
    return filename


def equiv(num1, num2, thres=1e-8):
    """Two floats agree to a relative `thres` of the first (smm:53-66); 0 only equals 0."""
    if num1 == 0:
        return num2 == 0
    return abs((num1 - num2) / num1) < thres


def listbands(isomol, lines):
    """Prints `upper -> lower : n lines` for every pair of levels of isomol that has lines
    (smm:117-130) and returns {(upper, lower): n}."""
    mine = [lin for lin in lines if lin.Mol == isomol.mol and lin.Iso == isomol.iso]
    out = dict()
    for lev in isomol.levels:
        up = getattr(isomol, lev)
        for lev2 in isomol.levels:
            lo = getattr(isomol, lev2)
            n = len([lin for lin in mine if lo.equiv(lin.Lo_lev_str) and up.equiv(lin.Up_lev_str)])
            if n > 0:
                print('{} -> {} : {} lines'.format(lev, lev2, n))
                out[(lev, lev2)] = n
    return out


def lut_name(mol, iso, LTE):
    """LUT_molMM_isoI_{LTE|nonLTE} (smm:659-680)."""
    return 'LUT_mol{:02d}_iso{:1d}_{}'.format(mol, iso, 'LTE' if LTE else 'nonLTE')


def lut_name_split(mol, iso, LTE, split):
    """LUT_csplitNN_molMM_isoI_{LTE|nonLTE} (smm:666-672)."""
    return 'LUT_csplit{:02d}_mol{:02d}_iso{:1d}_{}'.format(split, mol, iso, 'LTE' if LTE else 'nonLTE')


def lut_name_wsplits(mol, iso, LTE, n_split):
    """LUT_NNsplits_molMM_isoI_{LTE|nonLTE} (smm:674-680)."""
    return 'LUT_{:02d}splits_mol{:02d}_iso{:1d}_{}'.format(n_split, mol, iso, 'LTE' if LTE else 'nonLTE')


def prepare_spe_grid(wn_range, sp_step=5.e-4, units='cm_1'):
    """Zero spectrum on np.arange(w0, w1 + step/2, step) (smm:1262-1272)."""
    grid = spcl.SpectralGrid(np.arange(wn_range[0], wn_range[1] + sp_step / 2, sp_step,
                                       dtype=float), units=units)
    return spcl.SpectralObject(np.zeros(len(grid.grid)), grid, units=units)


# ---------------------------------------------------------------------------------------------
# (P,T) cells of the LUT
# ---------------------------------------------------------------------------------------------
def calc_PT_couples_atmosphere(lines, molecs, atmosphere, pres_step_log=0.4, temp_step=5.0,
                               max_pres=None, thres=0.01, add_lowpres=True):
    """The [P_hPa, T_K] cells a LUT needs for `atmosphere` (smm:1746-1844).

    ln P on multiples of pres_step_log between the lowest atmospheric pressure and max_pres; for
    every pressure the temperature range found within +-1 pressure level, widened by one
    temp_step on both sides and rounded to the temp_step ladder; cells where the Lorentz width of
    the broadest line is below thres x Doppler width collapse onto two pressures (the largest such
    pressure and, with add_lowpres, the lowest ladder pressure)."""
    atm_p = np.asarray(atmosphere.pres, dtype=float)
    atm_t = np.asarray(atmosphere.temp, dtype=float)
    top = mt.log(np.max(atm_p)) if max_pres is None else mt.log(max_pres)
    n_hi = mt.ceil(top / pres_step_log)
    n_lo = mt.floor(mt.log(np.min(atm_p)) / pres_step_log)
    ln_lo, ln_hi = n_lo * pres_step_log, n_hi * pres_step_log
    pressures = np.exp(ln_lo + np.arange(0, (ln_hi - ln_lo) + 0.5 * pres_step_log, pres_step_log))

    def t_range(mask):
        t = atm_t[mask]
        return [np.min(t), np.max(t)]

    temps = [t_range(atm_p <= pressures[1])]
    for p0, p2 in zip(pressures[:-2], pressures[2:]):
        temps.append(t_range((atm_p >= p0) & (atm_p <= p2)))
    temps.append(t_range((atm_p >= pressures[-2]) & (atm_p <= pressures[-1])))

    couples = []
    for pres, (t_min, t_max) in zip(pressures, temps):
        t_0 = (np.floor(t_min / temp_step) - 1) * temp_step
        t_1 = (np.ceil(t_max / temp_step) + 1) * temp_step
        couples += [[pres, temp] for temp in np.arange(t_0, t_1 + 0.5 * temp_step, temp_step)]

    if isinstance(molecs, (sbm.Molec, sbm.IsoMolec)):
        molecs = [molecs]
    molecs = list(molecs)        # dict.values() of a Python-3 caller (planet.gases.values())
    mms = []
    for mol in molecs:
        if isinstance(mol, sbm.Molec):
            mms += [getattr(mol, isom).MM for isom in mol.all_iso]
        elif isinstance(mol, sbm.IsoMolec):
            mms.append(mol.MM)
    broadest = lines[int(np.argmax(np.array([lin.Air_broad for lin in lines])))]

    kept, doppler_temps, pres_0 = [], [], 1.e-8
    for pres, temp in couples:
        dw, lw, _ = broadest.CheckWidths(temp, pres, min(mms))
        if lw < thres * dw:
            pres_0 = max(pres_0, pres)
            if temp not in doppler_temps:
                doppler_temps.append(temp)
        else:
            kept.append([pres, temp])
    for temp in doppler_temps:
        kept.insert(0, [pres_0, temp])
    if add_lowpres:
        for temp in doppler_temps:
            kept.insert(0, [mt.exp(ln_lo), temp])
    return kept


# ---------------------------------------------------------------------------------------------
# LUT classes
# ---------------------------------------------------------------------------------------------
class LutSet(object):
    """All (P,T) cells of one vibrational level (or of 'all' lines for an LTE isotopologue), three
    ctypes each (smm:841-1176).  The numbers live in the parent LookUpTable's device tensor;
    `sets` materialises host SpectralGcoeff objects on demand (calculate())."""

    def __init__(self, mol, iso, MM, level=None, filename=None):
        self.mol, self.iso, self.MM = mol, iso, MM
        self.level = copy.deepcopy(level)
        self.unidentified_lines = level is None
        self.filename = filename
        self.filenames = [filename]
        self.sets = []
        self.spectral_grid = None
        self.PTcouples = None
        self.temp_file = None  # the open per-level stream (prepare_read / prepare_export)
        self._table = None     # (LookUpTable, set index) once built on the device

    def find(self, Pres, Temp):
        """Index of [Pres, Temp] in PTcouples (smm:985-995)."""
        if [Pres, Temp] not in self.PTcouples:
            raise ValueError('{} couple not found!'.format([Pres, Temp]))
        return self.PTcouples.index([Pres, Temp])

    # -- the reference's per-level pickle stream (smm:864-979, 1170-1176): PTcouples header, then
    # one {ctype: SpectralGcoeff without grid} per cell.  Host side only; LookUpTable.import_levels
    # / export_levels move whole tables between these files and the resident device tensor.
    def add_file(self, filename, PTcouples):
        """Another file of the same set, written at a different time (smm:864-870)."""
        self.filenames.append(filename)
        self.PTcouples += PTcouples

    def _need_filename(self):
        if self.filename is None:
            raise ValueError('ERROR!: NO filename set for LutSet.')

    def prepare_read(self):
        """Opens the stream and returns its PTcouples header (smm:872-878)."""
        self._need_filename()
        self.temp_file = open(self.filename, 'rb')
        return _RefUnpickler(self.temp_file, encoding='latin1').load()

    def prepare_export(self, PTcouples, spectral_grid):
        """Opens the stream for writing and dumps PTcouples on top (smm:880-892)."""
        self._need_filename()
        self.temp_file = open(self.filename, 'wb')
        self.PTcouples = copy.deepcopy(PTcouples)
        self.spectral_grid = spectral_grid
        pickle.dump(PTcouples, self.temp_file, protocol=-1)

    def finalize_IO(self):
        self.temp_file.close()
        self.temp_file = None

    def _restore(self, set_, spectral_grid):
        for co in set_.values():
            if co is None:
                continue
            co.double_precision()
            grid = self.spectral_grid if self.spectral_grid is not None else spectral_grid
            if grid is None:
                raise ValueError('No spectral grid given.')
            co.restore_grid(grid)
        return set_

    def load_from_file(self, load_just_PT=False, spectral_grid=None):
        """Header and, unless load_just_PT, every cell of self.filename into self.sets in double
        precision with the grid restored (smm:899-921)."""
        with open(self.filename, 'rb') as f:
            self.PTcouples = _RefUnpickler(f, encoding='latin1').load()
            if load_just_PT:
                return
            for _ in self.PTcouples:
                self.sets.append(self._restore(_RefUnpickler(f, encoding='latin1').load(), spectral_grid))

    def load_from_files(self, load_just_PT=False, spectral_grid=None, cartLUTs=None):
        """load_from_file over all of self.filenames (smm:924-959); a file that is not found where
        it was written is looked for under cartLUTs."""
        self.PTcouples = []
        for filename in self.filenames:
            if not os.path.isfile(filename) and cartLUTs is not None:
                filename = os.path.join(cartLUTs, os.path.basename(filename))
            with open(filename, 'rb') as f:
                pts = _RefUnpickler(f, encoding='latin1').load()
                self.PTcouples += pts
                if load_just_PT:
                    continue
                for _ in pts:
                    self.sets.append(self._restore(_RefUnpickler(f, encoding='latin1').load(),
                                                   spectral_grid))

    def load_singlePT_from_file(self, spectral_grid=None):
        """The next cell of the open stream (smm:961-979); opens it (and skips the header) first
        when needed."""
        if getattr(self, 'temp_file', None) is None:   # (objects unpickled from older files)
            self.prepare_read()
        return self._restore(_RefUnpickler(self.temp_file, encoding='latin1').load(), spectral_grid)

    def add_dump(self, set_):
        pickle.dump(set_, self.temp_file, protocol=-1)

    def export(self, filename):
        """The whole object as one pickle (smm:1170-1172)."""
        state = copy.copy(self)
        state._table = None
        state.temp_file = None
        with open(filename, 'wb') as f:
            pickle.dump(state, f, protocol=-1)

    def make(self, spectral_grid, lines, PTcouples, control=True):
        """Every cell of PTcouples for this level, kept in memory (the reference's older
        whole-set builder, smm:1069-1119, without its per-cell progress file): one batched GPU
        build of this level's three ctypes."""
        lines = [lin for lin in lines if lin.Mol == self.mol and lin.Iso == self.iso]
        self.spectral_grid = copy.deepcopy(spectral_grid)
        self.PTcouples = [list(map(float, pt)) for pt in PTcouples]
        lev_str = '' if self.unidentified_lines else self.level.minimal_level_string()
        grid = self.spectral_grid.grid
        tab = spcl.line_table(lines, None)
        if not self.unidentified_lines:
            for i, lin in enumerate(lines):
                u, l = lin.minimal_level_string_up() == lev_str, lin.minimal_level_string_lo() == lev_str
                # set 0 = this level; lines that do not touch it are dropped (-1)
                tab["up_set"][i] = 0 if u else (1 if l else -1)
                tab["lo_set"][i] = 0 if l else (1 if u else -1)
                tab["e_vib_up"][i] = self.level.energy if u else 0.0
                tab["e_vib_lo"][i] = self.level.energy if l else 0.0
        n_sets = 1 if self.unidentified_lines else 2
        ls = engine.LineSet(tab, grid, self.MM, n_sets)
        G = ls.gcoeff_cells(self.PTcouples).cpu().numpy()[:, 0]          # [n_cells, 3, n_grid]
        ls.close()
        self.sets = []
        for c, (P, T) in enumerate(self.PTcouples):
            self.sets.append(dict((ct, spcl.SpectralGcoeff(
                ct, self.spectral_grid, self.mol, self.iso, self.MM, lev_str,
                unidentified_lines=self.unidentified_lines, spectrum=G[c, k], Pres=P, Temp=T))
                for k, ct in enumerate(CTYPES)))
        return self.sets

    def _host_sets(self):
        if not self.sets and getattr(self, '_table', None) is not None:
            lut, s = self._table
            g = lut.g32[:, s].cpu().numpy()
            lev_str = '' if self.unidentified_lines else self.level.minimal_level_string()
            for c, (P, T) in enumerate(self.PTcouples):
                d = dict()
                for k, ctype in enumerate(CTYPES):
                    # all-zero spectra are stored as None by split_and_compress_LUTS (smm:1674-1679)
                    d[ctype] = None if not np.any(g[c, k]) else spcl.SpectralGcoeff(
                        ctype, self.spectral_grid, self.mol, self.iso, self.MM, lev_str,
                        unidentified_lines=self.unidentified_lines,
                        spectrum=g[c, k].astype(np.float64), Pres=P, Temp=T)
                self.sets.append(d)
        return self.sets

    def free_memory(self):
        self.sets = []

    def calculate(self, Pres, Temp):
        """{ctype: SpectralGcoeff} interpolated to (Pres, Temp) with the reference's rule
        (smm:997-1066): nearest and second-nearest node in P and in T, linear in P then in T; at or
        below the lowest pressure node only T is interpolated; above the highest: ValueError."""
        sets = self._host_sets()
        Ps = np.unique(np.array([PT[0] for PT in self.PTcouples]))
        Ts = np.unique(np.array([PT[1] for PT in self.PTcouples]))
        t_near = np.argsort(np.abs(Ts - Temp), kind='stable')
        T1, T2 = Ts[np.argmin(np.abs(Ts - Temp))], Ts[t_near[1]]
        out = dict()
        if Pres <= np.min(Ps):
            a, b = sets[self.find(np.min(Ps), T1)], sets[self.find(np.min(Ps), T2)]
            for ctype in CTYPES:
                out[ctype] = None if a[ctype] is None or b[ctype] is None else \
                    a[ctype].interpolate(b[ctype], Temp=Temp)
        elif Pres <= np.max(Ps):
            P1 = Ps[np.argmin(np.abs(Ps - Pres))]
            P2 = Ps[np.argsort(np.abs(Ps - Pres), kind='stable')[1]]
            c11, c12 = sets[self.find(P1, T1)], sets[self.find(P1, T2)]
            c21, c22 = sets[self.find(P2, T1)], sets[self.find(P2, T2)]
            for ctype in CTYPES:
                if any(c[ctype] is None for c in (c11, c12, c21, c22)):
                    out[ctype] = None
                    continue
                lo = c11[ctype].interpolate(c21[ctype], Pres=Pres)
                hi = c12[ctype].interpolate(c22[ctype], Pres=Pres)
                out[ctype] = lo.interpolate(hi, Temp=Temp)
        else:
            raise ValueError('Extrapolating in P')
        return out

    def add_PT(self, spectral_grid, lines, Pres, Temp, keep_memory=False, control=True,
               n_threads=n_threads):
        """One (P,T) cell from lines that already carry shapes and G coefficients
        (calc_shapes_lines), through BuildCoeff -> sum_all_lines (smm:1122-1168).  The batched
        builder is LookUpTable.make; this method is the reference-shaped single-cell form."""
        if self.spectral_grid is None:
            self.spectral_grid = copy.deepcopy(spectral_grid)
        if self.PTcouples is None:
            self.PTcouples = []
        lev_str = '' if self.unidentified_lines else self.level.minimal_level_string()
        set_ = dict()
        for ctype in CTYPES:
            co = spcl.SpectralGcoeff(ctype, spectral_grid, self.mol, self.iso, self.MM, lev_str,
                                     unidentified_lines=self.unidentified_lines)
            co.BuildCoeff(lines, Temp, Pres, preCalc_shapes=True, n_threads=n_threads)
            set_[ctype] = co
        if getattr(self, 'temp_file', None) is not None:
            # stream opened by prepare_export: the cell goes to the file without its grid
            # (smm:1142-1161); the header already lists the cell
            dumped = dict()
            for ctype, co in set_.items():
                co = copy.copy(co)
                co.erase_grid()
                dumped[ctype] = co
            pickle.dump(dumped, self.temp_file, protocol=-1)
            if keep_memory:
                self.sets.append(set_)
        elif keep_memory:
            self.sets.append(set_)
            self.PTcouples.append([Pres, Temp])
        return set_


class LookUpTable(object):
    """Look-up table of one isotopologue (smm:682-838): one LutSet per vibrational level for a
    non-LTE isotopologue, a single set 'all' for an LTE one."""

    def __init__(self, isomolec, wn_range, LTE):
        self.tag = lut_name(isomolec.mol, isomolec.iso, LTE)
        self.wn_range = copy.deepcopy(wn_range)
        self.mol, self.iso, self.MM = isomolec.mol, isomolec.iso, isomolec.MM
        self.isomolec = copy.deepcopy(isomolec)
        self.sets = dict()
        self.PTcouples = []
        self.LTE = LTE
        self.g32 = None           # CUDA float32 [n_cells][n_sets][3][n_grid]
        self._dev = None

    def set_names(self):
        return list(self.isomolec.levels) if not self.LTE else ['all']

    def make(self, spectral_grid, lines, PTcouples, export_levels=True, cartLUTs=None,
             control=True, n_threads=n_threads, cells=None):
        """Builds every cell of the LUT on the GPU (smm:718-788).  `cells` optionally restricts
        the build to a subset of PTcouples indices (multi-GPU sharding: parallel.shard_cells);
        the other cells are left zero until parallel.gather_lut fills them."""
        import torch
        self.PTcouples = [list(map(float, pt)) for pt in PTcouples]
        self.spectral_grid = copy.deepcopy(spectral_grid)
        lines = [lin for lin in lines if lin.Mol == self.mol and lin.Iso == self.iso]
        tab = spcl.line_table(lines, None if self.LTE else self.isomolec)
        names = self.set_names()
        n_sets = tab["n_sets"]
        if n_sets != len(names):
            raise ValueError('isotopologue {} has no levels but LTE is False'.format(self.tag))
        grid = self.spectral_grid.grid
        self.g32 = engine.lut_tensor(len(self.PTcouples), n_sets, len(grid), zero=True)
        idx = list(range(len(self.PTcouples))) if cells is None else list(cells)
        if idx:
            ls = engine.LineSet(tab, grid, self.MM, n_sets)
            if idx == list(range(idx[0], idx[0] + len(idx))):   # contiguous: built in place
                ls.gcoeff_cells_f32([self.PTcouples[i] for i in idx],
                                    out=self.g32[idx[0]:idx[0] + len(idx)])
            else:
                sub = ls.gcoeff_cells_f32([self.PTcouples[i] for i in idx])
                self.g32[torch.as_tensor(idx, device="cuda")] = sub
            ls.close()
        for s, nam in enumerate(names):
            level = None if self.LTE else getattr(self.isomolec, nam)
            st = LutSet(self.mol, self.iso, self.MM, level=level)
            st.PTcouples = self.PTcouples
            st.spectral_grid = self.spectral_grid
            st._table = (self, s)
            self.sets[nam] = st
        self._dev = None
        if cartLUTs is not None and export_levels:
            self.export(os.path.join(cartLUTs, self.tag + date_stamp() + '.pic'))
        return self

    def device_lut(self):
        """engine.Lut handle over the resident float32 table."""
        if self._dev is None:
            energies = None if self.LTE else self.isomolec.level_energies()
            self._dev = engine.Lut(self.g32, self.PTcouples, self.mol, self.iso,
                                   self.isomolec.ratio, level_energies=energies)
        return self._dev

    def CPU_time_estimate(self, lines, PTcouples):
        """The reference's estimate of ITS build time in minutes (smm:791-801): 3 min per 30000
        lines and cell.  Kept for scripts that print it; the GPU build takes milliseconds per cell."""
        n_lin = len([lin for lin in lines if lin.Mol == self.mol and lin.Iso == self.iso])
        return n_lin * 3. / 30000. * len(PTcouples)

    def find_lev(self, lev_string):
        for lev, st in self.sets.items():
            if st.level is not None and st.level.equiv(lev_string):
                return True, lev
        return False, None

    def merge(self, LUT):
        """Concatenate the cells of another LUT of the same isotopologue and range (smm:699-716)."""
        import torch
        if self.wn_range != LUT.wn_range:
            raise ValueError('Incompatible LUTs, different wn_ranges: {} {}'.format(
                self.wn_range, LUT.wn_range))
        self.PTcouples += LUT.PTcouples
        self.g32 = engine.lut_cat([self.g32, LUT.g32])
        for st in self.sets.values():
            st.PTcouples = self.PTcouples
            st.free_memory()
        self._dev = None

    def export_levels(self, cartLUTs, stamp=None, dtype=np.float64, for_reference=False):
        """Writes the table in the reference's on-disk form (smm:726-788, 880-892, 1122-1161): one
        pickle stream per vibrational level (`<tag>_<lev><date>.pic`, `_alllev` for an LTE
        isotopologue) holding the PTcouples header and then, per cell, the dict
        {ctype: SpectralGcoeff} with the grid erased.  Returns {set name: filename}.
        for_reference: pickles the reference itself can load (_RefPickler: its module names,
        protocol 2) plus the skeleton file `<tag><date>.pic` that its check_LUT_exists /
        check_and_build_allluts look for (export_skeleton)."""
        stamp = date_stamp() if stamp is None else stamp
        names = self.set_names()
        host = self.g32.cpu().numpy()
        files = dict()
        for s, nam in enumerate(names):
            st = self.sets[nam]
            fn = os.path.join(cartLUTs, self.tag + ('_alllev' if self.LTE else '_' + nam) + stamp + '.pic')
            lev_string = '' if st.level is None else st.level.minimal_level_string()
            with open(fn, 'wb') as f:
                _dump(self.PTcouples, f, for_reference)
                for c, (P, T) in enumerate(self.PTcouples):
                    set_ = dict()
                    for k, ct in enumerate(CTYPES):
                        gigi = spcl.SpectralGcoeff(ct, self.spectral_grid, self.mol, self.iso, self.MM,
                                                   lev_string, unidentified_lines=st.unidentified_lines,
                                                   spectrum=host[c, s, k].astype(dtype), Pres=P, Temp=T)
                        gigi.erase_grid()
                        set_[ct] = gigi
                    _dump(set_, f, for_reference)
            st.filename = fn
            st.filenames = [fn]
            files[nam] = fn
        if for_reference:
            files['skeleton'] = self.export_skeleton(os.path.join(cartLUTs, self.tag + stamp + '.pic'),
                                                     for_reference=True)
        return files

    def skeleton(self):
        """A copy without spectra: level structure, PTcouples, spectral grid and, per LutSet, the
        names of the per-level files - what the reference pickles as `<tag><date>.pic` after a
        build (makeLUT_nonLTE_Gcoeffs, smm:1872-1875) and reads back in check_LUT_exists /
        check_and_build_allluts (smm:1405-1420, 1480-1485)."""
        sk = copy.copy(self)
        sk.g32, sk._dev = None, None
        sk.sets = dict()
        for nam, st in self.sets.items():
            s2 = copy.copy(st)
            s2._table, s2.temp_file, s2.sets = None, None, []
            sk.sets[nam] = s2
        return sk

    def export_skeleton(self, filename, for_reference=False):
        with open(filename, 'wb') as f:
            _dump(self.skeleton(), f, for_reference)
        return filename

    def import_levels(self, files, spectral_grid):
        """Reads per-level streams written by export_levels (or by the reference's
        LookUpTable.make / LutSet.add_PT; see read_lutset_stream) back into the resident float32
        table.  files: {set name: filename or list of filenames}; cells of several files are
        concatenated like LutSet.load_from_files (smm:923-956)."""
        self.spectral_grid = copy.deepcopy(spectral_grid)
        names = self.set_names()
        n_grid = len(spectral_grid.grid)
        table, pts = None, None
        for s, nam in enumerate(names):
            fl = files[nam]
            fl = [fl] if isinstance(fl, str) else list(fl)
            pt_s, rows = [], []
            for fn in fl:
                pt_f, sets_f = read_lutset_stream(fn)
                pt_s += pt_f
                rows += sets_f
            if pts is None:
                pts = pt_s
                table = np.zeros((len(pts), len(names), 3, n_grid), dtype=np.float32)
            elif pt_s != pts:
                raise ValueError('LUT files of {} hold different PTcouples'.format(nam))
            for c, set_ in enumerate(rows):
                for k, ct in enumerate(CTYPES):
                    spe = set_[ct]
                    spe = spe.spectrum if hasattr(spe, 'spectrum') else spe
                    if spe is not None:                      # None: all-zero spectrum (smm:1674-1679)
                        table[c, s, k] = np.asarray(spe, dtype=np.float32)
            level = None if self.LTE else getattr(self.isomolec, nam)
            st = LutSet(self.mol, self.iso, self.MM, level=level, filename=fl[0])
            st.filenames = fl
            st.spectral_grid = self.spectral_grid
            st._table = (self, s)
            self.sets[nam] = st
        self.PTcouples = [list(map(float, pt)) for pt in pts]
        for st in self.sets.values():
            st.PTcouples = self.PTcouples
        self.g32 = engine.lut_from_host(table)
        self._dev = None
        return self

    def add_split_file(self, filename):
        if not hasattr(self, 'splitfiles'):
            self.splitfiles = []
        self.splitfiles.append(filename)

    def load_split(self, nsp):
        """Loads wavenumber chunk `nsp` written by split_and_compress_LUTS (here or by the
        reference, smm:822-838) as THE table of this LUT: the sets, the spectral grid and the
        resident float32 tensor then cover that chunk only."""
        split = read_split_file(self.splitfiles[nsp])
        names = self.set_names()
        first = split[names[0]]
        self.spectral_grid = first.spectral_grid
        self.PTcouples = [list(map(float, pt)) for pt in first.PTcouples]
        n_grid = len(self.spectral_grid.grid)
        table = np.zeros((len(self.PTcouples), len(names), 3, n_grid), dtype=np.float32)
        for s, nam in enumerate(names):
            st = split[nam]
            for c, set_ in enumerate(st.sets):
                for k, ct in enumerate(CTYPES):
                    if set_[ct] is not None:
                        table[c, s, k] = set_[ct].spectrum
                        set_[ct].double_precision()
                        set_[ct].restore_grid(st.spectral_grid, link_grid=True)
            st._table = None
            self.sets[nam] = st
        self.g32 = engine.lut_from_host(table)
        self._dev = None
        return self

    def export(self, filename):
        """Header (PTcouples) first, then the table, like the reference's per-level files
        (smm:880-892): resumable by check_LUT_exists()."""
        with open(filename, 'wb') as f:
            pickle.dump(self.PTcouples, f, protocol=-1)
            pickle.dump(dict(tag=self.tag, wn_range=self.wn_range, mol=self.mol, iso=self.iso,
                             LTE=self.LTE, sets=self.set_names(),
                             grid=self.spectral_grid.grid, g32=self.g32.cpu().numpy()), f,
                        protocol=-1)
        self.filename = filename
        return filename


class _RefUnpickler(pickle.Unpickler):
    """Resolves the reference's top-level module names (`spect_classes`, `spect_main_module`,
    `spect_base_module`) to this package, so that pickles written by the reference load here."""

    def find_class(self, module, name):
        if module in ('spect_classes', 'spect_main_module', 'spect_base_module'):
            module = __package__ + '.' + module
        return pickle.Unpickler.find_class(self, module, name)


_REF_MODULE_NAMES = {
    __package__ + '.spect_classes': 'spect_classes',
    __package__ + '.spect_main_module': 'spect_main_module',
    __package__ + '.spect_base_module': 'spect_base_module',
    'numpy._core.multiarray': 'numpy.core.multiarray',      # NumPy >= 2 names its reconstructors
    'numpy._core.numeric': 'numpy.core.numeric',            # under numpy._core
}


class _RefPickler(pickle._Pickler):
    """The inverse of _RefUnpickler: writes pickles the REFERENCE can load - classes of this
    package are stored under the reference's top-level module names (`spect_classes`,
    `spect_main_module`, `spect_base_module`), NumPy's reconstructors under their pre-2.0 module
    path, protocol 2 (the highest Python 2 reads; bytes travel as latin-1 text, which Python 2 loads
    as str).  Pure-Python pickler: meant for LUT files of interoperable size, not for speed."""

    def __init__(self, file):
        pickle._Pickler.__init__(self, file, protocol=2, fix_imports=True)

    def save_global(self, obj, name=None):
        new = _REF_MODULE_NAMES.get(getattr(obj, '__module__', None))
        if new is None:
            return pickle._Pickler.save_global(self, obj, name)
        name = name or getattr(obj, '__qualname__', None) or obj.__name__
        self.write(pickle.GLOBAL + new.encode('ascii') + b'\n' + name.encode('ascii') + b'\n')
        self.memoize(obj)

    dispatch = dict(pickle._Pickler.dispatch)
    dispatch[types.FunctionType] = save_global          # (type objects reach it through save_type)


def _dump(obj, f, for_reference=False):
    """One pickle.dump of the LUT writers: this interpreter's highest protocol, or the
    reference-readable form."""
    if for_reference:
        _RefPickler(f).dump(obj)
    else:
        pickle.dump(obj, f, protocol=-1)


def read_lutset_stream(filename):
    """(PTcouples, [ {ctype: SpectralGcoeff}, ... ]) of one per-level LUT stream (smm:900-921):
    the header, then one dict per cell until the stream ends (a build that was interrupted
    leaves fewer cells than the header announces; only the complete ones are returned)."""
    with open(filename, 'rb') as f:
        # every pickle.dump of the writer is a stream of its own (fresh memo): one Unpickler each
        pts = [list(map(float, pt)) for pt in _RefUnpickler(f, encoding='latin1').load()]
        sets = []
        for _ in pts:
            try:
                sets.append(_RefUnpickler(f, encoding='latin1').load())
            except (EOFError, pickle.UnpicklingError):
                break
    return pts[:len(sets)], sets


def read_split_file(filename):
    """{set name: LutSet} of one split / compressed LUT file (smm:1684-1712, read back by
    LookUpTable.load_split, smm:822-838): one pickle `[set name, LutSet]` per vibrational level
    ('all' for an LTE isotopologue); the LutSet carries the chunk's SpectralGrid, its PTcouples and
    per cell {ctype: float32 SpectralGcoeff without grid, or None for an all-zero spectrum}."""
    out = dict()
    with open(filename, 'rb') as f:
        while True:
            try:
                nam, st = _RefUnpickler(f, encoding='latin1').load()
            except EOFError:
                break
            out[nam] = st
    return out


def split_and_compress_LUTS(spectral_grid, allLUTs, cartLUTs, n_threads=n_threads, n_split=None,
                            ram_max=8., dim_tot=20., low_thres=1.e-30, for_reference=False):
    """Writes every LUT as n_split contiguous wavenumber chunks in the reference's split format
    (smm:1614-1728): chunks of ceil(n_grid/n_split) points, float32, all-zero spectra as None, one
    file `LUT_csplitNN_...<date>.pic` per chunk.  The LOS path here keeps the whole float32 table
    resident on the device and does not need the files; they are written for interoperability with
    the reference's radtrans / load_split.  Returns (allLUTs, n_split, chunk SpectralGrids)."""
    if n_split is None:
        n_split = int(np.ceil(dim_tot * n_threads / ram_max))
    grid = spectral_grid.grid
    len_split = int(np.ceil(1.0 * len(grid) / n_split))
    sp_grids = []
    for nsp in range(n_split):
        g = copy.deepcopy(spectral_grid)
        g.grid = grid[nsp * len_split:(nsp + 1) * len_split]
        sp_grids.append(g)
    for key, LUT in allLUTs.items():
        if LUT is None:
            continue
        host = LUT.g32.cpu().numpy()
        LUT.splitfiles = []
        for nsp, spgri in enumerate(sp_grids):
            fn = os.path.join(cartLUTs, lut_name_split(LUT.mol, LUT.iso, LUT.LTE, nsp) + date_stamp() + '.pic')
            lo = nsp * len_split
            with open(fn, 'wb') as f:
                for s, nam in enumerate(LUT.set_names()):
                    src = LUT.sets[nam]
                    st = LutSet(LUT.mol, LUT.iso, LUT.MM, level=src.level)
                    st.PTcouples = copy.deepcopy(LUT.PTcouples)
                    st.spectral_grid = spgri
                    lev_string = '' if src.level is None else src.level.minimal_level_string()
                    for c, (P, T) in enumerate(LUT.PTcouples):
                        d = dict()
                        for k, ct in enumerate(CTYPES):
                            spe = host[c, s, k, lo:lo + len(spgri.grid)]
                            if not spe.max() > 0.0:
                                d[ct] = None
                                continue
                            co = spcl.SpectralGcoeff(ct, spgri, LUT.mol, LUT.iso, LUT.MM, lev_string,
                                                     unidentified_lines=src.unidentified_lines,
                                                     spectrum=spe, Pres=P, Temp=T)
                            co.spectrum = spe.astype(np.float32)
                            co.erase_grid()
                            d[ct] = co
                        st.sets.append(d)
                    del st._table
                    _dump([nam, st], f, for_reference)     # for_reference: see _RefPickler
            LUT.splitfiles.append(fn)
    return allLUTs, n_split, sp_grids


def best_compressed_grid(simuls, thress=(1.e-3, 1.e-4, 1.e-5), factors=(5, 20, 100), skip_thres=1.e-20,
                         consider_derivatives=True, factor_minor=10, thres_minor=1.e-2, alg=2):
    """Union of the degraded grids of a collection of simulated spectra (smm:1513-1539).
    simuls: list of {tag: SpectralObject}.  Spectra whose maximum is below skip_thres are skipped;
    those below thres_minor x the overall maximum use thresholds factor_minor times larger.
    alg 1: degrade_grid(thress, factors, consider_derivatives); alg 2: degrade_grid2(max threshold,
    no derivatives).  (alg 3, degrade_grid3, raises NameError in the reference and is not here.)"""
    maxo = max([0.] + [spe.max() for singles in simuls for spe in singles.values()])
    grids = [np.array([])]
    for singles in simuls:
        for spe in singles.values():
            if spe.max() < skip_thres:
                continue
            th = list(factor_minor * np.array(thress)) if spe.max() < thres_minor * maxo else list(thress)
            if alg == 1:
                low = spe.degrade_grid(thress=th, factors=factors, consider_derivatives=consider_derivatives)
            elif alg == 2:
                low = spe.degrade_grid2(thres=max(th), consider_derivatives=False)
            else:
                raise ValueError('alg has to be 1 or 2')
            grids.append(low.spectral_grid.grid)
    return np.unique(np.concatenate(grids))


ORBIT_FIELDS = ('num dist sub_obs_lat sub_obs_lon limb_tg_alt limb_tg_lat limb_tg_lon limb_tg_sza '
                'pixel_rot phase_ang sub_solar_lat sub_solar_lon sun_dist time').split()


def read_orbits(filename, formato='VIMSselect', tag=None):
    """Geometry records of a 'VIMSselect' orbit file (smm:2328-2347): after the '#' header line,
    one line of 14 numbers per observation -> list of dicts (ORBIT_FIELDS + filename, tag) ready
    for sbm.VIMSPixel(orb.keys(), orb.values()).  The spectra / band / noise readers that go with
    it (sbm.read_obs, read_bands, read_noise) belong to the module that is missing upstream and
    their file formats are not documented anywhere in the reference: not provided."""
    if formato != 'VIMSselect':
        raise ValueError('formato {} not available'.format(formato))
    orbits = []
    with open(filename, 'r') as infile:
        sbm.find_spip(infile)
        for lin in infile.readlines():
            if not lin.strip():
                continue
            cose = list(map(float, lin.split()))
            orb = dict(zip(ORBIT_FIELDS, cose))
            orb['num'] = int(cose[0])
            orb['filename'] = filename
            orb['tag'] = tag
            orbits.append(orb)
    return orbits


def tolowres(hires, obs):
    """hires (converted IN PLACE to nm and to a regular grid) convolved to the observation's grid
    with its band widths through the host convolve_to_grid, in W/m2 (smm:3472-3477)."""
    hires.convertto_nm()
    hires.interp_to_regular_grid()
    lowres = hires.convolve_to_grid(obs.spectral_grid, spectral_widths=obs.bands.spectrum)
    lowres.convertto('Wm2')
    return lowres


def check_LUT_exists(PTcouples, cartLUTs, mol, iso, LTE):
    """(exists, PTcouples still to do, [[file, its PTcouples], ...], wn_ranges) for the LUT files of
    this isotopologue in cartLUTs, with the reference's rule (smm:1390-1456): the files whose name
    holds the LUT tag and not 'lev' - the pickled LookUpTable skeletons written next to the
    per-level streams (export_skeleton, or the reference's own) and the single-file tables of
    LookUpTable.export - are opened for their PTcouples and spectral range; a couple counts as done
    when both numbers are close (sbm.isclose).  (False, PTcouples, None, None) when there is no
    such file; the reference returns three values there and its caller fails to unpack them.)"""
    tag = lut_name(mol, iso, LTE)
    found = []
    if cartLUTs is not None and os.path.isdir(cartLUTs):
        found = sorted(fn for fn in os.listdir(cartLUTs) if tag in fn and 'lev' not in fn)
    if len(found) == 0:
        return False, PTcouples, None, None
    pt_done, pt_map, wn_ranges = [], [], []
    for fn in found:
        path = os.path.join(cartLUTs, fn)
        with open(path, 'rb') as f:
            first = _RefUnpickler(f, encoding='latin1').load()
            if isinstance(first, LookUpTable):                   # skeleton
                pts, grid = first.PTcouples, first.spectral_grid.grid
            else:                                                # LookUpTable.export: header, table
                pts, grid = first, _RefUnpickler(f, encoding='latin1').load()['grid']
        pts = [list(map(float, pt)) for pt in pts]
        pt_done += pts
        pt_map.append([path, pts])
        wn_ranges.append([float(np.min(grid)), float(np.max(grid))])
    pt_to_do = [pt for pt in PTcouples
                if not any(np.all(sbm.isclose(np.array(pt, dtype=float), np.array(ptd))) for ptd in pt_done)]
    return True, pt_to_do, pt_map, wn_ranges


def makeLUT_nonLTE_Gcoeffs(spectral_grid, lines, isomolec, LTE=True, atmosphere=None,
                           cartLUTs=None, n_threads=n_threads, test=False, PTcouples=None,
                           LUTopt=dict()):
    """Build the LUT of one isotopologue (smm:1847-1877)."""
    if PTcouples is None:
        PTcouples = calc_PT_couples_atmosphere(lines, isomolec, atmosphere, **LUTopt)
    if test:
        PTcouples = PTcouples[:2]
    wn_range = [spectral_grid.grid[0], spectral_grid.grid[-1]]
    LUT = LookUpTable(isomolec, wn_range, LTE)
    LUT.make(spectral_grid, lines, PTcouples, cartLUTs=cartLUTs, n_threads=n_threads)
    return LUT


def check_and_build_allluts(inputs, sp_grid, lines, molecs, atmosphere=None, PTcouples=None,
                            LUTopt=dict(), check_wn_range=True):
    """{(mol_name, iso): LookUpTable} for every isotopologue of `molecs` that has lines in the
    range (None otherwise, smm:1459-1510)."""
    if PTcouples is None:
        PTcouples = calc_PT_couples_atmosphere(lines, list(molecs), atmosphere, **LUTopt)
    cart = inputs.get('cart_LUTS') if isinstance(inputs, dict) else None
    allLUTs = dict()
    for molec in molecs:
        for isoname in molec.all_iso:
            isomol = getattr(molec, isoname)
            mine = [lin for lin in lines if lin.Mol == isomol.mol and lin.Iso == isomol.iso]
            if not mine:
                allLUTs[(isomol.mol_name, isomol.iso)] = None
                continue
            LTE = isomol.is_in_LTE or len(isomol.levels) == 0
            allLUTs[(isomol.mol_name, isomol.iso)] = makeLUT_nonLTE_Gcoeffs(
                sp_grid, mine, isomol, LTE=LTE, cartLUTs=cart, PTcouples=PTcouples)
    return allLUTs


def read_Gcoeffs_from_LUTs(cartLUTs, fileLUTs):
    """The pickled LookUpTable skeleton the reference's makeLUT_nonLTE_Gcoeffs leaves next to its
    per-level files (smm:1380-1388): level structure and file names, no spectra."""
    with open(cartLUTs + fileLUTs, 'rb') as f:
        return _RefUnpickler(f, encoding='latin1').load()


class AbsSetLOS(object):
    """Absorption or emission coefficients along one LOS, one SpectralObject per step
    (smm:1179-1258): kept in `set` (add_set) or streamed to `filename` without their grid
    (prepare_export / add_dump) and read back one at a time (prepare_read / read_one).
    Indexing, len() and iteration go over `set`."""

    def __init__(self, filename, spectral_grid=None, indices=None):
        self.indices = indices if indices is not None else []
        self.counter = 0
        self.remaining = 0
        self.filename = filename
        self.temp_file = None
        self.set = []
        self.spectral_grid = spectral_grid

    def __getitem__(self, k):
        return self.set[k]

    def __len__(self):
        return len(self.set)

    def __iter__(self):
        return iter(self.set)

    def _need_filename(self):
        if self.filename is None:
            raise ValueError('ERROR!: NO filename set for LutSet.')

    def prepare_read(self, read_spectral_grid=True):
        self._need_filename()
        self.temp_file = open(self.filename, 'rb')
        self.remaining = self.counter
        if read_spectral_grid:
            self.spectral_grid = _RefUnpickler(self.temp_file, encoding='latin1').load()

    def prepare_export(self):
        """Opens the file and writes the spectral grid on top (when there is one)."""
        self._need_filename()
        self.temp_file = open(self.filename, 'wb')
        if self.spectral_grid is not None:
            pickle.dump(self.spectral_grid, self.temp_file, protocol=-1)

    def finalize_IO(self):
        self.temp_file.close()
        self.temp_file = None

    def add_dump(self, set_, no_spectral_grid=True):
        if no_spectral_grid:
            for co in (set_.values() if type(set_) is dict else [set_]):
                co.erase_grid()
        pickle.dump(set_, self.temp_file, protocol=-1)
        self.counter += 1

    def add_set(self, set_):
        self.set.append(set_)
        self.counter += 1

    def read_one(self):
        if self.temp_file is None:
            self.prepare_read()
        set_ = _RefUnpickler(self.temp_file, encoding='latin1').load()
        for co in (set_.values() if type(set_) is dict else [set_]):
            co.restore_grid(self.spectral_grid, link_grid=True)
        self.remaining -= 1
        return set_


# ---------------------------------------------------------------------------------------------
# absorption / emission coefficients and the LOS integral
# ---------------------------------------------------------------------------------------------
def _abs_set_writer(spectral_grid, isomolec, tagLOS, cartDROP, to_disk):
    """(fill(kind, tag, rows) -> AbsSetLOS, tag of this LOS and isotopologue): file names and the
    keep-or-stream choice shared by the two make_abscoeff routines (smm:2164-2183, 2259-2274)."""
    if cartDROP is None:
        cartDROP = 'stuff_' + date_stamp() + '/'
    if to_disk and not os.path.exists(cartDROP):
        os.makedirs(cartDROP)

    def fill(kind, tag, rows):
        out = AbsSetLOS(cartDROP + kind + '_' + tag + '.pic', spectral_grid=spectral_grid)
        if to_disk:
            out.prepare_export()
        for row in rows:
            co = spcl.SpectralObject(row, spectral_grid, link_grid=True)
            out.add_dump(co) if to_disk else out.add_set(co)
        if to_disk:
            out.finalize_IO()
        return out

    return fill, ('LOS' if tagLOS is None else tagLOS) + '_mol_{}_iso_{}'.format(isomolec.mol,
                                                                                isomolec.iso)


def make_abscoeff_LUTS_fast(spectral_grid, isomolec, Temps, Press, LTE=True, tagLOS=None,
                            allLUTs=None, cartDROP=None, store_in_memory=False, track_levels=None,
                            time_control=False):
    """Absorption and emission coefficients of one isotopologue at the LOS steps (Temps, Press)
    from the LUT (smm:2134-2299), as two AbsSetLOS with one SpectralObject per step (indexable);
    with store_in_memory (the reference's name for "stream them to cartDROP") they are written to
    `abscoeff_<tagLOS>_mol_M_iso_I.pic` / `emicoeff_...` instead of being kept in `.set`.
    Vibrational temperatures are read from Level.local_vibtemp[step] when LTE is False.
    track_levels (a list of level names): two more dicts {level: AbsSetLOS}, the emission and the
    absorption of those levels alone - and, like the reference (:2263, :2273), the tracked
    ABSORPTION sets receive the tracked EMISSION coefficients.  None / (None x 4) when the
    isotopologue has no LUT in this range."""
    LUT = allLUTs[(isomolec.mol_name, isomolec.iso)]
    if LUT is None:
        return (None, None) if track_levels is None else (None, None, None, None)
    try:
        len(Press), len(Temps)
    except TypeError:
        Press, Temps = [Press], [Temps]
    Temps, Press = np.asarray(Temps, dtype=float), np.asarray(Press, dtype=float)
    n = len(Temps)
    names = LUT.set_names()

    def tvib_of(levels):
        if LUT.LTE:
            return None
        tv = np.empty((1, len(levels), 1, n))
        for s, lev in enumerate(levels):
            tv[0, s, 0] = Temps if LTE else np.asarray(getattr(isomolec, lev).local_vibtemp[:n], dtype=float)
        return tv

    # unit column per unit isotopic ratio: tau = abs coefficient, J = emission coefficient
    column = np.full((1, 1, n), 1.0 / LUT.isomolec.ratio)
    steps = engine.LosSteps([n], Temps[None], Press[None], column, tvib_of(names))
    a, e = engine.los_abs_emi([LUT.device_lut()], steps)
    a, e = a.cpu().numpy()[0], e.cpu().numpy()[0]

    fill, tagg = _abs_set_writer(spectral_grid, isomolec, tagLOS, cartDROP, store_in_memory)
    abs_c, emi_c = fill('abscoeff', tagg, a), fill('emicoeff', tagg, e)
    if track_levels is None:
        return abs_c, emi_c
    emi_tr, abs_tr = dict(), dict()
    for lev in track_levels:
        if LUT.LTE or lev not in names:
            rows = np.zeros((n, len(spectral_grid.grid)))      # never added to (:2250-2258)
        else:
            s_ = names.index(lev)
            sub = engine.lut_tensor(LUT.g32.shape[0], 1, LUT.g32.shape[3])
            sub.copy_(LUT.g32[:, s_:s_ + 1])
            one = engine.Lut(sub, LUT.PTcouples, LUT.mol, LUT.iso, LUT.isomolec.ratio,
                             level_energies=[getattr(isomolec, lev).energy])
            st1 = engine.LosSteps([n], Temps[None], Press[None], column, tvib_of([lev]))
            rows = engine.los_abs_emi([one], st1)[1].cpu().numpy()[0]
            one.close()
        tagl = tagg + '_{}'.format(lev)
        emi_tr[lev] = fill('tracklevel_emicoeff', tagl, rows)
        abs_tr[lev] = fill('tracklevel_abscoeff', tagl, rows)
    return abs_c, emi_c, emi_tr, abs_tr


def make_abscoeff_isomolec(wn_range_tot, isomolec, Temps, Press, LTE=True, allLUTs=None,
                           useLUTs=False, lines=None, store_in_memory=False, tagLOS=None,
                           cartDROP=None, track_levels=None, n_threads=n_threads):
    """The reference's slow twin of make_abscoeff_LUTS_fast (smm:1880-2131): with useLUTs the
    coefficients come from the LUT (on the LUT's grid), without it from a line-by-line evaluation
    of the G coefficients at every step's own (P, T) on prepare_spe_grid(wn_range_tot) - one
    batched FP64 GPU call (K1) instead of the per-step calc_shapes_lines / add_PT / pickle round
    trip through cartDROP.  Same populations, same return values and file names as the fast
    routine; more than 10 steps switch store_in_memory on like the reference (:1894-1895)."""
    try:
        len(Press), len(Temps)
    except TypeError:
        Press, Temps = [Press], [Temps]
    if len(Temps) > 10:
        store_in_memory = True
    if useLUTs:
        LUT = allLUTs[(isomolec.mol_name, isomolec.iso)]
        return make_abscoeff_LUTS_fast(LUT.spectral_grid, isomolec, Temps, Press, LTE=LTE,
                                       tagLOS=tagLOS, allLUTs=allLUTs, cartDROP=cartDROP,
                                       store_in_memory=store_in_memory, track_levels=track_levels)
    if lines is None:
        raise ValueError('when calling smm.make_abscoeff_isomolec() with useLUTs = False, you need '
                         'to give the list of spectral lines of isomolec as input')
    spectral_grid = prepare_spe_grid(wn_range_tot).spectral_grid
    lte_unid = len(isomolec.levels) == 0
    # a table with exactly one cell per step: interpolating at a node returns the node
    PT = [[float(p), float(t)] for p, t in zip(Press, Temps)]
    cells = []
    for pt in PT:
        if pt not in cells:
            cells.append(pt)
    mine = [lin for lin in lines if lin.Mol == isomolec.mol and lin.Iso == isomolec.iso]
    tab = spcl.line_table(mine, None if lte_unid else isomolec)
    grid = spectral_grid.grid
    Temps, Press = np.asarray(Temps, dtype=float), np.asarray(Press, dtype=float)
    n, names = len(Temps), (['all'] if lte_unid else list(isomolec.levels))
    import torch
    ls = engine.LineSet(tab, grid, isomolec.MM, tab["n_sets"])
    G = ls.gcoeff_cells(cells)                                        # [n_cells, n_sets, 3, n_grid]
    ls.close()
    idx = torch.as_tensor([cells.index(pt) for pt in PT], device="cuda")
    q = np.array([spcl.CalcPartitionSum(isomolec.mol, isomolec.iso, temp=t) for t in Temps])
    if lte_unid:
        pop = (1.0 / q)[:, None]
    else:
        tv = np.array([Temps if LTE else np.asarray(getattr(isomolec, lev).local_vibtemp[:n], dtype=float)
                       for lev in names]).T                            # [n, n_sets]
        pop = spcl.Boltz_ratio_nodeg(isomolec.level_energies()[None, :], tv) / q[:, None]
    w = torch.as_tensor(pop, dtype=torch.float64, device="cuda")
    Gs = G[idx]                                                       # [n, n_sets, 3, n_grid]
    a = torch.einsum('ks,ksp->kp', w, Gs[:, :, 2] - Gs[:, :, 1]).cpu().numpy()
    e = torch.einsum('ks,ksp->kp', w, Gs[:, :, 0]).cpu().numpy()

    fill, tagg = _abs_set_writer(spectral_grid, isomolec, tagLOS, cartDROP, store_in_memory)
    abs_c, emi_c = fill('abscoeff', tagg, a), fill('emicoeff', tagg, e)
    if track_levels is None:
        return abs_c, emi_c
    emi_tr, abs_tr = dict(), dict()
    for lev in track_levels:
        if lte_unid or lev not in names:
            rows = np.zeros((n, len(grid)))
        else:
            s_ = names.index(lev)
            rows = (w[:, s_, None] * Gs[:, s_, 0]).cpu().numpy()
        tagl = tagg + '_{}'.format(lev)
        emi_tr[lev] = fill('tracklevel_emicoeff', tagl, rows)
        abs_tr[lev] = fill('tracklevel_abscoeff', tagl, rows)
    return abs_c, emi_c, emi_tr, abs_tr


# ---------------------------------------------------------------------------------------------
# parameter space of the retrieval (smm:161-296, 319-352, 442-656): only what the forward model
# and its Jacobians need; the update step itself (inversion_algebra) is further down, host NumPy
# ---------------------------------------------------------------------------------------------
def alt_triangle(alt_grid, node_alt, step=None, node_lo=None, node_up=None, first=False,
                 last=False):
    """Triangular weight of one altitude node on alt_grid (smm:319-352), as an sbm.AtmGridMask.
    first / last: constant 1 below / above the node."""
    if step is not None:
        node_lo, node_up = node_alt - step, node_alt + step
    alt_grid = np.asarray(alt_grid, dtype=float)
    cos = np.zeros(len(alt_grid))
    up = None if node_up is None else 1.0 - (alt_grid - node_alt) / (node_up - node_alt)
    lo = None if node_lo is None else 1.0 - (node_alt - alt_grid) / (node_alt - node_lo)
    if first:
        cos = np.where(alt_grid < node_alt, 1.0, np.where(alt_grid < node_up, up, 0.0))
    elif last:
        cos = np.where(alt_grid > node_alt, 1.0, np.where(alt_grid > node_lo, lo, 0.0))
    else:
        inside = (alt_grid >= node_lo) & (alt_grid <= node_up)
        cos = np.where(inside, np.where(alt_grid >= node_alt, up, lo), 0.0)
    return sbm.AtmGridMask(sbm.AtmGrid('alt', alt_grid), cos, 'lin')


def lat_box(lat_limits, lat_ok):
    """Latitude mask with a box function (smm:355-374): lat_limits are the START latitudes of the
    boxes, ascending, the last box is open-ended; 1 in the box that contains lat_ok.  Kept literal:
    the last box needs lat_ok STRICTLY above its start (:368), so LinearProfile_2D, which passes the
    box starts themselves, leaves the parameters of its last box with an all-zero mask."""
    lat_limits = np.asarray(lat_limits, dtype=float)
    cos = []
    for lat1, lat2 in zip(lat_limits[:-1], lat_limits[1:]):
        cos.append(1.0 if lat1 <= lat_ok < lat2 else 0.0)
    cos.append(1.0 if lat_ok > lat_limits[-1] else 0.0)
    return sbm.AtmGridMask(sbm.AtmGrid('lat', lat_limits), np.array(cos), 'box')


def centre_boxes(lat_limits):
    """[-90, -60, -30, ...] -> box centres [-75, -45, ...] (smm:377-385)."""
    return [(la1 + la2) / 2.0 for la1, la2 in zip(lat_limits[:-1], lat_limits[1:])]


class RetParam(object):
    """A single parameter of the parameter space (smm:600-656)."""

    def __init__(self, nameset, key, maskgrid, apriori, apriori_err, first_guess=None,
                 constrain_positive=True):
        self.nameset = nameset
        self.key = key
        self.maskgrid = copy.deepcopy(maskgrid)
        self.value = apriori if first_guess is None else first_guess
        self.apriori = apriori
        self.apriori_err = apriori_err
        self.derivatives = []
        self.old_values = []
        self.constrain_positive = constrain_positive
        self.not_involved = False
        self.is_used = False
        self.hires_deriv = None

    def set_not_involved(self):
        self.not_involved = True

    def set_used(self):
        self.is_used = True

    def set_involved(self):
        self.not_involved = False

    def update_par(self, delta_par):
        self.old_values.append(self.value)
        new_value = self.value + delta_par
        if self.constrain_positive:
            while new_value <= 0.0:
                delta_par /= 2
                new_value = self.value + delta_par
        self.value = new_value

    def add_hires_deriv(self, derivative):
        self.hires_deriv = copy.deepcopy(derivative)

    def erase_hires_deriv(self):
        self.hires_deriv = None

    def store_deriv(self, derivative, num):
        try:
            self.derivatives[num] = copy.deepcopy(derivative)
        except IndexError:
            self.derivatives.append(copy.deepcopy(derivative))


class RetSet(object):
    """Parameters of one quantity, e.g. the VMR profile of a gas (smm:259-283); `name` is the
    gas name for a VMR set."""

    def __init__(self, name, params):
        self.name = name
        self.set = [copy.deepcopy(par) for par in params]
        self.n_par = len(self.set)

    def items(self):
        return zip([par.key for par in self.set], self.set)

    def keys(self):
        return [par.key for par in self.set]

    def profile(self):
        """sum_p maskgrid_p * value_p (smm:483-489) as an AtmProfile named 'vmr' (and self.name)."""
        grid = sbm.mask_profile_grid(self.set[0].maskgrid)
        prof = sbm.AtmProfZeros(grid, 'vmr', self.set[0].maskgrid.interp['mask'])
        for par in self.set:
            prof += par.maskgrid * par.value
        prof.add_profile(prof.values['vmr'], self.name, prof.interp['vmr'])
        return prof


class LinearProfile_1D_new(RetSet):
    """Profile through linear interpolation of altitude nodes (smm:442-489)."""

    def __init__(self, name, alt_grid, alt_nodes, apriori_prof, apriori_prof_err,
                 first_guess_prof=None):
        self.name = name
        self.set = []
        self.n_par = len(alt_nodes)
        self.alts = list(alt_nodes)
        z = alt_grid.grid[0] if hasattr(alt_grid, 'grid') else np.asarray(alt_grid, dtype=float)
        if first_guess_prof is None:
            first_guess_prof = apriori_prof
        n = len(alt_nodes)
        for i in range(n):
            if i == 0:
                mask = alt_triangle(z, alt_nodes[0], node_up=alt_nodes[1], first=True)
            elif i == n - 1:
                mask = alt_triangle(z, alt_nodes[-1], node_lo=alt_nodes[-2], last=True)
            else:
                mask = alt_triangle(z, alt_nodes[i], node_up=alt_nodes[i + 1], node_lo=alt_nodes[i - 1])
            self.set.append(RetParam(name, alt_nodes[i], mask, apriori_prof[i], apriori_prof_err[i],
                                     first_guess=first_guess_prof[i]))

    def check_involved(self, parkey, coord_range):
        indp = self.alts.index(parkey)
        if indp == len(self.alts) - 1:
            return True
        return not coord_range['alt'][0] > self.alts[indp + 1]


class LinearProfile_1D(LinearProfile_1D_new):
    """Same with the reference's older signature taking the atmosphere (smm:563-597)."""

    def __init__(self, name, atmosphere, alt_nodes, apriori_prof, apriori_prof_err,
                 first_guess_prof=None):
        LinearProfile_1D_new.__init__(self, name, atmosphere.grid.coords['alt'], alt_nodes,
                                      apriori_prof, apriori_prof_err, first_guess_prof)
        self.orig_atmosphere = atmosphere


class LinearProfile_2D(RetSet):
    """Profile linear in altitude between the nodes, with latitude boxes (smm:388-439): one
    LinearProfile_1D_new per box, every parameter keyed (box start latitude, altitude node) and
    masked by alt_triangle x lat_box.  lat_limits are the box START latitudes (ascending)."""

    def __init__(self, name, atmosphere, alt_nodes, lat_limits, apriori_profs, apriori_prof_errs,
                 first_guess_profs=None):
        self.name = name
        self.set = []
        self.n_par = len(alt_nodes) * len(lat_limits)
        self.alts = list(alt_nodes)
        self.lats = list(lat_limits)
        z = atmosphere.grid.coords['alt']
        if first_guess_profs is None:
            first_guess_profs = apriori_profs
        for ap, er, fg, lat in zip(apriori_profs, apriori_prof_errs, first_guess_profs, lat_limits):
            coso = LinearProfile_1D_new(name, z, alt_nodes, ap, er, first_guess_prof=fg)
            latbox = lat_box(lat_limits, lat)
            for cos in coso.set:
                self.set.append(RetParam(name, (lat, cos.key), cos.maskgrid.merge(latbox),
                                         cos.apriori, cos.apriori_err, first_guess=cos.value))

    def check_involved(self, parkey, coord_range):
        indp = self.alts.index(parkey[1])
        involved = True
        if indp != len(self.alts) - 1 and coord_range['alt'][0] > self.alts[indp + 1]:
            involved = False
        latz = coord_range['lat']
        indl = self.lats.index(parkey[0])
        if indl == len(self.lats) - 1:
            if latz[1] < parkey[0]:
                involved = False
        elif latz[1] < parkey[0] or latz[0] > self.lats[indl + 1]:
            involved = False
        return involved


class BayesSet(object):
    """The full parameter space that drives the forward model (smm:161-257)."""

    def __init__(self, tag=None):
        self.tag = tag
        self.sets = dict()
        self.n_tot = 0
        self.order = []
        self.old_params = []

    def add_set(self, set_):
        self.sets[set_.name] = copy.deepcopy(set_)
        self.n_tot += set_.n_par
        self.order.append(set_.name)

    def values(self):
        return [par.value for par in self.params()]

    def params(self):
        return [par for nam in self.order for par in self.sets[nam].set]

    def n_used_par(self):
        return sum([par.is_used for par in self.params()])

    def param_vector(self):
        return np.array([par.value for par in self.params()])

    def apriori_vector(self):
        return np.array([par.apriori for par in self.params()])

    def VCM_apriori(self):
        return np.diag(np.array([par.apriori_err for par in self.params()], dtype=float) ** 2)

    def update_params(self, delta_x):
        """(smm:224-229)"""
        self.old_params.append(copy.deepcopy(self.param_vector()))
        for par, dx in zip(self.params(), delta_x):
            par.update_par(dx)

    def update_parerror(self):
        """ret_error of every parameter = sqrt of its diagonal element of the stored VCM (smm:191-194)."""
        for num, par in enumerate(self.params()):
            par.ret_error = np.sqrt(self.VCM[num, num])

    def store_avk(self, av_kernel):
        self.av_kernel = copy.deepcopy(av_kernel)

    def store_VCM(self, VCM):
        self.VCM = copy.deepcopy(VCM)

    def build_jacobian(self, masks=None):
        """[n_obs_points x n_tot] from the per-pixel derivative spectra stored on the parameters
        (smm:197-222)."""
        masktot = None
        if masks is not None:
            masktot = np.concatenate([np.asarray(m) for m in masks]).astype(bool)
        jac = []
        for par in self.params():
            dertot = np.concatenate([np.asarray(der.spectrum, dtype=float) for der in par.derivatives])
            jac.append(dertot if masktot is None else dertot[masktot])
        self.jacobian = np.array(jac).T
        return self.jacobian


def los_step_tables(loss, planet):
    """engine.LosSteps of a list of sbm.LineOfSight that went through calc_radtran_steps."""
    gi = loss[0].radtran_steps['gas_isos']
    nmax = max(len(l.radtran_steps['step']) for l in loss)
    tabs = [l.step_tables(planet, n_steps_max=nmax) for l in loss]
    n_steps = np.array([t[0] for t in tabs], dtype=np.int32)
    temp = np.stack([t[1] for t in tabs])
    pres = np.stack([t[2] for t in tabs])
    col = np.stack([t[3] for t in tabs], axis=1)            # [gas][los][step]
    tvib = np.stack([t[4] for t in tabs], axis=2)           # [gas][set][los][step]
    return gi, engine.LosSteps(n_steps, temp, pres, col, tvib)


def planet_atmosphere_tables(planet, gas_isos):
    """engine.Atmosphere (the tables of sr_atmosphere) from an sbm planet: T (linear), P
    (log-linear), one VMR profile per (gas, iso) entry and the vibrational-temperature profiles of
    the levels, all on the atmosphere's altitude grid; vibrational temperatures may carry a
    solar-zenith-angle axis (3-D profiles on ('lat', 'sza', 'alt'), radtran_3D_ch4.py:249-250)."""
    atm = planet.atmosphere
    z = atm.grid.coords['alt']
    two_d = 'lat' in atm.grid.names
    n_band = atm.n_band()

    def table(prof, name, sza_nodes=None):
        if not np.array_equal(prof.grid.coords['alt'], z):
            raise ValueError('profile %s is not on the atmosphere altitude grid' % name)
        v = np.asarray(prof.values[name], dtype=float)
        has_lat, has_sza = 'lat' in prof.grid.names, 'sza' in prof.grid.names
        if has_lat and prof.lat_interp.get(name, 'box') != 'box':
            raise NotImplementedError(
                'profile %s is interpolated linearly in latitude: the device step builder works '
                'on latitude bands (interp [\'box\', ...]); use the per-LOS host methods '
                'LineOfSight.calc_atm_intersections / calc_radtran_steps for such an atmosphere' % name)
        if has_lat and not np.array_equal(prof.lat_edges(), atm.lat_edges()):
            raise ValueError('profile %s is not on the latitude bands of the atmosphere' % name)
        if has_sza and (sza_nodes is None or not np.array_equal(prof.grid.coords['sza'], sza_nodes)):
            raise ValueError('profile %s: all SZA-dependent profiles must share their SZA nodes' % name)
        if has_lat and v.shape[0] != n_band:
            raise ValueError('profile %s has %d latitude bands, atmosphere has %d'
                             % (name, v.shape[0], n_band))
        if not has_lat:
            v = np.broadcast_to(v, (n_band,) + v.shape)
        if sza_nodes is not None and not has_sza:
            v = np.broadcast_to(v[:, None, :], (n_band, len(sza_nodes), len(z)))
        return v

    sza_nodes = None
    for g, iso in gas_isos:
        im = getattr(planet.gases[g], iso)
        for lev in im.levels:
            vt = getattr(im, lev).vibtemp
            if vt is not None and 'sza' in vt.grid.names and sza_nodes is None:
                sza_nodes = np.asarray(vt.grid.coords['sza'], dtype=float)
    n_lev = max([len(getattr(planet.gases[g], iso).levels) for g, iso in gas_isos] + [0])
    vmr = np.stack([table(planet.gases[g].abundance, 'vmr') for g, iso in gas_isos])
    tvib_on = -np.ones((len(gas_isos), n_lev), dtype=np.int32)
    shape = (n_band, len(z)) if sza_nodes is None else (n_band, len(sza_nodes), len(z))
    tvib = np.full((len(gas_isos), n_lev) + shape, 100.0)
    for m, (g, iso) in enumerate(gas_isos):
        im = getattr(planet.gases[g], iso)
        for j, lev in enumerate(im.levels):
            L = getattr(im, lev)
            if L.vibtemp is None:
                tvib_on[m, j] = 0
            else:
                tvib_on[m, j] = 1
                tvib[m, j] = table(L.vibtemp, 'vibtemp', sza_nodes)
    return engine.Atmosphere(z, table(atm, 'temp'), table(atm, 'pres'), vmr,
                             tvib=tvib if n_lev else None, tvib_on=tvib_on if n_lev else None,
                             lat_edges=atm.lat_edges() if two_d else None,
                             radius_km=planet.radius, top_km=planet.atm_extension,
                             sza_nodes=sza_nodes)


def latitude_linear_profiles(planet):
    """Names of the profiles of the planet (atmosphere, VMRs, vibrational temperatures) that are
    interpolated LINEARLY in latitude between band centres (radtran_3Dvs2D_radtrans_new.py:82-111,
    `lat_interp='lin'`) - the device step builder works on latitude bands and cannot take them."""
    found = []

    def look(prof, label):
        if prof is not None and 'lat' in prof.grid.names:
            found.extend(label + ':' + n for n in prof.names if prof.lat_interp.get(n, 'box') != 'box')

    look(planet.atmosphere, 'atmosphere')
    for g in sorted(planet.gases):
        look(planet.gases[g].abundance, g)
        for iso in planet.gases[g].all_iso:
            im = getattr(planet.gases[g], iso)
            for lev in im.levels:
                look(getattr(im, lev).vibtemp, '{}/{}/{}'.format(g, iso, lev))
    return found


def los_step_tables_host(loss, planet, bayes_set=None, set_name=None, delta_x=5.0,
                         max_T_variation=5.0, max_Plog_variation=1.0, max_opt_depth=None,
                         lines=None, ssps=None, fszas=None, use_tangent_sza=False,
                         LOS_order='radtran'):
    """The same tables as los_step_tables_device from the per-LOS host methods
    (calc_atm_intersections, calc_SZA_along_los or the tangent SZA, calc_radtran_steps - what the
    reference itself runs per LOS, smm:3133-3147): the route for atmospheres the device builder
    does not take (latitude-linear profiles).  Host time grows with the number of LOS; the
    radiances are computed on the device from these tables like any other."""
    for k, los in enumerate(loss):
        los.calc_atm_intersections(planet, delta_x=delta_x, LOS_order=LOS_order)
        if use_tangent_sza:
            if fszas is None:
                raise ValueError('use_tangent_sza needs the tangent-point SZA of every LOS')
            los.szas = np.full(len(los.intersections), float(fszas[k]))
        elif ssps is not None:
            los.calc_SZA_along_los(planet, ssps[k])
        los.calc_radtran_steps(planet, lines, calc_derivatives=bayes_set is not None,
                               bayes_set=bayes_set, max_T_variation=max_T_variation,
                               max_Plog_variation=max_Plog_variation, max_opt_depth=max_opt_depth)
    gi, steps = los_step_tables(loss, planet)
    dfrac = None
    if bayes_set is not None and set_name is not None and set_name in planet.gases:
        dfrac = los_jac_tables(loss, bayes_set, set_name, steps.n_steps_max)
    return gi, steps, dfrac


def los_step_tables_device(loss, planet, bayes_set=None, set_name=None, delta_x=5.0,
                           max_T_variation=5.0, max_Plog_variation=1.0, max_opt_depth=None,
                           lines=None, ssps=None, fszas=None, use_tangent_sza=False,
                           LOS_order='radtran'):
    """calc_atm_intersections + calc_SZA_along_los + calc_radtran_steps for a whole list of
    sbm.LineOfSight in ONE library call (sr_los_steps_build_rays).

    ssps: sub-solar point (sbm.Coords) per LOS -> SZA at every sample of the ray; fszas with
    use_tangent_sza: one SZA per LOS (smm:3138-3141); LOS_order 'photon' = invert_LOS_direction
    (smm:3135-3137); max_opt_depth needs `lines` (peak cross-sections, sbm.peak_cross_sections).
    Returns (gas_isos, engine.LosSteps, dfrac): dfrac [n_los][n_steps_max][n_par] for the parameters
    of bayes_set.sets[set_name] (a VMR set named like a gas of the planet), else None.  Also fills
    los.involved_retparams."""
    gi = [(g, iso) for g in sorted(planet.gases) for iso in planet.gases[g].all_iso]
    if latitude_linear_profiles(planet):
        return los_step_tables_host(loss, planet, bayes_set=bayes_set, set_name=set_name,
                                    delta_x=delta_x, max_T_variation=max_T_variation,
                                    max_Plog_variation=max_Plog_variation,
                                    max_opt_depth=max_opt_depth, lines=lines, ssps=ssps, fszas=fszas,
                                    use_tangent_sza=use_tangent_sza, LOS_order=LOS_order)
    atm = planet_atmosphere_tables(planet, gi)
    org = np.array([l.starting_point.Cartesian() for l in loss])
    drc = np.array([l.direction for l in loss])
    sun = sza_fixed = None
    if use_tangent_sza:
        if fszas is None:
            raise ValueError('use_tangent_sza needs the tangent-point SZA of every LOS')
        sza_fixed = np.asarray(fszas, dtype=float)
    elif ssps is not None:
        sun = np.array([p.Cartesian() for p in ssps])
    elif atm.n_sza > 1:
        raise ValueError('the vibrational temperatures depend on the SZA: pass the sub-solar points '
                         '(ssps) or use_tangent_sza with fszas')
    sigma = None
    if max_opt_depth is not None and max_opt_depth > 0.0:
        if not lines:
            raise ValueError('max_opt_depth needs the line list (peak cross-sections)')
        sg = sbm.peak_cross_sections(planet, lines)
        seen, sigma = set(), []
        for g, iso in gi:     # the column of a gas is counted once, on its first isotopologue entry
            sigma.append(0.0 if g in seen else sg[g])
            seen.add(g)
    masks, jac_gas, pars = None, -1, []
    if bayes_set is not None and set_name is not None and set_name in planet.gases:
        pars = bayes_set.sets[set_name].set
        lat_edges = planet.atmosphere.lat_edges() if 'lat' in planet.atmosphere.grid.names else None
        masks = np.array([np.broadcast_to(p.maskgrid.table(atm.z, lat_edges), (atm.n_band, len(atm.z)))
                          for p in pars])
        jac_gas = [g for g, iso in gi].index(set_name)
    steps, dfrac = engine.los_steps_build(atm, org, drc, delta_x=delta_x,
                                          max_T_variation=max_T_variation,
                                          max_Plog_variation=max_Plog_variation, masks=masks,
                                          jac_gas=jac_gas, sun=sun, sza_fixed=sza_fixed,
                                          max_opt_depth=max_opt_depth, sigma_peak=sigma,
                                          photon_order=(LOS_order == 'photon'))
    if bayes_set is not None:
        for l, los in enumerate(loss):
            for par in bayes_set.params():
                los.involved_retparams.setdefault((par.nameset, par.key), False)
            for q, par in enumerate(pars):
                los.involved_retparams[(par.nameset, par.key)] = bool(np.any(dfrac[l, :, q] != 0.0))
    return gi, steps, dfrac


def observation_channels(obs, sp_gri):
    """(centres, widths, units mode, intensity factor) that turn the device convolution into
    SpectralIntensity.hires_to_lowres(obs, spectral_widths=obs.bands.spectrum) (spcl:1180-1191):
    the hi-res radiances are in 'ergscm2' per cm-1 on sp_gri; the observation may be on a cm-1 axis
    (units mode 'same'), or on a wavelength axis in nm or micron (mode 'nm': converted per point
    and convolved on the wavelength axis on the device; micron channels are rescaled to nm), and in
    any of the intensity units of SpectralIntensity.convertto."""
    hi_units = getattr(sp_gri, 'units', 'cm_1')
    lo_units = obs.spectral_grid.units
    centres = np.asarray(obs.spectral_grid.grid, dtype=float)
    widths = np.asarray(obs.bands.spectrum, dtype=float)
    factor = spcl.SpectralIntensity.INTENSITY_TO_WM2['ergscm2'] / \
        spcl.SpectralIntensity.INTENSITY_TO_WM2[getattr(obs, 'units', 'ergscm2')]
    if lo_units == hi_units:
        return centres, widths, 'same', factor
    if hi_units == 'cm_1' and lo_units == 'nm':
        return centres, widths, 'nm', factor
    if hi_units == 'cm_1' and lo_units == 'mum':      # per micron = per nm * 1e3 (spcl:780-784)
        return centres * 1.e3, widths * 1.e3, 'nm', factor * 1.e3
    raise ValueError('observation axis in {} over a hi-res grid in {}: convert the observation to '
                     'cm_1, nm or mum first'.format(lo_units, hi_units))


def los_batch_radiances(loss, sp_grid, planet, LUTS, solo_absorption=False,
                        initial_intensity=None, lowres=None, pt0=0, n_pts=None, tables=None,
                        lowres_units='same'):
    """Radiances of a batch of lines of sight in ONE launch sequence.

    loss: sbm.LineOfSight objects with radtran_steps; LUTS: {(mol_name, iso): LookUpTable}.
    Returns a list of hi-res SpectralIntensity on sp_grid[pt0:pt0+n_pts], or, with
    lowres = (channel centres, channel widths) in the units of sp_grid, the CUDA tensor
    [n_los][n_chan] convolved on the device (hires_to_lowres)."""
    import torch
    gi, steps = los_step_tables(loss, planet) if tables is None else tables[:2]
    luts, keep = [], []
    for m, (g, iso) in enumerate(gi):
        im = getattr(planet.gases[g], iso)
        L = LUTS.get((im.mol_name, im.iso))
        if L is not None:
            luts.append(L.device_lut())
            keep.append(m)
    if not luts:
        raise ValueError('no LUT for any gas of the planet in this spectral range')
    if len(keep) != len(gi):
        steps = engine.LosSteps(steps.n_steps, steps.temp, steps.pres, steps.column[keep],
                                None if steps.tvib is None else steps.tvib[keep])
    grid = sp_grid.grid if hasattr(sp_grid, 'grid') else np.asarray(sp_grid)
    n_pts = len(grid) - pt0 if n_pts is None else n_pts
    i0 = None
    if initial_intensity is not None:
        i0 = torch.as_tensor(np.broadcast_to(np.asarray(initial_intensity, dtype=float),
                                             (len(loss), n_pts)).copy(), device="cuda")
    if lowres is not None:   # LOS blocks are reduced to the channels on the device, batch of any size
        gdev = torch.as_tensor(np.ascontiguousarray(grid, dtype=float), device="cuda")
        return engine.los_rt_lut_lowres(luts, steps, gdev, lowres[0], lowres[1], pt0=pt0,
                                        n_pts=n_pts, i0=i0, solo_absorption=solo_absorption,
                                        units=lowres_units)
    rad = engine.los_rt_lut(luts, steps, pt0=pt0, n_pts=n_pts, i0=i0,
                            solo_absorption=solo_absorption)
    rad = rad.cpu().numpy()
    units = getattr(sp_grid, 'units', 'cm_1')
    sub = spcl.SpectralGrid(grid[pt0:pt0 + n_pts], units=units)
    return [spcl.SpectralIntensity(rad[i], sub) for i in range(len(loss))]


def los_radiance_line_by_line(los, sp_grid, planet, lines, calc_derivatives=False, bayes_set=None,
                              solo_absorption=False, initial_intensity=None):
    """Hi-res radiance of ONE line of sight WITHOUT look-up tables (`useLUTs=False`, the reference's
    make_abscoeff_isomolec path, smm:1880-2131): the G-coefficient spectra of every isotopologue
    are evaluated line by line (K1: sr_gcoeff_cells_dev, FP64) at each step's own Curtis-Godson
    (P, T), weighted with the level populations (smm:2200-2250 arithmetic) into tau and J per
    step, and run through the layer recursion (K3: sr_los_rt_layers[_jac]_dev).  Returns
    [SpectralIntensity, {}, bayes_set copy (with hires_deriv when calc_derivatives)]."""
    import torch
    steps = los.radtran_steps['step']
    gi = los.radtran_steps['gas_isos']
    grid = sp_grid.grid
    n, n_grid = len(steps), len(grid)
    PT = [[st['pres'], st['temp']] for st in steps]
    temps = np.array([st['temp'] for st in steps])
    dev = dict(dtype=torch.float64, device="cuda")
    tau = torch.zeros((1, n, n_grid), **dev)
    emi = torch.zeros((1, n, n_grid), **dev)
    per_gas = dict()
    for g, iso in gi:
        im = getattr(planet.gases[g], iso)
        mine = [lin for lin in lines if lin.Mol == im.mol and lin.Iso == im.iso]
        if not mine:
            continue
        lte_unid = len(im.levels) == 0
        tab = spcl.line_table(mine, None if lte_unid else im)
        ls = engine.LineSet(tab, grid, im.MM, tab["n_sets"])
        G = ls.gcoeff_cells(PT)                                   # [n, n_sets, 3, n_grid]
        ls.close()
        q = np.array([spcl.CalcPartitionSum(im.mol, im.iso, temp=t) for t in temps])
        if lte_unid:
            pop = (1.0 / q)[:, None]
        else:
            tv = np.array([[st['vibtemps'][(g, iso, lev)] for lev in im.levels] for st in steps])
            pop = spcl.Boltz_ratio_nodeg(im.level_energies()[None, :], tv) / q[:, None]
        col = np.array([st['columns'][g] for st in steps]) * im.ratio
        w = torch.as_tensor(pop * col[:, None], **dev)            # [n, n_sets]
        a = torch.einsum('ks,ksp->kp', w, G[:, :, 2] - G[:, :, 1])
        e = torch.einsum('ks,ksp->kp', w, G[:, :, 0])
        del G
        tau[0] += a
        emi[0] += e
        ta, te = per_gas.get(g, (0, 0))
        per_gas[g] = (ta + a, te + e)
    nst = torch.tensor([n], dtype=torch.int32, device="cuda")
    i0 = None
    if initial_intensity is not None:
        i0 = torch.as_tensor(np.broadcast_to(np.asarray(initial_intensity, dtype=float), (1, n_grid)).copy(),
                             device="cuda")
    out_set = copy.deepcopy(bayes_set)
    if calc_derivatives and bayes_set is not None:
        rad = None
        jacs = []
        for nam in bayes_set.order:
            n_par = bayes_set.sets[nam].n_par
            if nam not in per_gas:
                jacs.append(torch.zeros((1, n_par, n_grid), **dev))
                continue
            dfrac = torch.as_tensor(los_jac_tables([los], bayes_set, nam, n), **dev)
            only = len(per_gas) == 1
            tg = None if only else per_gas[nam][0][None].contiguous()
            eg = None if only else per_gas[nam][1][None].contiguous()
            rad, jac = engine.los_rt_layers_jac(tau, emi, dfrac, nst, tau_g=tg, emi_g=eg, i0=i0,
                                                solo_absorption=solo_absorption)
            jacs.append(jac)
        jac = torch.cat(jacs, dim=1).cpu().numpy()[0]
        for par, d in zip(out_set.params(), jac):
            par.add_hires_deriv(spcl.SpectralIntensity(d, sp_grid))
    else:
        rad = None
    if rad is None:
        src = torch.where(tau == 0.0, torch.zeros_like(tau), emi / tau)
        rad = engine.los_rt_layers(tau, src, nst, i0=i0, solo_absorption=solo_absorption)
    return [spcl.SpectralIntensity(rad.cpu().numpy()[0], sp_grid), dict(), out_set]


def los_jac_tables(loss, bayes_set, set_name, n_steps_max):
    """dfrac [n_los][n_steps_max][n_par] = (d column / d parameter) / column of gas `set_name` for
    the parameters of bayes_set.sets[set_name], from the `dcolumns` that
    calc_radtran_steps(calc_derivatives=True, bayes_set=...) stored on every step."""
    pars = bayes_set.sets[set_name].set
    dfrac = np.zeros((len(loss), n_steps_max, len(pars)))
    for l, los in enumerate(loss):
        for k, st in enumerate(los.radtran_steps['step']):
            col = st['columns'][set_name]
            if col == 0.0:
                continue
            for q, par in enumerate(pars):
                dfrac[l, k, q] = st['dcolumns'][(par.nameset, par.key)] / col
    return dfrac


def los_batch_jacobians(loss, sp_grid, planet, LUTS, bayes_set, solo_absorption=False,
                        initial_intensity=None, lowres=None, pt0=0, n_pts=None, tables=None,
                        lowres_units='same'):
    """Radiances AND their derivatives with respect to every parameter of bayes_set for a batch of
    lines of sight (the `calc_derivatives=True` path of radtran_fast, smm:2837-2881), one fused
    library call per retrieved gas.  Parameter sets whose name is not a gas of the planet get zero
    derivatives.  Returns (rad, jac): lists of SpectralIntensity / lists of per-parameter
    SpectralIntensity (order of bayes_set.params()); with lowres = (centres, widths) the CUDA
    tensors low [n_los][n_chan], jac_low [n_los][n_tot][n_chan]."""
    import torch
    # tables: {set name: (gas_isos, LosSteps, dfrac)} from los_step_tables_device, else the step
    # dictionaries that calc_radtran_steps(calc_derivatives=True) left on every LOS are used
    if tables is None:
        gi, steps = los_step_tables(loss, planet)
    else:
        gi, steps = list(tables.values())[0][:2]
    luts, keep = [], []
    for m, (g, iso) in enumerate(gi):
        im = getattr(planet.gases[g], iso)
        L = LUTS.get((im.mol_name, im.iso))
        if L is not None:
            luts.append(L.device_lut())
            keep.append(m)
    if not luts:
        raise ValueError('no LUT for any gas of the planet in this spectral range')
    if len(keep) != len(gi):
        steps = engine.LosSteps(steps.n_steps, steps.temp, steps.pres, steps.column[keep],
                                None if steps.tvib is None else steps.tvib[keep])
    grid = sp_grid.grid if hasattr(sp_grid, 'grid') else np.asarray(sp_grid)
    n_pts = len(grid) - pt0 if n_pts is None else n_pts
    i0 = None
    if initial_intensity is not None:
        i0 = torch.as_tensor(np.broadcast_to(np.asarray(initial_intensity, dtype=float),
                                             (len(loss), n_pts)).copy(), device="cuda")
    gdev = None
    if lowres is not None:
        gdev = torch.as_tensor(np.ascontiguousarray(grid, dtype=float), device="cuda")
    n_out = n_pts if lowres is None else len(lowres[0])
    rad, blocks = None, []
    for nam in bayes_set.order:
        n_par = bayes_set.sets[nam].n_par
        in_jac = [1 if gi[m][0] == nam else 0 for m in keep]
        if nam not in planet.gases or not any(in_jac):
            blocks.append(torch.zeros((len(loss), n_par, n_out), dtype=torch.float64, device="cuda"))
            continue
        dfrac = (los_jac_tables(loss, bayes_set, nam, steps.n_steps_max) if tables is None
                 else tables[nam][2])
        if lowres is None:
            rad, jac = engine.los_rt_lut_jac(luts, steps, dfrac, gas_in_jac=in_jac, pt0=pt0,
                                             n_pts=n_pts, i0=i0, solo_absorption=solo_absorption)
        else:
            rad, jac = engine.los_rt_lut_jac_lowres(luts, steps, dfrac, gdev, lowres[0], lowres[1],
                                                    gas_in_jac=in_jac, pt0=pt0, n_pts=n_pts, i0=i0,
                                                    solo_absorption=solo_absorption,
                                                    units=lowres_units)
        blocks.append(jac)
    if rad is None:   # no parameter touches a gas with a LUT: plain forward model
        rad = (engine.los_rt_lut(luts, steps, pt0=pt0, n_pts=n_pts, i0=i0,
                                 solo_absorption=solo_absorption) if lowres is None else
               engine.los_rt_lut_lowres(luts, steps, gdev, lowres[0], lowres[1], pt0=pt0,
                                        n_pts=n_pts, i0=i0, solo_absorption=solo_absorption,
                                        units=lowres_units))
    jac = torch.cat(blocks, dim=1)
    if lowres is not None:
        return rad, jac
    rad, jac = rad.cpu().numpy(), jac.cpu().numpy()
    sub = spcl.SpectralGrid(grid[pt0:pt0 + n_pts], units=getattr(sp_grid, 'units', 'cm_1'))
    return ([spcl.SpectralIntensity(rad[i], sub) for i in range(len(loss))],
            [[spcl.SpectralIntensity(jac[i, q], sub) for q in range(jac.shape[1])]
             for i in range(len(loss))])


def fov_weights(pixel_rot=0.0):
    """(dmax, W0, W2) of FOV_integr_1D (smm:3342-3374): the pixel response across the limb is a
    trapezoid w(x) = esse for |x| <= delta, esse*(dmax-|x|)/(dmax-delta) beyond, on [-dmax, dmax]
    (pixel_rot in degrees).  W0 = int w dx, W2 = int x^2 w dx."""
    rot = abs(sbm.rad(pixel_rot))
    dmax = np.sqrt(2.) / 2. * np.cos(np.pi / 4 - rot)
    delta = dmax - np.sin(rot)
    esse = 1 / np.cos(rot)
    W0 = dmax + delta
    W2 = 2. * delta ** 3 / 3.
    if dmax - delta > 0.0:
        W2 += 2. * (dmax ** 4 / 12. - dmax * delta ** 3 / 3. + delta ** 4 / 4.) / (dmax - delta)
    return dmax, esse * W0, esse * W2


def fov_integrate(low, pixel_rot=0.0):
    """FOV integration of [..., 3, n_chan] low / centre / up LOS spectra -> [..., n_chan].

    The reference builds RectBivariateSpline(x = [-dmax, 0, dmax], wavenumbers, kx=2, ky=2) and
    integrates spline(x, ww)*w(x) with scipy quad for every channel (smm:3342-3374).  With three
    nodes and kx = 2 the spline is, at every grid wavenumber, THE parabola through the three LOS
    values, so the integral has the closed form  y1*W0 + (y0 + y2 - 2*y1)/(2 dmax^2) * W2  (the odd
    term integrates to zero against the even weight); quad only approximates it (~1e-8)."""
    low = np.asarray(low, dtype=float)
    dmax, W0, W2 = fov_weights(pixel_rot)
    y0, y1, y2 = low[..., 0, :], low[..., 1, :], low[..., 2, :]
    return y1 * W0 + (y0 + y2 - 2. * y1) / (2. * dmax ** 2) * W2


def FOV_integr_1D(radtrans, pixel_rot=0.0):
    """Reference signature (smm:3342): three SpectralIntensity (low, centre, up LOS) -> the
    FOV-integrated SpectralIntensity."""
    integ_rad = copy.deepcopy(radtrans[0])
    integ_rad.spectrum = fov_integrate(np.array([rad.spectrum for rad in radtrans]), pixel_rot)
    return integ_rad


def make_group_observations(pixels, alt_step=50., alt_first_los=None):
    """A ladder of lines of sight with a fixed step in tangent altitude that stands in for the
    pixels' own LOS (restated from smm:3290-3338, same rules in the same order): pixels are assumed to share the cube and to have close
    tangent latitude / longitude / SZA; the ladder runs from alt_first_los (at most the lowest LOS
    of the lowest pixel) to past the highest LOS of the highest pixel, at the mean tangent
    latitude / longitude, seen from the first pixel's spacecraft position."""
    pixels.sort(key=lambda x: x.limb_tg_alt)
    sim_LOSs = [pix.LOS() for pix in pixels]
    first_los = pixels[0].low_LOS()
    if first_los.get_tangent_altitude() > pixels[0].limb_tg_alt:
        first_los = pixels[0].up_LOS()
    last_los = pixels[-1].up_LOS()
    if last_los.get_tangent_altitude() < pixels[-1].limb_tg_alt:
        last_los = pixels[-1].low_LOS()
    alt_range = [first_los.get_tangent_altitude(), last_los.get_tangent_altitude()]
    if alt_first_los is None or alt_first_los > alt_range[0]:
        alt_first_los = alt_range[0]
    mea_lat = np.mean([pi.limb_tg_lat for pi in pixels])
    mea_lon = np.mean([pi.limb_tg_lon for pi in pixels])
    mea_sza = np.mean([pix.limb_tg_sza for pix in pixels])
    ssp = pixels[0].sub_solar_point()
    spacecraft = sim_LOSs[0].starting_point
    alts = np.arange(alt_first_los, alt_range[1] + alt_step, alt_step)
    LOS_ok = [sbm.LineOfSight(spacecraft, sbm.Coords([mea_lat, mea_lon, alt], s_ref='Spherical'))
              for alt in alts]
    return LOS_ok, alts, [ssp] * len(alts), [mea_sza] * len(alts)


def make_radtran_spline(alts, radtrans):
    """Function of the tangent altitude that interpolates the simulated LOS spectra
    (restated from smm:3377-3396): the same RectBivariateSpline(alts, grid, spectra, kx=2, ky=2) the reference
    builds, evaluated on the spectra's own grid."""
    from scipy.interpolate import RectBivariateSpline as spline2D
    alts = np.array(alts)
    spectrums = np.array([rad.spectrum for rad in radtrans])
    grid = radtrans[0].spectral_grid.grid
    radsample = radtrans[0]
    intens_spl = spline2D(alts, grid, spectrums, kx=2, ky=2)

    def radtran_alt(x):
        res_spe = copy.deepcopy(radsample)
        res_spe.spectrum = np.array(intens_spl(x, grid)).reshape(-1)
        return res_spe

    return radtran_alt


def _pixel_los_altitudes(pix):
    return np.array([lin.get_tangent_point().Spherical()[2]
                     for lin in (pix.low_LOS(), pix.LOS(), pix.up_LOS())])


def check_lines_mols(lines, molecs):
    """Only the lines of the given molecules; for an isotopologue that carries levels, only the
    lines whose upper AND lower level it knows (smm:69-91)."""
    lines_ok = []
    for mol in molecs:
        for iso in mol.all_iso:
            isomol = getattr(mol, iso)
            mine = [lin for lin in lines if lin.Mol == isomol.mol and lin.Iso == isomol.iso]
            if len(isomol.levels) > 0:
                mine = [lin for lin in mine if isomol.has_level(lin.Lo_lev_str)[0]
                        and isomol.has_level(lin.Up_lev_str)[0]]
            lines_ok += mine
    return lines_ok


def keep_levels_wlines(planet, lines):
    """Erases the levels of the planet's isotopologues that no line touches (smm:94-115)."""
    for gas in planet.gases:
        mol = planet.gases[gas]
        for iso in mol.all_iso:
            isomol = getattr(mol, iso)
            iso_lines = [lin for lin in lines if lin.Mol == isomol.mol and lin.Iso == isomol.iso]
            for lev in list(isomol.levels):
                levvo = getattr(isomol, lev)
                if not any(levvo.equiv(lin.Lo_lev_str) or levvo.equiv(lin.Up_lev_str) for lin in iso_lines):
                    isomol.erase_level(lev)


def keep_levels(planet, keep_levels, lines=None):
    """Keeps only the levels listed in keep_levels[(gas, iso)] (smm:133-149)."""
    for gas in planet.gases:
        mol = planet.gases[gas]
        for iso in mol.all_iso:
            isomol = getattr(mol, iso)
            for lev in list(isomol.levels):
                if lev not in keep_levels[(gas, iso)]:
                    isomol.erase_level(lev)


def _simulated_los(pixels, group_observations, alt_step_sims, alt_first_los):
    """(sim_LOSs, alts_sim, ssps, fszas): three LOS per pixel (low, centre, up) with the pixel's
    sub-solar point and tangent SZA (smm:3091-3100), or the altitude ladder (smm:3056-3058)."""
    if group_observations:
        return make_group_observations(pixels, alt_step=alt_step_sims, alt_first_los=alt_first_los)
    sim_LOSs, ssps, fszas = [], [], []
    for pix in pixels:
        sim_LOSs += [pix.low_LOS(), pix.LOS(), pix.up_LOS()]
        ssps += 3 * [pix.sub_solar_point()]
        fszas += 3 * [pix.limb_tg_sza]
    alts_sim = [los.get_tangent_point().Spherical()[2] for los in sim_LOSs]
    return sim_LOSs, alts_sim, ssps, fszas


def _spectral_grid_of(pixels, wn_range, sp_gri):
    """The hi-res grid of a run (smm:3000-3012): sp_gri, or prepare_spe_grid(wn_range), or - with
    neither - the observation's own range widened by two channel widths, converted to cm-1."""
    if sp_gri is not None:
        return sp_gri
    if wn_range is None:
        obs = pixels[0].observation
        g = copy.deepcopy(obs.spectral_grid)
        g.grid[0] -= 2 * obs.bands.spectrum[0]
        g.grid[-1] += 2 * obs.bands.spectrum[-1]
        g.convertto_cm_1()
        wn_range = [g.grid[0], g.grid[-1]]
    return prepare_spe_grid(wn_range).spectral_grid


def _luts_for(inputs, planet, lines, pixels, sim_LOSs, sp_gri, LUTopt):
    LUTopt = dict(LUTopt)
    if 'max_pres' not in LUTopt:   # the deepest point any simulated LOS reaches (the reference looks
        # at the pixels' low LOS only, :3010-3024, and stops with 'Extrapolating in P' when the
        # ladder starts below them)
        LUTopt['max_pres'] = max([planet.atmosphere.calc(p.low_LOS().get_tangent_point(), 'pres')
                                  for p in pixels] +
                                 [planet.atmosphere.calc(sim_LOSs[0].get_tangent_point(), 'pres')])
    gases = list(planet.gases.values())
    PT = calc_PT_couples_atmosphere(lines, gases, planet.atmosphere, **LUTopt)
    return check_and_build_allluts(inputs, sp_gri, lines, gases, PTcouples=PT, LUTopt=LUTopt)


def track_all_levels(planet):
    """{(gas, iso): all its levels} - the `track_levels` argument that follows every level."""
    return dict(((g, iso), list(getattr(planet.gases[g], iso).levels))
                for g in planet.gases for iso in planet.gases[g].all_iso)


def single_radiances(sim_LOSs, sp_gri, planet, LUTS, tables, lowres, lowres_units, factor, obs,
                     track_levels=None, **kwargs):
    """single_rads of radtrans (smm:3176-3186, 3242-3246): per (gas, iso) - and per tracked level
    (gas, iso, lev) - the low-res radiance each LOS receives from THAT emitter alone, absorbed by
    the whole mixture.  The layer source is linear in the emitters, so these contributions add up
    to the total radiance (the reference's own definition lives in the missing radtran_fast).
    One batched device run per emitter, with the spontaneous-emission rows of every other set
    switched off (sr_lut_set_emission_mask)."""
    gi = tables[0] if tables is not None else sim_LOSs[0].radtran_steps['gas_isos']
    keys = []
    for g, iso in gi:
        im = getattr(planet.gases[g], iso)
        if LUTS.get((im.mol_name, im.iso)) is None:
            continue
        keys.append((g, iso))
        if track_levels is not None and (g, iso) in track_levels:
            keys += [(g, iso, lev) for lev in track_levels[(g, iso)]]
    luts = dict(((g, iso), LUTS[(getattr(planet.gases[g], iso).mol_name,
                                 getattr(planet.gases[g], iso).iso)]) for g, iso in gi
                if LUTS.get((getattr(planet.gases[g], iso).mol_name,
                             getattr(planet.gases[g], iso).iso)) is not None)
    out = dict()
    obs_units = getattr(obs, 'units', 'ergscm2') if obs is not None else 'ergscm2'
    try:
        for key in keys:
            for gk, L in luts.items():
                if gk != key[:2]:
                    L.device_lut().set_emission_mask(0)
                elif len(key) == 2:
                    L.device_lut().set_emission_mask(None)
                else:
                    L.device_lut().set_emission_mask(1 << L.set_names().index(key[2]))
            if lowres is None:   # hi-res contributions (LineOfSight.radtran_fast)
                rads = los_batch_radiances(sim_LOSs, sp_gri, planet, LUTS, tables=tables, **kwargs)
                out[key] = dict((los.tag, r) for los, r in zip(sim_LOSs, rads))
                continue
            low = los_batch_radiances(sim_LOSs, sp_gri, planet, LUTS, lowres=lowres, tables=tables,
                                      lowres_units=lowres_units, **kwargs).cpu().numpy() * factor
            out[key] = dict((los.tag, spcl.SpectralIntensity(low[i], obs.spectral_grid, units=obs_units))
                            for i, los in enumerate(sim_LOSs))
    finally:
        for L in luts.values():
            L.device_lut().set_emission_mask(None)
    return out


def radtrans(inputs, planet, lines, pixels, wn_range=None, sp_gri=None, radtran_opt=dict(),
             save_hires=True, save_lowres=True, LUTopt=dict(), test=False, use_tangent_sza=False,
             group_observations=False, invert_LOS_direction=False, nome_inv='1',
             track_levels=None, alt_step_sims=50., alt_first_los=None):
    """Forward model for a list of pixels (smm:2990-3287), batched on the GPU.

    Three lines of sight per pixel (low, centre, up; :3091-3096) or the altitude ladder of
    make_group_observations -> geometry, SZA along every LOS (or the tangent SZA with
    use_tangent_sza) and radtran steps for ALL lines of sight in one library call -> ONE batched
    call for the radiances, reduced to the instrument channels on the device (any observation axis
    hires_to_lowres accepts).  invert_LOS_direction runs the layers in LOS_order='photon'.
    Returns (sims, radtrans, single_rads): `radtrans` = {LOS tag: low-res SpectralIntensity};
    `sims` = per pixel the FOV integral of its three LOS (FOV_integr_1D, :3273-3277), from its own
    LOS or, with group_observations, from the ladder through make_radtran_spline (:3263-3272);
    `single_rads` = {(gas, iso[, lev]): {LOS tag: contribution}} (see single_radiances).
    With inputs['out_dir']: `lowres_radtran_<nome_inv>.pic` (save_lowres) and, with save_hires, the
    hi-res radiances `hires_radtran_<nome_inv>.pic` as [0, {LOS tag: SpectralIntensity}]."""
    lines = check_lines_mols(lines, planet.gases.values())
    pixels = sorted(pixels, key=lambda p: p.limb_tg_alt)
    sp_gri = _spectral_grid_of(pixels, wn_range, sp_gri)
    sim_LOSs, alts_sim, ssps, fszas = _simulated_los(pixels, group_observations, alt_step_sims,
                                                     alt_first_los)
    for num, los in enumerate(sim_LOSs):
        los.tag = 'LOS{:02d}'.format(num)
    LUTS = _luts_for(inputs, planet, lines, pixels, sim_LOSs, sp_gri, LUTopt)

    # geometry + SZA + radtran steps of ALL lines of sight in one library call; the per-LOS host
    # methods calc_atm_intersections / calc_SZA_along_los / calc_radtran_steps stay on sbm
    tables = los_step_tables_device(sim_LOSs, planet, lines=lines, ssps=ssps, fszas=fszas,
                                    use_tangent_sza=use_tangent_sza,
                                    LOS_order='photon' if invert_LOS_direction else 'radtran',
                                    **radtran_opt)

    obs = pixels[0].observation
    centres, widths, ch_units, factor = observation_channels(obs, sp_gri)
    obs_units = getattr(obs, 'units', 'ergscm2')
    out_dir = inputs.get('out_dir') if isinstance(inputs, dict) else None
    if save_hires and out_dir:
        # the hi-res spectra are wanted on disk: compute them once, convolve them on the device
        import torch
        hi = los_batch_radiances(sim_LOSs, sp_gri, planet, LUTS, tables=tables)
        with open(os.path.join(out_dir, 'hires_radtran_{}.pic'.format(nome_inv)), 'wb') as f:
            pickle.dump([0, dict((los.tag, h) for los, h in zip(sim_LOSs, hi))], f, protocol=-1)
        gdev = torch.as_tensor(np.ascontiguousarray(sp_gri.grid, dtype=float), device="cuda")
        spec = torch.as_tensor(np.array([h.spectrum for h in hi]), device="cuda")
        low = engine.convolve_lowres(gdev, spec, centres, widths, units=ch_units).cpu().numpy() * factor
    else:
        # one call for the whole batch and range: the library cuts it into LOS blocks x wavenumber
        # chunks itself (the reference's n_split loop, smm:3190, was a host-memory workaround)
        low = los_batch_radiances(sim_LOSs, sp_gri, planet, LUTS, lowres=(centres, widths),
                                  tables=tables, lowres_units=ch_units)
        low = low.cpu().numpy() * factor
    radtrans_out = dict()
    for i, los in enumerate(sim_LOSs):
        radtrans_out[los.tag] = spcl.SpectralIntensity(low[i], obs.spectral_grid, units=obs_units)
    single_rads = single_radiances(sim_LOSs, sp_gri, planet, LUTS, tables, (centres, widths),
                                   ch_units, factor, obs, track_levels)
    if group_observations:   # spectra at the pixels' LOS altitudes from the ladder (smm:3263-3272)
        radtran_spline = make_radtran_spline(alts_sim, [radtrans_out[los.tag] for los in sim_LOSs])
        sims = []
        for pix in pixels:
            three = np.array([radtran_spline(al).spectrum for al in _pixel_los_altitudes(pix)])
            sims.append(spcl.SpectralIntensity(
                fov_integrate(three, getattr(pix, 'pixel_rot', 0.0) or 0.0), obs.spectral_grid,
                units=obs_units))
    else:
        sims = [spcl.SpectralIntensity(fov_integrate(low[3 * k:3 * k + 3],
                                                     getattr(pixels[k], 'pixel_rot', 0.0) or 0.0),
                                       obs.spectral_grid, units=obs_units) for k in range(len(pixels))]
    if out_dir and save_lowres:
        with open(os.path.join(out_dir, 'lowres_radtran_{}.pic'.format(nome_inv)), 'wb') as f:
            pickle.dump([pixels, sims, sim_LOSs, radtrans_out, single_rads], f, protocol=-1)
    return sims, radtrans_out, single_rads


def forward_jacobian_limb(inputs, planet, lines, bayes_set, pixels, wn_range=None, sp_gri=None,
                          radtran_opt=dict(), save_lowres=True, LUTopt=dict(),
                          use_tangent_sza=False, group_observations=False, nome_inv='1',
                          invert_LOS_direction=False, alt_step_sims=50., alt_first_los=None,
                          LUTS=None):
    """ONE forward + Jacobian evaluation of the fast limb retrieval (the body of the iteration
    loop, smm:2626-2940), batched on the GPU: the current VMR profiles of bayes_set are installed
    on the planet (:2626-2627), all lines of sight of all pixels go through the radtran steps with
    derivative columns and ONE fused library call per retrieved gas returns low-res radiances and
    derivative spectra; both are FOV-integrated per pixel (:2934-2940) and the derivatives are
    stored on the parameters (`par.store_deriv`), so that `bayes_set.build_jacobian()` gives the
    Jacobian inversion_algebra consumes.  Returns (sims, radtrans, derivs) with
    derivs[(LOS tag, nameset, key)] = low-res derivative SpectralIntensity."""
    lines = check_lines_mols(lines, planet.gases.values())
    pixels = sorted(pixels, key=lambda p: p.limb_tg_alt)
    sp_gri = _spectral_grid_of(pixels, wn_range, sp_gri)
    for gas in bayes_set.sets.keys():
        if gas in planet.gases:
            planet.gases[gas].add_clim(bayes_set.sets[gas].profile())
    sim_LOSs, alts_sim, ssps, fszas = _simulated_los(pixels, group_observations, alt_step_sims,
                                                     alt_first_los)
    for num, los in enumerate(sim_LOSs):
        los.tag = 'LOS{:02d}'.format(num)
    if LUTS is None:
        LUTS = _luts_for(inputs, planet, lines, pixels, sim_LOSs, sp_gri, LUTopt)

    # geometry, radtran steps and derivative columns of all LOS on the device, one call per
    # retrieved gas (the step tables themselves are identical between the calls)
    kw = dict(lines=lines, ssps=ssps, fszas=fszas, use_tangent_sza=use_tangent_sza,
              LOS_order='photon' if invert_LOS_direction else 'radtran')
    kw.update(radtran_opt)
    tables = dict()
    for nam in bayes_set.order:
        if nam in planet.gases:
            tables[nam] = los_step_tables_device(sim_LOSs, planet, bayes_set=bayes_set,
                                                 set_name=nam, **kw)
    if not tables:
        tables[None] = los_step_tables_device(sim_LOSs, planet, bayes_set=bayes_set, **kw)

    obs = pixels[0].observation
    centres, widths, ch_units, factor = observation_channels(obs, sp_gri)
    obs_units = getattr(obs, 'units', 'ergscm2')
    low, jlow = los_batch_jacobians(sim_LOSs, sp_gri, planet, LUTS, bayes_set,
                                    lowres=(centres, widths), tables=tables, lowres_units=ch_units)
    low, jlow = low.cpu().numpy() * factor, jlow.cpu().numpy() * factor
    radtrans_out, derivs = dict(), dict()
    pars = bayes_set.params()
    for i, los in enumerate(sim_LOSs):
        radtrans_out[los.tag] = spcl.SpectralIntensity(low[i], obs.spectral_grid, units=obs_units)
        for q, par in enumerate(pars):
            if not los.involved_retparams.get((par.nameset, par.key), False):
                jlow[i, q] = 0.0                                   # zeroder, smm:2871-2872
            else:
                par.set_used()
            derivs[(los.tag, par.nameset, par.key)] = spcl.SpectralIntensity(
                jlow[i, q], obs.spectral_grid, units=obs_units)
    sims = []
    if group_observations:   # radiances and derivatives interpolated from the ladder (:2904-2929)
        radtran_spline = make_radtran_spline(alts_sim, [radtrans_out[los.tag] for los in sim_LOSs])
        deriv_splines = [make_radtran_spline(alts_sim, [derivs[(los.tag, par.nameset, par.key)]
                                                        for los in sim_LOSs]) for par in pars]
    for k, pix in enumerate(pixels):
        rot = getattr(pix, 'pixel_rot', 0.0) or 0.0
        if group_observations:
            alts_pix = _pixel_los_altitudes(pix)
            three = np.array([radtran_spline(al).spectrum for al in alts_pix])
            ders = [np.array([spl(al).spectrum for al in alts_pix]) for spl in deriv_splines]
        else:
            three = low[3 * k:3 * k + 3]
            ders = [jlow[3 * k:3 * k + 3, q] for q in range(len(pars))]
        sims.append(spcl.SpectralIntensity(fov_integrate(three, rot), obs.spectral_grid, units=obs_units))
        for q, par in enumerate(pars):
            par.store_deriv(spcl.SpectralIntensity(fov_integrate(ders[q], rot), obs.spectral_grid,
                                                   units=obs_units), num=k)
    for par in pars:
        par.hires_deriv = None
    out_dir = inputs.get('out_dir') if isinstance(inputs, dict) else None
    if out_dir and save_lowres:
        with open(os.path.join(out_dir, 'lowres_radtran_{}.pic'.format(nome_inv)), 'wb') as f:
            pickle.dump([pixels, sims, sim_LOSs, radtrans_out, dict()], f, protocol=-1)
    return sims, radtrans_out, derivs


# ---------------------------------------------------------------------------------------------
# the update step of the retrieval (smm:3399-3469): tiny dense algebra on the host, kept so that
# inversion_fast_limb runs like the reference's; not part of the GPU hot path
# ---------------------------------------------------------------------------------------------
def genvec(obs, sims, noise, masks=None):
    """Concatenated observation / simulation / noise vectors of all pixels (smm:3399-3425)."""
    cat = lambda xs: np.concatenate([np.asarray(x.spectrum, dtype=float) for x in xs])   # noqa: E731
    obs_vec, sim_vec, noi_vec = cat(obs), cat(sims), cat(noise)
    if masks is not None:
        keep = np.concatenate([np.asarray(m) for m in masks]).astype(bool)
        obs_vec, sim_vec, noi_vec = obs_vec[keep], sim_vec[keep], noi_vec[keep]
    return obs_vec, sim_vec, noi_vec


def chicalc(obs, sims, noise, masks, n_ret):
    """Reduced chi square (smm:3428-3433)."""
    obs_vec, sim_vec, noi_vec = genvec(obs, sims, noise, masks=masks)
    return np.sum(((obs_vec - sim_vec) / noi_vec) ** 2) / (len(obs_vec) - n_ret)


def inversion_algebra(obs, sims, noise, bayes_set, lambda_LM=0.1, L1_reg=False, masks=None):
    """Bayesian optimal estimation, one Levenberg-Marquardt step (smm:3435-3469):
    dx = (K^T Sy^-1 K + Sa^-1 + lambda diag(.))^-1 [K^T Sy^-1 (y - F) + Sa^-1 (xa - x)]."""
    from numpy.linalg import inv
    jac = bayes_set.build_jacobian(masks=masks)
    xi = bayes_set.param_vector()
    obs_vec, sim_vec, noi_vec = genvec(obs, sims, noise, masks=masks)
    KtSy = jac.T * (1.0 / noi_vec ** 2)      # K^T Sy^-1 for the diagonal Sy (the reference forms inv(S_y), :3445-3452)
    G_inv = KtSy @ jac
    Sa_inv = inv(bayes_set.VCM_apriori())
    S_inv = G_inv + Sa_inv
    SxLM = inv(S_inv + lambda_LM * np.diag(np.diag(S_inv)))
    S_x = inv(S_inv)
    deltax = SxLM @ (KtSy @ (obs_vec - sim_vec) + Sa_inv @ (bayes_set.apriori_vector() - xi))
    bayes_set.update_params(deltax)
    bayes_set.store_avk(S_x @ G_inv)
    bayes_set.store_VCM(S_x)


def inversion(inputs, planet, lines, bayes_set, pixels, wn_range=None, chi_threshold=0.01, max_it=10,
              lambda_LM=0.1, L1_reg=False, radtran_opt=dict(), useLUTs=True, debugfile=None,
              save_hires=False, save_lowres=True, LUTopt=dict(), test=False, g3D=False):
    """The per-LOS retrieval of the reference (smm:2422-2595; radtran_test_CO.py:194): every pixel
    is simulated by three LineOfSight.radtran calls (low / centre / up LOS, hi-res radiance and
    hi-res derivative spectra each), convolved with hires_to_lowres and FOV-integrated; then chi,
    the convergence tests and the Levenberg-Marquardt step.  The reference returns nothing; here
    (chi, obs, sims, bayes_set) is returned for convenience.  useLUTs=False runs the line-by-line
    path without look-up tables (LineOfSight.radtran)."""
    lines = check_lines_mols(lines, planet.gases.values())
    sp_gri = _spectral_grid_of(pixels, wn_range, None)
    wn_range = [sp_gri.grid[0], sp_gri.grid[-1]]
    for gas in bayes_set.sets.keys():
        if gas in planet.gases:
            planet.gases[gas].add_clim(bayes_set.sets[gas].profile())
    LUTS = None
    if useLUTs:
        gases = list(planet.gases.values())
        PT = calc_PT_couples_atmosphere(lines, gases, planet.atmosphere, **LUTopt)
        LUTS = check_and_build_allluts(inputs, sp_gri, lines, gases, PTcouples=PT, LUTopt=LUTopt)
    obs = [pix.observation for pix in pixels]
    masks = [pix.observation.mask for pix in pixels]
    noise = [pix.observation.noise for pix in pixels]
    out_dir = inputs.get('out_dir') if isinstance(inputs, dict) else None
    sims = [None] * len(pixels)
    chi_old = chi = None
    for num_it in range(max_it):
        hires = []
        for num, pix in enumerate(pixels):
            ssp = pix.sub_solar_point() if g3D else None
            widths = pix.observation.bands.spectrum
            three, ders = [], []
            for linea in (pix.low_LOS(), pix.LOS(), pix.up_LOS()):
                res = linea.radtran(wn_range, planet, lines, cartLUTs=inputs.get('cart_LUTS'),
                                    cartDROP=out_dir, calc_derivatives=True, bayes_set=bayes_set,
                                    LUTS=LUTS, useLUTs=useLUTs, radtran_opt=radtran_opt, g3D=g3D,
                                    sub_solar_point=ssp, sp_grid=sp_gri)
                three.append(res[0].hires_to_lowres(pix.observation, spectral_widths=widths))
                row = []
                for par in bayes_set.params():
                    if par.not_involved or par.hires_deriv is None:
                        row.append(pix.observation * 0.0)
                    else:
                        row.append(par.hires_deriv.hires_to_lowres(pix.observation, spectral_widths=widths))
                ders.append(row)
                if linea is not None and len(three) == 2:
                    hires.append([num, pix.limb_tg_alt, res])
            sims[num] = FOV_integr_1D(three, pix.pixel_rot)
            for q, par in enumerate(bayes_set.params()):
                par.store_deriv(FOV_integr_1D([d[q] for d in ders], pix.pixel_rot), num=num)
        if save_hires and out_dir:
            with open(os.path.join(out_dir, 'hires_radtran.pic'), 'wb') as f:
                for h in hires:
                    pickle.dump(h, f, protocol=-1)
        chi = chicalc(obs, sims, noise, masks, bayes_set.n_tot)
        if chi_old is not None and (abs(chi - chi_old) / chi_old < chi_threshold or chi > chi_old):
            return chi, obs, sims, bayes_set
        chi_old = chi
        inversion_algebra(obs, sims, noise, bayes_set, lambda_LM=lambda_LM, L1_reg=L1_reg, masks=masks)
        for par in bayes_set.params():
            par.hires_deriv = None
        if debugfile is not None:
            pickle.dump([num_it, obs, sims, bayes_set], debugfile)
        for gas in bayes_set.sets.keys():
            if gas in planet.gases:
                planet.gases[gas].add_clim(bayes_set.sets[gas].profile())
    return chi, obs, sims, bayes_set


def inversion_fast_limb(inputs, planet, lines, bayes_set, pixels, wn_range=None, sp_gri=None,
                        chi_threshold=0.01, max_it=10, lambda_LM=0.1, L1_reg=False,
                        radtran_opt=dict(), debugfile=None, save_hires=False, save_lowres=True,
                        LUTopt=dict(), test=False, use_tangent_sza=False, group_observations=False,
                        nome_inv='1', solo_simulation=False, invert_LOS_direction=False,
                        alt_step_sims=50., alt_first_los=None, track_levels=None, check_log=None,
                        g3D=None):
    """The fast limb retrieval (smm:2598-2987): up to max_it iterations of forward model +
    analytic Jacobians on the GPU (forward_jacobian_limb), reduced chi square, convergence tests
    (relative change below chi_threshold, or chi rising) and the Levenberg-Marquardt update
    (inversion_algebra); the retrieved VMR profiles are installed on the planet after every step.
    Returns (chi, obs, sims, bayes_set) like the reference; None after one simulation with
    solo_simulation (:2903-2905).  When max_it iterations end without meeting a stopping rule the
    reference falls off its loop and returns None (:2987); here the state after the last
    evaluation is returned instead, so that the work is not lost.  The LUTs
    are built once and reused by every iteration.  save_hires is not supported here (the
    derivative spectra of a batch exist on the device per LOS block only): it raises."""
    if save_hires:
        raise NotImplementedError('inversion_fast_limb(save_hires=True): hi-res radiances and '
                                  'derivatives are reduced to the channels on the device; use '
                                  'radtrans(save_hires=True) or LineOfSight.radtran_fast')
    if track_levels is not None:
        raise NotImplementedError('inversion_fast_limb(track_levels=...): per-level contributions '
                                  'are available from radtrans(track_levels=...)')
    lines = check_lines_mols(lines, planet.gases.values())
    pixels = sorted(pixels, key=lambda p: p.limb_tg_alt)
    sp_gri = _spectral_grid_of(pixels, wn_range, sp_gri)
    obs = [pix.observation for pix in pixels]
    masks = [pix.observation.mask for pix in pixels]
    noise = [pix.observation.noise for pix in pixels]
    for gas in bayes_set.sets.keys():
        if gas in planet.gases:
            planet.gases[gas].add_clim(bayes_set.sets[gas].profile())
    sim_LOSs = _simulated_los(pixels, group_observations, alt_step_sims, alt_first_los)[0]
    LUTS = _luts_for(inputs, planet, lines, pixels, sim_LOSs, sp_gri, LUTopt)
    chi_old = None
    for num_it in range(max_it):
        sims, radtrans_out, derivs = forward_jacobian_limb(
            inputs, planet, lines, bayes_set, pixels, sp_gri=sp_gri, radtran_opt=radtran_opt,
            save_lowres=save_lowres, use_tangent_sza=use_tangent_sza,
            group_observations=group_observations, nome_inv=nome_inv,
            invert_LOS_direction=invert_LOS_direction, alt_step_sims=alt_step_sims,
            alt_first_los=alt_first_los, LUTS=LUTS)
        if solo_simulation:
            return None
        chi = chicalc(obs, sims, noise, masks, bayes_set.n_used_par())
        if debugfile is not None:
            pickle.dump([num_it, chi, obs, sims, bayes_set, radtrans_out, derivs], debugfile)
        if check_log is not None:
            check_log.write('Iteration {:2d}: chi is {:8.3f}\n'.format(num_it, chi))
        if chi_old is not None:
            if abs(chi - chi_old) / chi_old < chi_threshold:
                if check_log is not None:
                    check_log.write('Finished!\n')
                return chi, obs, sims, bayes_set
            if chi > chi_old:
                if check_log is not None:
                    check_log.write('Chi has raised.. Finished!\n')
                return chi, obs, sims, bayes_set
        chi_old = chi
        inversion_algebra(obs, sims, noise, bayes_set, lambda_LM=lambda_LM, L1_reg=L1_reg, masks=masks)
        if check_log is not None:
            check_log.write(('Params: ' + len(bayes_set.params()) * '{:9.2e} ' + '\n').format(
                *[par.value for par in bayes_set.params()]))
        for gas in bayes_set.sets.keys():
            if gas in planet.gases:
                planet.gases[gas].add_clim(bayes_set.sets[gas].profile())
    return chi, obs, sims, bayes_set
