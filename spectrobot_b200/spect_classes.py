"""`spect_classes` for the hot path, Python 3, backed by libspectrobot.so.

Same module-level names, class names, method signatures and attribute names as the reference's
spect_classes.py for everything the line-by-line path touches (SURVEY 8b "who calls it"), so that
code written against the reference keeps working:

    SpectLine, SpectralGrid, SpectralObject, SpectralIntensity, SpectralGcoeff,
    calc_shapes_lines, closest_grid, Lorenz_width, Doppler_width, MakeShape, convert_to_atm,
    ImportPartitionSumTable, CalcPartitionSum, CalcStrength_at_T, Einstein_A_to_B, ...

What differs is where the arithmetic runs: per-line Python loops, fork/Queue fan-out and the three
f2py modules are replaced by batched calls into the CUDA library (engine.LineSet), and the level
book-keeping by string comparison (spect_classes.py:1304-1313) is resolved once into integer set
ids (`line_table`).  `read_line_database` parses HITRAN / "gbb" fixed-width files (SURVEY 8f row
3).  The host-side helpers of SpectralObject (slicing, arithmetic, re-gridding, degrade_grid /
degrade_grid2, the older convolve_to_grid chain) and the scalar shape / black-body helpers are
plain NumPy restatements pinned by tests/test_ref_golden2.py.  Not here: plotting (`plot`,
`norm_plot`) and `degrade_grid3`, which raises NameError in the reference.
"""
import copy
import math as mt

import numpy as np
import scipy.constants as const

from . import engine, fparts_mod, lineshape
from . import spect_base_module as sbm

n_threads = 8            # kept for signature compatibility; the GPU path ignores it
imxsig = 13010           # parameters.inc:65
imxlines = 40000         # parameters.inc:64
imxsig_long = 2000000    # parameters.inc:64

T_ref = 296.0
hpa_to_atm = 0.00098692326671601
# cgs constants "as in HITRAN", taken from the installed scipy like the reference does (:44-47)
h_cgs = const.physical_constants['Planck constant'][0] * 1.e7
c_cgs = const.c * 1.e2
k_cgs = const.physical_constants['Boltzmann constant'][0] * 1.e7
c2 = h_cgs * c_cgs / k_cgs

cose = ('Mol', 'Iso', 'Freq', 'Strength', 'A_coeff', 'Air_broad', 'Self_broad', 'E_lower',
        'T_dep_broad', 'P_shift', 'Up_lev_str', 'Lo_lev_str', 'Q_num_up', 'Q_num_lo')
cose_hit = cose + ('others', 'g_up', 'g_lo')
cose_mas = ('Up_lev_id', 'Lo_lev_id', 'Up_lev', 'Lo_lev')
CTYPES = ['sp_emission', 'ind_emission', 'absorption']


# ---------------------------------------------------------------------------------------------
# scalar physics helpers (reference: spect_classes.py:1713-1878, 1967-2036)
# ---------------------------------------------------------------------------------------------
def convert_to_atm(Pres, units='hPa'):
    if units == 'hPa':
        return Pres * hpa_to_atm
    raise ValueError('unknown pressure units ' + str(units))


def Lorenz_width(Temp, Pres_atm, T_dep_broad, Air_broad, Self_broad=0.0, Self_pres_atm=0.0):
    """Pressure HWHM in cm-1 (:1967-1972)."""
    return (T_ref / Temp) ** T_dep_broad * (Air_broad * (Pres_atm - Self_pres_atm)
                                            + Self_broad * Self_pres_atm)


def Doppler_width(Temp, MM, wn_0):
    """Doppler HWHM in cm-1 (:1976-1986)."""
    return wn_0 / c_cgs * mt.sqrt(2 * const.Avogadro * k_cgs * Temp * mt.log(2.0) / MM)


def Boltz_ratio_nodeg(wavenumber, temp):
    return np.exp(-c2 * wavenumber / temp)


def Einstein_A_to_B(A_coeff, wavenumber, units='cm3ergcm2'):
    """B21 from A21 in the HITRAN cgs convention (:1736-1754)."""
    if units != 'cm3ergcm2':
        raise ValueError('only the cm3ergcm2 convention is supported')
    return A_coeff / (2 * h_cgs * c_cgs ** 2 * wavenumber ** 3)


def Einstein_B21_to_A(B21, wavenumber, units='cm3ergcm2'):
    return B21 * (2 * h_cgs * c_cgs ** 2 * wavenumber ** 3)


def Einstein_B21_to_B12(B_21, g_1, g_2):
    return B_21 * g_2 / g_1


def Einstein_A_to_Gcoeff_spem(line, Temp, E_vib):
    pop_rot = line.g_up * Boltz_ratio_nodeg(line.E_lower + line.Freq - E_vib, Temp)
    return h_cgs * c_cgs * line.Freq * pop_rot * line.A_coeff / (4 * np.pi)


def Einstein_A_to_Gcoeff_indem(line, Temp, E_vib):
    pop_rot = line.g_up * Boltz_ratio_nodeg(line.E_lower + line.Freq - E_vib, Temp)
    B21 = Einstein_A_to_B(line.A_coeff, line.Freq)
    return h_cgs * c_cgs * line.Freq * pop_rot * B21 / (4 * np.pi)


def Einstein_A_to_Gcoeff_abs(line, Temp, E_vib):
    pop_rot = line.g_lo * Boltz_ratio_nodeg(line.E_lower - E_vib, Temp)
    B12 = Einstein_B21_to_B12(Einstein_A_to_B(line.A_coeff, line.Freq), line.g_lo, line.g_up)
    return h_cgs * c_cgs * line.Freq * pop_rot * B12 / (4 * np.pi)


def Einstein_A_to_LineStrength_nonLTE(A_coeff, wavenumber, E_lower, T_vib_lower, T_vib_upper,
                                      g_lower, g_upper, Q_part, iso_ab=1.0):
    B21 = Einstein_A_to_B(A_coeff, wavenumber)
    B12 = Einstein_B21_to_B12(B21, g_lower, g_upper)
    pop_lo = g_lower * Boltz_ratio_nodeg(E_lower, T_vib_lower) / Q_part
    pop_up = g_upper * Boltz_ratio_nodeg(E_lower + wavenumber, T_vib_upper) / Q_part
    return iso_ab * h_cgs * c_cgs * wavenumber / (4 * np.pi) * (pop_lo * B12 - pop_up * B21)


def ImportPartitionSumTable(mol, iso):
    """(gi, T grid, Q grid) from the TIPS-2003 tables (:1680-1689, fparts_mod.f:33-295)."""
    return fparts_mod.bd_tips_2003(mol, iso)


def CalcPartitionSum(mol, iso, temp=296.0):
    """4-point Lagrange interpolation of the TIPS table (:1692-1710)."""
    return engine.partition_sum(mol, iso, temp)


def CalcStrength_at_T(mol, iso, S_ref, w, E_low, T, T_ref=296.):
    """HITRAN temperature scaling of a line strength (:1713-1733)."""
    def fu_exp(E, w_, T_):
        return np.exp(-c2 * E / T_) * (1 - np.exp(-c2 * w_ / T_))
    return S_ref * CalcPartitionSum(mol, iso, T_ref) / CalcPartitionSum(mol, iso, T) * \
        fu_exp(E_low, w, T) / fu_exp(E_low, w, T_ref)


def closest_grid(wn_arr, wn_0):
    """Index and value of the grid point closest to wn_0 (:1937-1943)."""
    g = wn_arr.grid if hasattr(wn_arr, 'grid') else np.asarray(wn_arr)
    ind = int(np.argmin(np.abs(g - wn_0)))
    return ind, g[ind]


def gaussian(arr, mu, sig):
    return 1 / (sig * np.sqrt(2. * np.pi)) * np.exp(-0.5 * ((arr - mu) / sig) ** 2)


def MakeShape(wn_arr, wn_0, lw, dw, Strength=1.0):
    """Normalised Voigt profile on a 13010-point grid through the humliv_bb drop-in (:1990-2008)."""
    fac = dw * mt.sqrt(np.pi / mt.log(2.0))
    y = lineshape.humliv_bb(wn_arr.grid, 1, len(wn_arr.grid), wn_0, lw,
                            dw / mt.sqrt(mt.log(2.0)))
    return SpectralObject(Strength * y / fac, wn_arr)


def closest_grid_ext(wn_arr, wn_0):
    """closest_grid for a wavenumber that may lie outside the grid: the grid is continued with its
    own step beyond the end that wn_0 passes (:1945-1964).  The index is that of the continued
    grid (negative below the start)."""
    g, step = wn_arr.grid, wn_arr.step()
    if (g[0] <= wn_0 <= g[-1]) or sbm.isclose(wn_0, g[0]) or sbm.isclose(wn_0, g[-1]):
        ind = int(np.argmin(np.abs(g - wn_0)))
        return ind, g[ind]
    if wn_0 < g[0]:
        ext = np.arange(g[0], wn_0 - step, -step)
        ind = int(np.argmin(np.abs(ext - wn_0)))
        return -ind, ext[ind]
    ext = np.arange(g[-1], wn_0 + step, step)
    ind = int(np.argmin(np.abs(ext - wn_0)))
    return len(g) + ind, ext[ind]


def Lorentz_shape(wn, wn_0, lw):
    """Normalised Lorentzian of HWHM lw (:1906-1913)."""
    return 1 / np.pi * lw / (lw ** 2 + (wn - wn_0) ** 2)


def Doppler_shape(wn, wn_0, dw):
    """Normalised Gaussian of HWHM dw (:1916-1923)."""
    return mt.sqrt(mt.log(2.0) / (np.pi * dw ** 2)) * np.exp(-(wn - wn_0) ** 2 * mt.log(2.0) / dw ** 2)


def MakeShape_py(wn_arr, wn_0, lw, dw, Strength=1.0):
    """Voigt profile as the discrete convolution of the two shapes on the grid (:2011-2026); the
    reference's pure-Python cross-check of MakeShape."""
    lor = Lorentz_shape(wn_arr.grid - wn_0, 0.0, lw)
    dop = Doppler_shape(wn_arr.grid - wn_0, 0.0, dw)
    return SpectralObject(np.convolve(dop, lor, mode='same') * wn_arr.step() * Strength, wn_arr)


def Einstein_A_to_LineStrength_hitran(A_coeff, wavenumber, temp, Q_part, g_upper, E_lower, iso_ab=1.0):
    """LTE line strength in cm-1/(molecule cm-2) from the Einstein A (:1856-1863)."""
    return iso_ab * A_coeff * g_upper * np.exp(-c2 * E_lower / temp) * \
        (1 - np.exp(-c2 * wavenumber / temp)) / (8 * np.pi * c_cgs * wavenumber ** 2 * Q_part)


def Boltz_pop_at_T(wavenumber, temp, g_level, Q_part):
    """LTE population of a level (:1866-1873)."""
    return g_level * Boltz_ratio_nodeg(wavenumber, temp) / Q_part


def alpha_nlte(Freq, Temp, r1, r2):
    """Non-LTE correction of the absorption strength for population ratios r1 (lower), r2 (upper)
    (:1485-1488)."""
    Gm = Boltz_ratio_nodeg(Freq, Temp)
    return r1 * (1 - Gm * r2 / r1) / (1 - Gm)


def Calc_BB_single(nu, T):
    """Planck function in erg/(s cm2 sr cm-1) at wavenumber nu (:1895-1903)."""
    return 2 * h_cgs * c_cgs ** 2 * nu ** 3 / (np.exp(c2 * nu / T) - 1)


def Calc_BB(spectral_grid, T, units='ergscm2'):
    """SpectralIntensity with the Planck spectrum at T on a cm-1 grid (:1881-1892)."""
    out = SpectralIntensity(Calc_BB_single(spectral_grid.grid, T), spectral_grid, units='ergscm2')
    if units != 'ergscm2':
        out.convertto(units)
    return out


def BB(T, w):
    """Black body in nW/(cm2 sr cm-1) with the reference's rounded constants (:2080-2092)."""
    return 1.1904e-3 * w ** 3 / (mt.exp(w * 1.4388 / T) - 1)


def BB_erg(T, w):
    """Black body in erg/(s cm2 sr cm-1), rounded constants (:2095-2107)."""
    return 1.1904e-5 * w ** 3 / (mt.exp(w * 1.4388 / T) - 1)


def BB_nm(T, w):
    """Black body in W/(m2 sr nm) at wavelength w in nm; Wien form for T*w <= 5e5 (:2110-2125)."""
    if T * w > 5e5:
        return 1.1904 * mt.pow(1.e4 / w, 5) / (mt.exp(1.e7 / w * 1.4388 / T) - 1)
    return 1.1904 * mt.pow(1.e4 / w, 5) * mt.exp(-1.e7 / w * 1.4388 / T)


def convert_cm_1_to_J(w):
    return const.h * 1.e2 * const.c * w


def convert_J_to_eV(J):
    """Joule -> eV.  (The reference's body returns its argument unchanged, :2047-2049, and its
    convert_cm_1_to_eV raises NameError, :2043-2045; both are given their evident meaning.)"""
    return J / const.eV


def convert_cm_1_to_eV(w):
    return convert_J_to_eV(convert_cm_1_to_J(w))


def read_mw_list(db_cart, nome='mw_list.dat'):
    """(n, tags, [w1, w2] ranges) of a microwindow list file (:1465-1482): the count on the first
    line, then `index tag w1 w2` per line."""
    with open(db_cart + nome, 'r') as f:
        n_mws = int(f.readline())
        tags, ranges = [], []
        for _ in range(n_mws):
            tok = f.readline().split()
            tags.append(tok[1])
            ranges.append([float(tok[2]), float(tok[3])])
    return n_mws, tags, ranges


# ---------------------------------------------------------------------------------------------
# SpectLine
# ---------------------------------------------------------------------------------------------
class SpectLine(object):
    """One spectral line with the HITRAN fields of `cose_hit` (:56-351)."""

    def __init__(self, linea, nomi=None):
        if nomi is None:
            if isinstance(linea, dict):
                nomi = tuple(linea.keys())
            elif isinstance(linea, np.void):
                nomi = linea.dtype.names
            else:
                raise ValueError('Missing names for line quantities')
        else:
            linea = dict(zip(nomi, linea))
        for nome in nomi:
            setattr(self, nome, linea[nome])
        for nome in cose_mas + cose_hit:
            if not hasattr(self, nome):
                setattr(self, nome, None)
        self.E_vib_up = None
        self.E_vib_lo = None

    def Print(self, ofile=None):
        """Short listing of the line (:93-98); the file form leaves the rotational quanta out."""
        if ofile is None:
            print('{:4d}{:2d}{:8.2f}{:10.3e}{:8.2f}{:15s}{:15s}{:15s}{:15s}{:8.3f}{:8.3f}'.format(
                self.Mol, self.Iso, self.Freq, self.Strength, self.E_lower, self.Up_lev_str,
                self.Lo_lev_str, self.Q_num_up, self.Q_num_lo, self.g_up, self.g_lo))
        else:
            ofile.write('{:4d}{:2d}{:8.2f}{:10.3e}{:8.2f}{:15s}{:15s}{:8.3f}{:8.3f}\n'.format(
                self.Mol, self.Iso, self.Freq, self.Strength, self.E_lower, self.Up_lev_str,
                self.Lo_lev_str, self.g_up, self.g_lo))

    def Print_hitran(self, ofile=None):
        """The line as a HITRAN-like record with the reference's own field formats (:100-109);
        format_line_record writes the fixed-width layout read_line_database reads back."""
        oth = '' if self.others is None else self.others
        stringa = ('{:2d}{:1d}{:12.6f}{:10.3e}{:10.3e}{:5.3f}{:5.3f}{:10.4f}{:4.2f}{:8.5f}{:15s}{:15s}'
                   '{:15s}{:15s}{:19s}{:7.2f}{:7.2f}').format(
            self.Mol, self.Iso, self.Freq, self.Strength, self.A_coeff, self.Air_broad,
            self.Self_broad, self.E_lower, self.T_dep_broad, self.P_shift, self.Up_lev_str,
            self.Lo_lev_str, self.Q_num_up, self.Q_num_lo, oth, self.g_up, self.g_lo)
        if ofile is None:
            print(stringa)
        else:
            ofile.write(stringa + '\n')
        return stringa

    def ShowCalc(self, T, P=1, nlte_ratio=1):
        pass

    def CalcStrength(self, T):
        return CalcStrength_at_T(self.Mol, self.Iso, self.Strength, self.Freq, self.E_lower, T)

    def CalcStrength_from_Strength(self, Temp, Q_part=None, iso_ab=None, isomolec=None,
                                   T_vib_lower=None, T_vib_upper=None, E_vib_lo=None, E_vib_up=None):
        """(absorption, emission) strengths in non-LTE from the tabulated strength at 296 K
        (:256-288): S(T) x alpha_nlte(r1, r2) and S(T) x r2 x BB_erg(T, nu), r = non-LTE / LTE
        population ratio of the lower / upper vibrational level."""
        T_vib_lower = Temp if T_vib_lower is None else T_vib_lower
        T_vib_upper = Temp if T_vib_upper is None else T_vib_upper
        if E_vib_lo is None and E_vib_up is None:
            E_vib_lo, E_vib_up = self.E_vib_lo, self.E_vib_up
        r1 = sbm.vibtemp_to_ratio(E_vib_lo, T_vib_lower, Temp)
        r2 = sbm.vibtemp_to_ratio(E_vib_up, T_vib_upper, Temp)
        S = self.CalcStrength(Temp)
        return S * alpha_nlte(self.Freq, Temp, r1, r2), S * r2 * BB_erg(Temp, self.Freq)

    def minimal_level_string_up(self):
        return sbm.extract_quanta_HITRAN(self.Mol, self.Iso, self.Up_lev_str)[0]

    def minimal_level_string_lo(self):
        return sbm.extract_quanta_HITRAN(self.Mol, self.Iso, self.Lo_lev_str)[0]

    def LinkToMolec(self, isomolec):
        """Sets Up_lev_id / Lo_lev_id / E_vib_up / E_vib_lo from the IsoMolec levels (:122-150)."""
        if isomolec is None:
            return False
        self.Up_lev_id = self.Lo_lev_id = self.E_vib_up = self.E_vib_lo = None
        up, lo = self.minimal_level_string_up(), self.minimal_level_string_lo()
        for lev in isomolec.levels:
            Level = getattr(isomolec, lev)
            ms = Level.minimal_level_string()
            if ms == up:
                self.Up_lev_id, self.E_vib_up = lev, Level.energy
            elif ms == lo:
                self.Lo_lev_id, self.E_vib_lo = lev, Level.energy
        return self.Up_lev_id is not None and self.Lo_lev_id is not None

    def Einstein_A_to_B(self):
        return Einstein_A_to_B(self.A_coeff, self.Freq)

    def CheckWidths(self, Temp, Pres, MM):
        """(Doppler HWHM, Lorentz HWHM, pressure shift) (:161-171)."""
        Pres_atm = convert_to_atm(Pres)
        return (Doppler_width(Temp, MM, self.Freq),
                Lorenz_width(Temp, Pres_atm, self.T_dep_broad, self.Air_broad),
                self.P_shift * Pres_atm)

    def MakeShapeLine(self, Temp, Pres, grid=None, MM=None, Strength=1.0, verbose=False,
                      keep_memory=False):
        """Voigt shape of this line.  As in the reference the pressure shift is computed but NOT
        applied and only air broadening is used (:187-197)."""
        if MM is None:
            MM = sbm.find_molec_metadata(self.Mol, self.Iso)['iso_MM']
        if grid is None:
            sp_step = 5.e-4
            g = np.arange(-imxsig * sp_step / 2, imxsig * sp_step / 2, sp_step, dtype=float)
            grid = SpectralGrid(g + self.Freq, units='cm_1')
        dw, lw, _ = self.CheckWidths(Temp, Pres, MM)
        shape = MakeShape(grid, self.Freq, lw, dw, Strength=Strength)
        if keep_memory:
            self.shape = shape
        return shape

    def CalcStrength_nonLTE(self, Temp, T_vib_lower, T_vib_upper, Q_part=None):
        if Q_part is None:
            Q_part = CalcPartitionSum(self.Mol, self.Iso, temp=Temp)
        return Einstein_A_to_LineStrength_nonLTE(self.A_coeff, self.Freq, self.E_lower,
                                                 T_vib_lower, T_vib_upper, self.g_lo, self.g_up,
                                                 Q_part)

    def CalcStrength_from_Einstein(self, Temp, Q_part=None, iso_ab=None, isomolec=None,
                                   T_vib_lower=None, T_vib_upper=None):
        """(absorption, emission) strengths from the G coefficients (:219-254)."""
        T_vib_lower = Temp if T_vib_lower is None else T_vib_lower
        T_vib_upper = Temp if T_vib_upper is None else T_vib_upper
        if Q_part is None:
            Q_part = CalcPartitionSum(self.Mol, self.Iso, temp=Temp)
        if iso_ab is None:
            iso_ab = sbm.find_molec_metadata(self.Mol, self.Iso)['iso_ratio']
        G = self.Calc_Gcoeffs(Temp, isomolec=isomolec)
        E_lo = 0.0 if self.E_vib_lo is None else self.E_vib_lo
        E_up = 0.0 if self.E_vib_up is None else self.E_vib_up
        S_ab = (G['absorption'] * Boltz_ratio_nodeg(E_lo, T_vib_lower)
                - G['ind_emission'] * Boltz_ratio_nodeg(E_up, T_vib_upper)) / Q_part
        S_em = G['sp_emission'] * Boltz_ratio_nodeg(E_up, T_vib_upper) / Q_part
        return iso_ab * S_ab, iso_ab * S_em

    def calc_A_coeff_from_strength(self, iso_ab=None, Q_part=None, set_attr=False):
        """Inverse of CalcStrength_from_Einstein at 296 K (:291-309)."""
        Temp = 296.0
        if Q_part is None:
            Q_part = CalcPartitionSum(self.Mol, self.Iso, temp=Temp)
        if iso_ab is None:
            iso_ab = sbm.find_molec_metadata(self.Mol, self.Iso)['iso_ratio']
        B_21 = self.Strength * (4 * np.pi * Q_part) / (
            (Boltz_ratio_nodeg(self.E_lower, Temp) - Boltz_ratio_nodeg(self.E_lower + self.Freq, Temp))
            * h_cgs * c_cgs * self.Freq * self.g_up * iso_ab)
        A = Einstein_B21_to_A(B_21, self.Freq)
        if set_attr:
            self.A_coeff = A
        return A

    def Calc_Gcoeffs(self, Temp, isomolec=None):
        """{'sp_emission', 'ind_emission', 'absorption'} G coefficients (:312-343)."""
        ok = self.LinkToMolec(isomolec)
        e_lo, e_up = (self.E_vib_lo, self.E_vib_up) if ok else (0.0, 0.0)
        if self.A_coeff != 0.0 and self.g_lo != 0.0 and self.g_up != 0.0:
            values = [Einstein_A_to_Gcoeff_spem(self, Temp, e_up),
                      Einstein_A_to_Gcoeff_indem(self, Temp, e_up),
                      Einstein_A_to_Gcoeff_abs(self, Temp, e_lo)]
        else:
            values = [0., 0., 0.]
        self.G_coeffs = dict(zip(CTYPES, values))
        return self.G_coeffs


# ---------------------------------------------------------------------------------------------
# line list -> device line table
# ---------------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------
# line database reader (reference: spect_classes.py:1532-1601) and statistical weights (:1604-1678)
# ---------------------------------------------------------------------------------------------
DB_FORMATS = {
    # field widths of one record, names, converters (reference :1546-1553)
    'gbb': ((2, 1, 12, 10, 10, 6, 6, 10, 4, 8, 15, 15, 15, 15), cose),
    'HITRAN': ((2, 1, 12, 10, 10, 5, 5, 10, 4, 8, 15, 15, 15, 15, 19, 7, 7), cose_hit),
}
_DB_NUMERIC = ('Freq', 'Strength', 'A_coeff', 'Air_broad', 'Self_broad', 'E_lower', 'T_dep_broad',
               'P_shift', 'g_up', 'g_lo')


def parse_line_record(text, db_format='HITRAN'):
    """One fixed-width record -> dict of the reference's field names.  Numbers as np.genfromtxt
    reads them (blank -> nan for floats, -1 for ints), strings kept with their padding."""
    widths, names = DB_FORMATS[db_format]
    rec, pos = dict(), 0
    for w, nam in zip(widths, names):
        tok = text[pos:pos + w]
        pos += w
        if nam in ('Mol', 'Iso'):
            try:
                rec[nam] = int(tok)
            except ValueError:
                rec[nam] = -1
        elif nam in _DB_NUMERIC:
            try:
                rec[nam] = float(tok)
            except ValueError:
                rec[nam] = float('nan')
        else:
            rec[nam] = tok
    return rec


def read_line_database(nome_sp, mol=None, iso=None, up_lev=None, down_lev=None,
                       fraction_to_keep=None, db_format='HITRAN', freq_range=None, n_skip=0,
                       link_to_isomolecs=None, verbose=False):
    """List of SpectLine from a HITRAN2012 (160-character) or "gbb" (MAKE_MW) line file, with the
    reference's selection rules (:1532-1601): optional (mol, iso, upper/lower level string)
    filter, frequency window on a frequency-sorted file (stops at the first line beyond it), zero
    air/self widths replaced by 0.05 / 0.07, and the `fraction_to_keep` strength cut."""
    if db_format not in DB_FORMATS:
        raise ValueError('Allowed values for db_format: {}, {}'.format('gbb', 'HITRAN'))
    linee_ok = []
    with open(nome_sp, 'r') as infi:
        if n_skip == -1:
            sbm.trova_spip(infi)
        else:
            for _ in range(n_skip):
                infi.readline()
        for text in infi:
            text = text.rstrip('\r\n')
            if not text.strip():
                continue
            linea = parse_line_record(text, db_format)
            if verbose:
                print(linea['Mol'], linea['Iso'], linea['Freq'])
            if freq_range is not None:
                if linea['Freq'] < freq_range[0]:
                    continue
                if linea['Freq'] > freq_range[1]:
                    break
            if ((linea['Mol'] == mol or mol is None) and (linea['Iso'] == iso or iso is None) and
                    (linea['Up_lev_str'] == up_lev or up_lev is None) and
                    (linea['Lo_lev_str'] == down_lev or down_lev is None)):
                line = SpectLine(linea)
                if line.Air_broad == 0.0:
                    line.Air_broad = 0.05
                if line.Self_broad == 0.0:
                    line.Self_broad = 0.07
                if link_to_isomolecs is not None:
                    found = [m for m in link_to_isomolecs
                             if m.mol == line.Mol and m.iso == line.Iso]
                    if len(found) > 1:
                        raise ValueError('Multiple isotopologues corresponding to line!')
                    if found:
                        line.LinkToMolec(found[0])
                linee_ok.append(line)
    if fraction_to_keep is not None and linee_ok:
        essesss = np.sort(np.array([lin.Strength for lin in linee_ok]))
        essort = essesss[int(fraction_to_keep * (len(linee_ok) - 1))]
        return [lin for lin in linee_ok if lin.Strength >= essort]
    return linee_ok


def format_line_record(line, db_format='HITRAN'):
    """Inverse of parse_line_record for a SpectLine / dict (used to write test and synthetic
    databases in the HITRAN2012 layout)."""
    get = (lambda k: line[k]) if isinstance(line, dict) else (lambda k: getattr(line, k))
    fmt = {'Mol': '{:2d}', 'Iso': '{:1d}', 'Freq': '{:12.6f}', 'Strength': '{:10.3E}',
           'A_coeff': '{:10.3E}', 'E_lower': '{:10.4f}', 'T_dep_broad': '{:4.2f}',
           'P_shift': '{:8.5f}', 'g_up': '{:7.1f}', 'g_lo': '{:7.1f}'}
    wid = 5 if db_format == 'HITRAN' else 6
    fmt['Air_broad'] = '{:%d.4f}' % wid
    fmt['Self_broad'] = '{:%d.3f}' % wid
    widths, names = DB_FORMATS[db_format]
    out = ''
    for w, nam in zip(widths, names):
        v = get(nam)
        if nam in fmt:
            tok = fmt[nam].format(v)
            if nam in ('Air_broad', 'Self_broad') and tok.startswith('0.'):
                tok = tok[1:] if len(tok) > w else tok      # HITRAN writes .0700
        else:
            tok = '' if v is None else str(v)
        out += tok[:w].rjust(w) if nam in fmt else tok[:w].ljust(w)
    return out


def calc_stat_weights_CH4(Q_num_up, Q_num_lo, formato='HITRAN'):
    """g = gi*gs*(2J+1) with gs = 5, 2, 3 for the A, E, F symmetry species of CH4 (:1604-1635)."""
    gs = dict(A=5, E=2, F=3)
    if formato == 'HITRAN':
        J_up, state_up = int(Q_num_up[2:5]), Q_num_up[5]
        J_lo, state_lo = int(Q_num_lo[2:5]), Q_num_lo[5]
    elif formato == 'hot_bands':
        J_up, J_lo = int(Q_num_up.split()[0]), int(Q_num_lo.split()[0])
        state_up, state_lo = Q_num_up.split()[1][0], Q_num_lo.split()[1][0]
    else:
        raise ValueError('formato {} not recognized'.format(formato))
    return gs[state_up] * (2 * J_up + 1), gs[state_lo] * (2 * J_lo + 1)


def calc_stat_weights_linear_molec(gi, gs, Q_num_up, Q_num_lo, formato='HITRAN'):
    """g = gi*gs*(2J+1) of a linear molecule from the branch symbol and lower-state J of the
    HITRAN local quanta, or from both J of a GEISA record (:1639-1678)."""
    def to_int(tok):
        try:
            return int(tok)
        except ValueError:
            return int(tok[:-1])
    if formato == 'HITRAN':
        coso = Q_num_lo.split()
        J, br = to_int(coso[1]), coso[0]
        q_lo = gs * gi * (2 * J + 1)
        if br == 'Q':
            q_up = gs * gi * (2 * J + 1)
        elif br == 'P':
            q_up = gs * gi * (2 * (J - 1) + 1)
        else:
            q_up = gs * gi * (2 * (J + 1) + 1)
    elif formato == 'GEISA':
        J_up, J_lo = to_int(Q_num_up.split()[0]), to_int(Q_num_lo.split()[0])
        q_up, q_lo = gs * gi * (2 * J_up + 1), gs * gi * (2 * J_lo + 1)
    else:
        raise ValueError('formato {} not recognized'.format(formato))
    return q_up, q_lo


def line_table(lines, isomolec=None):
    """Arrays for engine.LineSet from a list of SpectLine and an sbm.IsoMolec.

    Non-LTE isotopologue (isomolec.levels non-empty): up_set / lo_set are the positions of the
    line's upper / lower level in isomolec.levels, -1 when either is unknown, which drops the line
    exactly like calc_shapes_lines does (:1384-1388).  LTE isotopologue without levels: one set
    'all', E_vib = 0 (:318-321, smm:742-748)."""
    n = len(lines)
    tab = {k: np.empty(n) for k in ("freq", "a_coeff", "air_broad", "t_dep", "e_lower", "g_up",
                                    "g_lo", "e_vib_up", "e_vib_lo")}
    up = np.zeros(n, dtype=np.int32)
    lo = np.zeros(n, dtype=np.int32)
    levels = list(isomolec.levels) if isomolec is not None else []
    minstr = {getattr(isomolec, lev).minimal_level_string(): i for i, lev in enumerate(levels)}
    energy = [getattr(isomolec, lev).energy for lev in levels]
    for i, lin in enumerate(lines):
        tab["freq"][i], tab["a_coeff"][i] = lin.Freq, lin.A_coeff
        tab["air_broad"][i], tab["t_dep"][i] = lin.Air_broad, lin.T_dep_broad
        tab["e_lower"][i] = lin.E_lower
        tab["g_up"][i] = 0.0 if lin.g_up is None else lin.g_up
        tab["g_lo"][i] = 0.0 if lin.g_lo is None else lin.g_lo
        if levels:
            u = minstr.get(lin.minimal_level_string_up(), -1)
            l = minstr.get(lin.minimal_level_string_lo(), -1)
            if u < 0 or l < 0:
                u = l = -1
            up[i], lo[i] = u, l
            tab["e_vib_up"][i] = energy[u] if u >= 0 else 0.0
            tab["e_vib_lo"][i] = energy[l] if l >= 0 else 0.0
        else:
            tab["e_vib_up"][i] = tab["e_vib_lo"][i] = 0.0
    tab["up_set"], tab["lo_set"] = up, lo
    tab["n_sets"] = max(len(levels), 1)
    tab["level_energies"] = np.array(energy) if levels else None
    return tab


def calc_shapes_lines(wn_arr, lines, Temp, Pres, isomolec, n_threads=n_threads):
    """Shapes and G coefficients of every line at (Temp, Pres) (:1378-1415): returns the lines
    (non-LTE: only those linked to two known levels) with `.shape` (SpectralObject on the line's
    13010-point window centred on the nearest grid point) and `.G_coeffs`.  One batched GPU call
    replaces the fork/Queue fan-out."""
    if not isomolec.is_in_LTE:
        lines = [lin for lin in lines if lin.LinkToMolec(isomolec)]
    lines = [lin for lin in lines if lin.Mol == isomolec.mol and lin.Iso == isomolec.iso]
    if len(lines) == 0:
        return []
    grid = wn_arr.grid if hasattr(wn_arr, 'grid') else np.asarray(wn_arr)
    tab = line_table(lines, isomolec if not isomolec.is_in_LTE else None)
    ls = engine.LineSet(tab, grid, isomolec.MM, tab["n_sets"])
    shapes, g = ls.line_shapes(Pres, Temp)
    shapes, g = shapes.cpu().numpy(), g.cpu().numpy()
    order, centres = ls.order(), ls.centres()
    lin_grid = engine.line_window_offsets(grid)
    units = getattr(wn_arr, 'units', 'cm_1')
    for pos, src in enumerate(order):
        lin = lines[src]
        lin.shape = SpectralObject(shapes[pos], SpectralGrid(lin_grid + grid[centres[src]], units=units))
        lin.G_coeffs = dict(zip(CTYPES, g[pos]))
    ls.close()
    return [lines[i] for i in sorted(order)]


def PrepareCalcShapes(wn_arr, linee_mol, Temp, Pres, isomolec):
    """Core of the reference's worker (:1440-1462): shape (on the 13010-point window centred on the
    grid point nearest to the line) and G coefficients attached to EVERY line of linee_mol - no
    level filter, lines that cannot be linked get E_vib = 0 (:318-324).  One batched GPU call."""
    lines = list(linee_mol)
    if len(lines) == 0:
        return lines
    grid = wn_arr.grid if hasattr(wn_arr, 'grid') else np.asarray(wn_arr)
    tab = line_table(lines, None)                 # one set: every line is evaluated
    if len(isomolec.levels) > 0:
        for i, lin in enumerate(lines):
            ok = lin.LinkToMolec(isomolec)
            tab["e_vib_up"][i] = lin.E_vib_up if ok else 0.0
            tab["e_vib_lo"][i] = lin.E_vib_lo if ok else 0.0
    ls = engine.LineSet(tab, grid, isomolec.MM, 1)
    shapes, g = ls.line_shapes(Pres, Temp)
    shapes, g = shapes.cpu().numpy(), g.cpu().numpy()
    order, centres = ls.order(), ls.centres()
    lin_grid = engine.line_window_offsets(grid)
    for pos, src in enumerate(order):
        lin = lines[src]
        lin.shape = SpectralObject(shapes[pos], SpectralGrid(lin_grid + grid[centres[src]], units='cm_1'))
        lin.G_coeffs = dict(zip(CTYPES, g[pos]))
    ls.close()
    return lines


def do_for_th_calc(wn_arr, linee_tot, Temp, Pres, isomolec, i, coda, n_threads=n_threads):
    """The reference's per-process worker (:1418-1437): slice i of n_threads of the line list
    through PrepareCalcShapes, result on the queue `coda` (anything with .put)."""
    step_nlin = len(linee_tot) // n_threads
    linee = linee_tot[step_nlin * i:] if i == n_threads - 1 else \
        linee_tot[step_nlin * i:step_nlin * (i + 1)]
    coda.put(PrepareCalcShapes(wn_arr, linee, Temp, Pres, isomolec))


def sum_strength_lowres(lines, wn_range, mol, iso, Temp=None, ratios=None, res=5.0, plot=True):
    """Band-by-band sum of the line strengths in bins of `res` cm-1 (:1490-1529).  The reference
    only plots; here the sums are returned as (bin centres, {(upper, lower): array}) and plotted
    when matplotlib is importable and plot is True.  With Temp the strengths are the absorption
    strengths of CalcStrength_from_Einstein(Temp)."""
    arr_low = np.arange(wn_range[0], wn_range[1] + res / 2., res)
    mine = [lin for lin in lines if lin.Mol == mol and lin.Iso == iso]
    ups = [lin.minimal_level_string_up() for lin in mine]
    los = [lin.minimal_level_string_lo() for lin in mine]
    out = dict()
    for lev1 in np.unique(ups):
        for lev2 in np.unique(los):
            band = [lin for lin, u, l in zip(mine, ups, los) if u == lev1 and l == lev2]
            if len(band) == 0:
                continue
            freqs = np.array([lin.Freq for lin in band])
            if Temp is None:
                strengths = np.array([lin.Strength for lin in band])
            else:
                strengths = np.array([lin.CalcStrength_from_Einstein(Temp)[0] for lin in band])
            ratio = 1.0 if ratios is None else ratios[lev1]
            out[(str(lev1), str(lev2))] = np.array(
                [ratio * np.sum(strengths[(freqs > fr - res / 2.0) & (freqs < fr + res / 2.0)])
                 for fr in arr_low])
    if plot:
        try:
            import matplotlib.pyplot as pl
            for (lev1, lev2), spet in out.items():
                pl.plot(arr_low, spet, label=lev1 + ' -> ' + lev2)
            pl.legend()
            pl.grid()
        except ImportError:
            pass
    return arr_low, out


# ---------------------------------------------------------------------------------------------
# spectral containers
# ---------------------------------------------------------------------------------------------
class SpectralGrid(object):
    def __init__(self, spectral_grid, units='nm'):
        self.grid = np.array(spectral_grid, dtype=float)
        if len(self.grid) > imxsig_long:
            raise ValueError('grid longer than imxsig_long = %d' % imxsig_long)
        self.units = units

    def step(self):
        return self.grid[1] - self.grid[0]

    def len_wn(self):
        return len(self.grid)

    def wn_range(self):
        return [self.grid.min(), self.grid.max()]

    def min_wn(self):
        return self.grid.min()

    def max_wn(self):
        return self.grid.max()

    def half_precision(self):
        self.grid = self.grid.astype(np.float16)

    def double_precision(self):
        self.grid = self.grid.astype(float)

    # unit conversions of the axis (:396-432): nm <-> mum <-> cm_1 <-> hz; the grid is kept
    # ascending, so conversions between wavelength-like and wavenumber-like units reverse it
    def convertto_nm(self):
        if self.units == 'mum':
            self.grid = self.grid * 1.e3
        elif self.units == 'cm_1':
            self.grid = (1.e7 / self.grid)[::-1].copy()
        elif self.units == 'hz':
            self.grid = (const.c / (1.e-9 * self.grid))[::-1].copy()
        elif self.units != 'nm':
            raise ValueError('Cannot recognize units {}'.format(self.units))
        self.units = 'nm'
        return self.grid

    def convertto_cm_1(self):
        self.convertto_nm()
        self.grid = (1.e7 / self.grid)[::-1].copy()
        self.units = 'cm_1'
        return self.grid

    def convertto_mum(self):
        self.grid = self.convertto_nm() * 1.e-3
        self.units = 'mum'
        return self.grid

    def convertto_hz(self):
        self.convertto_nm()
        self.grid = (const.c / (1.e-9 * self.grid))[::-1].copy()
        self.units = 'hz'
        return self.grid


def convertto_nm(w, units):
    """Scalar / array axis value to nm (:2066-2078)."""
    if units == 'nm':
        return w
    if units == 'mum':
        return w * 1.e3
    if units == 'cm_1':
        return 1.e7 / w
    if units == 'hz':
        return const.c / (1.e-9 * w)
    raise ValueError('Cannot recognize units {}'.format(units))


def convertto_cm_1(w, units):
    return 1.e7 / convertto_nm(w, units)


def convertto_mum(w, units):
    return convertto_nm(w, units) * 1.e-3


def convertto_hz(w, units):
    return const.c / (1.e-9 * convertto_nm(w, units))


class SpectralObject(object):
    """A spectrum on a SpectralGrid (:435-1159; unit conversions and plotting left out)."""

    def __init__(self, spectrum, spectral_grid, direction=None, units='', link_grid=False):
        self.spectrum = np.array(spectrum, dtype=float)
        self.spectral_grid = spectral_grid if link_grid else copy.deepcopy(spectral_grid)
        self.direction = direction
        self.units = units

    def _other(self, obj2):
        return obj2.spectrum if isinstance(obj2, SpectralObject) else obj2

    def __getitem__(self, key):
        """self[w1, w2]: the part of the spectrum between w1 and w2, half a step of margin on both
        sides (:451-460); None when no grid point falls inside.  As in the reference the new
        SpectralGrid takes `self.units` (the units of the ordinate) as its units."""
        g = self.spectral_grid.grid
        half = self.spectral_grid.step() / 2.0
        cond = (g > key[0] - half) & (g < key[1] + half)
        if not np.any(cond):
            return None
        out = copy.deepcopy(self)
        out.spectral_grid = SpectralGrid(g[cond], units=self.units)
        out.spectrum = self.spectrum[cond]
        return out

    def __add__(self, obj2):
        """Spectra of the same length add point by point; a shorter spectrum on the same step is
        added where the grids overlap (add_to_spectrum, :462-472)."""
        out = copy.deepcopy(self)
        if isinstance(obj2, SpectralObject) and len(obj2.spectrum) != len(self.spectrum):
            out.add_to_spectrum(obj2)
        else:
            out.spectrum = self.spectrum + self._other(obj2)
        return out

    def __sub__(self, obj2):
        out = copy.deepcopy(self)
        if isinstance(obj2, SpectralObject) and len(obj2.spectrum) != len(self.spectrum):
            out.add_to_spectrum(obj2, Strength=-1.0)
        else:
            out.spectrum = self.spectrum - self._other(obj2)
        return out

    def __mul__(self, obj2):
        out = copy.deepcopy(self)
        out.spectrum = self.spectrum * self._other(obj2)
        return out

    def __truediv__(self, obj2):
        out = copy.deepcopy(self)
        out.spectrum = self.spectrum / self._other(obj2)
        return out

    __div__ = __truediv__

    def max(self):
        return np.max(self.spectrum)

    def min(self):
        return np.min(self.spectrum)

    def n_points(self):
        return len(self.spectrum)

    def add_mask(self, mask):
        self.mask = mask

    def multiply(self, factor, save=True):
        if save:
            self.spectrum = self.spectrum * factor
            return self
        out = copy.deepcopy(self)
        out.spectrum = self.spectrum * factor
        return out

    def exp_elementwise(self, exp_factor, save=False):
        """exp(spectrum * exp_factor) (:693-701)."""
        new = np.exp(self.spectrum * exp_factor)
        if save:
            self.spectrum = new
            return None
        out = copy.deepcopy(self)
        out.spectrum = new
        return out

    def multiply_elementwise(self, spectrum2, save=True):
        """(:975-990)"""
        if len(self.spectrum) != len(spectrum2.spectrum):
            raise ValueError('The two spectra have different lengths!')
        new = spectrum2.spectrum * self.spectrum
        if save:
            self.spectrum = new
            return None
        out = copy.deepcopy(self)
        out.spectrum = new
        return out

    def divide_elementwise(self, spectrum2, save=True):
        """(:993-1008)"""
        if len(self.spectrum) != len(spectrum2.spectrum):
            raise ValueError('The two spectra have different lengths!')
        new = self.spectrum / spectrum2.spectrum
        if save:
            self.spectrum = new
            return None
        out = copy.deepcopy(self)
        out.spectrum = new
        return out

    def sum_scalar(self, scalar):
        self.spectrum = self.spectrum + scalar
        return self.spectrum

    def interp_to_grid(self, nugrid):
        """Linear interpolation to another SpectralGrid, zero outside this one (:501-507)."""
        out = copy.deepcopy(self)
        out.spectrum = np.interp(nugrid.grid, self.spectral_grid.grid, self.spectrum, left=0.0,
                                 right=0.0)
        out.spectral_grid = copy.deepcopy(nugrid)
        return out

    def interp_to_regular_grid(self):
        """In place: the same number of points, evenly spaced between the ends (:920-927)."""
        g = self.spectral_grid.grid
        new = SpectralGrid(np.linspace(np.min(g), np.max(g), len(g)), units=self.spectral_grid.units)
        self.spectrum = np.interp(new.grid, g, self.spectrum)
        self.spectral_grid = new

    def _significant(self, thres, stride=1, consider_derivatives=True):
        """Points (every `stride`-th) where the spectrum, or its first / second np.gradient, exceeds
        thres x its own maximum - the selection rule of degrade_grid / degrade_grid2."""
        oks = self.spectrum[::stride] > thres * self.max()
        if consider_derivatives:
            d1 = np.gradient(self.spectrum)
            d2 = np.abs(np.gradient(d1))
            d1 = np.abs(d1)
            oks = oks | (d1[::stride] > thres * np.max(d1)) | (d2[::stride] > thres * np.max(d2))
        return oks

    def degrade_grid(self, thress=(1.e-3, 1.e-4, 1.e-5), factors=(5, 20, 100),
                     consider_derivatives=True):
        """Irregular grid that keeps full resolution where the spectrum (or a derivative) is above
        thress[0] x max, every factors[0]-th point above thress[1] x max, and every
        factors[-1]-th point everywhere; the spectrum is interpolated onto it (:509-547).  As in
        the reference only the first two thresholds and the strides (1, factors[0]) enter, plus
        the coarsest stride factors[-1]."""
        strides = [1] + list(factors)
        g = self.spectral_grid.grid
        keep = [g[::strides[-1]]]
        for thres, stride in list(zip(thress, strides[:-1]))[:2]:
            keep.append(g[::stride][self._significant(thres, stride, consider_derivatives)])
        new = SpectralGrid(np.unique(np.concatenate(keep)), units=self.spectral_grid.units)
        return self.interp_to_grid(new)

    def degrade_grid2(self, thres=1.e-4, num_aside=(10, 5, 5), res_low=(1, 2, 5),
                      consider_derivatives=True):
        """Keeps the points above thres x max (spectrum or derivatives) and, around them, num_aside[k]
        rings of neighbours at distance (i+1)*res_low[k], of which every res_low[k]-th candidate (in
        index order) is taken (:549-600)."""
        n = len(self.spectrum)
        hi = np.flatnonzero(self._significant(thres, 1, consider_derivatives))
        if len(hi) == 0:
            raise IndexError('no point above the threshold')
        part = hi
        for num, res in zip(num_aside, res_low):
            for i in range(num):
                up = hi + (i + 1) * res
                part = np.unique(np.append(part, up[up < n][::res]))
                dn = hi - (i + 1) * res
                part = np.unique(np.append(part, dn[dn >= 0][::res]))
            hi = part
        new = SpectralGrid(self.spectral_grid.grid[np.unique(hi)], units=self.spectral_grid.units)
        return self.interp_to_grid(new)

    def compress_spectrum(self, threshold=1.e-25):
        """Spectrum as a scipy.sparse.csr_matrix with the values below threshold zeroed (:748-755)."""
        import scipy.sparse
        coso = self.spectrum
        coso[self.spectrum < threshold] = 0.0
        self.spectrum = scipy.sparse.csr_matrix(coso)

    def erase_grid(self):
        self.spectral_grid = None

    def restore_grid(self, spectral_grid, link_grid=False):
        self.spectral_grid = spectral_grid if link_grid else copy.deepcopy(spectral_grid)

    def half_precision(self):
        """float32 spectrum (the reference's "half" precision, :722-737); a grid that is still
        attached goes to float16 like SpectralGrid.half_precision does."""
        self.spectrum = self.spectrum.astype(np.float32)
        if self.spectral_grid is not None:
            self.spectral_grid.half_precision()

    def double_precision(self):
        self.spectrum = self.spectrum.astype(np.float64)
        if self.spectral_grid is not None:
            self.spectral_grid.double_precision()

    # spectral density per unit of the axis: converting the axis multiplies by |d old / d new|
    # and, between wavelength-like and wavenumber-like axes, reverses the arrays (:753-797)
    def convertto_nm(self):
        u = self.spectral_grid.units
        g = self.spectral_grid.grid
        if u == 'mum':
            self.spectrum = self.spectrum * 1.e-3
        elif u == 'cm_1':
            self.spectrum = (self.spectrum * g ** 2 * 1.e-7)[::-1].copy()
        elif u == 'hz':
            self.spectrum = (self.spectrum * g ** 2 * 1.e-9 / const.c)[::-1].copy()
        self.spectral_grid.convertto_nm()
        return self.spectral_grid.grid, self.spectrum

    def convertto_mum(self):
        self.convertto_nm()
        self.spectrum = self.spectrum * 1.e3
        self.spectral_grid.convertto_mum()
        return self.spectral_grid.grid, self.spectrum

    def convertto_cm_1(self):
        self.convertto_nm()
        self.spectrum = (self.spectrum * self.spectral_grid.grid ** 2 * 1.e-7)[::-1].copy()
        self.spectral_grid.convertto_cm_1()
        return self.spectral_grid.grid, self.spectrum

    def convertto_hz(self):
        self.convertto_nm()
        self.spectrum = (self.spectrum * self.spectral_grid.grid ** 2 * 1.e-9 / const.c)[::-1].copy()
        self.spectral_grid.convertto_hz()
        return self.spectral_grid.grid, self.spectrum

    def convert_grid_to(self, units):
        """(:738-751)"""
        if units == self.spectral_grid.units:
            return self.spectral_grid.grid, self.spectrum
        conv = dict(nm=self.convertto_nm, mum=self.convertto_mum, cm_1=self.convertto_cm_1,
                    hz=self.convertto_hz)
        if units not in conv:
            raise ValueError('Cannot recognize units {}'.format(units))
        return conv[units]()

    def integrate(self, w1=None, w2=None, offset=None):
        g = self.spectral_grid.grid
        cond = ~np.isnan(self.spectrum)
        if w1 is not None:
            cond &= g >= w1
        if w2 is not None:
            cond &= g <= w2
        spect = self.spectrum if offset is None else self.spectrum - offset
        return np.trapezoid(spect[cond], x=g[cond])

    def convolve_to_grid_from_irregular(self, new_spectral_grid, spectral_widths=None,
                                        conv_type='gaussian', n_sigma=5.):
        """Gaussian instrument convolution to another grid (:883-918), on the GPU."""
        if conv_type != 'gaussian':
            raise ValueError('only the gaussian convolution exists in the reference')
        new_len = len(new_spectral_grid.grid)
        if spectral_widths is None:
            spectral_widths = [new_spectral_grid.step()] * new_len
        elif isinstance(spectral_widths, (int, float)):
            spectral_widths = [spectral_widths] * new_len
        if len(spectral_widths) != new_len:
            raise ValueError('{} spectral widths for {} grid points'.format(len(spectral_widths), new_len))
        out = copy.deepcopy(self)
        out.spectral_grid = copy.deepcopy(new_spectral_grid)
        out.spectrum = engine.convolve_lowres_host(self.spectral_grid.grid, self.spectrum,
                                                   new_spectral_grid.grid, spectral_widths,
                                                   n_sigma)[0]
        if hasattr(out, 'intensity'):
            out.intensity = out.spectrum
        return out

    def convolve_to_grid(self, new_spectral_grid, spectral_widths=None, conv_type='gaussian',
                         n_sigma=5.):
        """Gaussian convolution of a spectrum on a REGULAR grid to another grid, the reference's
        older host routine (:832-881; hires_to_lowres_old and tolowres use it): one window of
        2*int(n_sigma*max width/step) points of the old step, centred on the old grid point
        nearest to every new point (closest_grid_ext extrapolates the grid outside the range),
        the part of the spectrum inside it zero-extended to the window and integrated against the
        Gaussian with the trapezoid rule.  Host NumPy; the batched device convolution is
        convolve_to_grid_from_irregular."""
        if conv_type != 'gaussian':
            raise ValueError('only the gaussian convolution exists in the reference')
        new_len = len(new_spectral_grid.grid)
        weed = new_spectral_grid.step() if spectral_widths is None else max(spectral_widths)
        step_old = self.spectral_grid.step()
        n_points = int(n_sigma * weed / step_old)
        lin_grid = np.arange(-n_points * step_old, n_points * step_old, step_old, dtype=float)
        if spectral_widths is None:
            window = gaussian(lin_grid, 0.0, new_spectral_grid.step())
            spectral_widths = [None] * new_len
        spectrum = np.zeros(new_len, dtype=float)
        for num, (freq, wid) in enumerate(zip(new_spectral_grid.grid, spectral_widths)):
            if wid is not None:
                window = gaussian(lin_grid, 0.0, wid)
            _, fr_ok = closest_grid_ext(self.spectral_grid, freq)
            win_grid = SpectralGrid(lin_grid + fr_ok, units=self.spectral_grid.units)
            old = self[win_grid.grid[0], win_grid.grid[-1]]
            if old is None:
                continue
            if len(old.spectrum) < len(window):
                zero = SpectralObject(np.zeros(len(win_grid.grid)), win_grid)
                zero.add_to_spectrum(old)
                old = zero
            spectrum[num] = conv_single(old, window, new_spectral_grid.step())
        out = copy.deepcopy(self)
        out.spectral_grid = copy.deepcopy(new_spectral_grid)
        out.spectrum = spectrum
        if hasattr(out, 'intensity'):
            out.intensity = out.spectrum
        return out

    def add_to_spectrum(self, spectrum2, Strength=None, sumcheck=10.):
        """Adds a spectrum whose grid lies (partly) inside this one and has the SAME step, where
        the two overlap (tolerance step/10 at the ends, :929-953)."""
        g, g2 = self.spectral_grid.grid, spectrum2.spectral_grid.grid
        spino = self.spectral_grid.step() / 10.
        ok = (g > g2[0] - spino) & (g < g2[-1] + spino)
        ok2 = (g2 > g[0] - spino) & (g2 < g[-1] + spino)
        if Strength is not None:
            self.spectrum[ok] += Strength * spectrum2.spectrum[ok2]
        else:
            self.spectrum[ok] += spectrum2.spectrum[ok2]

    def add_to_spectrum_slow(self, spectrum2, Strength=None):
        """The argmin variant (:955-972): as in the reference the last overlapping point is left
        out (half-open slices)."""
        g, g2 = self.spectral_grid.grid, spectrum2.spectral_grid.grid
        ini_1, fin_1 = np.argmin(np.abs(g - g2[0])), np.argmin(np.abs(g - g2[-1]))
        ini_2, fin_2 = np.argmin(np.abs(g2 - g[0])), np.argmin(np.abs(g2 - g[-1]))
        add = spectrum2.spectrum[ini_2:fin_2]
        self.spectrum[ini_1:fin_1] += add if Strength is None else Strength * add

    def _fortran_rows(self, lines, fix_length=imxsig, Strengths=None):
        """(matrix [n_lines][fix_length] column-major, init, fin) of prepare_fortran_sum
        (:1100-1147): the points of every line inside the spectrum range (tolerance step/10),
        the rest of the row zero padding on the right, or on the left with a shifted start when
        the window sticks out at the high end; init / fin 1-based inclusive."""
        n_lines = len(lines)
        g = self.spectral_grid.grid
        spino = self.spectral_grid.step() / 10.
        matrix = np.zeros((n_lines, fix_length), order='F')
        init = np.zeros(n_lines, dtype=np.int32)
        fin = np.zeros(n_lines, dtype=np.int32)
        for i, line in enumerate(lines):
            lg = line.spectral_grid.grid
            y = line.spectrum if Strengths is None else line.spectrum * Strengths[i]
            ok = np.flatnonzero((g > lg[0] - spino) & (g < lg[-1] + spino))
            ok2 = (lg > g[0] - spino) & (lg < g[-1] + spino)
            ini, end = ok[0] + 1, ok[-1] + 1          # 1-based, inclusive
            n_zeri = fix_length - (end - ini + 1)
            if n_zeri > 0:
                if end + n_zeri < self.n_points():
                    matrix[i, :fix_length - n_zeri] = y[ok2]
                    end += n_zeri
                else:
                    matrix[i, n_zeri:] = y[ok2]
                    ini -= n_zeri
            else:
                matrix[i, :] = y
            init[i], fin[i] = ini, end
        return matrix, init, fin

    def prepare_fortran_sum(self, lines, i, coda, fix_length=imxsig):
        """The reference's worker of add_lines_to_spectrum (:1100-1147): puts [matrix, init, fin]
        of this slice of lines on the queue `coda` (anything with .put)."""
        matrix, init, fin = self._fortran_rows(lines, fix_length)
        coda.put([matrix, init.astype(int), fin.astype(int)])

    def add_lines_to_spectrum(self, lines, Strengths=None, fix_length=imxsig, n_threads=n_threads):
        """Adds line shapes (SpectralObjects on their own 13010-point windows) to this spectrum
        through the sum_all_lines drop-in (:1016-1097); window clipping as prepare_fortran_sum."""
        n_lines = len(lines)
        if n_lines == 0:
            return self.spectrum
        if n_lines > imxlines:
            raise ValueError('{} are too many lines (imxlines = {})'.format(n_lines, imxlines))
        if self.n_points() > imxsig_long:
            raise ValueError('The input spectrum is too long (imxsig_long)')
        matrix, init, fin = self._fortran_rows(lines, fix_length, Strengths)
        self.spectrum = lineshape.sum_all_lines(self.spectrum, matrix, init, fin, n_lines,
                                                self.n_points())
        return self.spectrum


def conv_single(spect, window, step):
    """Trapezoid integral of spectrum x window on the spectrum's grid (:1162-1164)."""
    return np.trapezoid(spect.spectrum * window, x=spect.spectral_grid.grid)


class SpectralIntensity(SpectralObject):
    def __init__(self, intensity, spectral_grid, direction=None, units='ergscm2'):
        SpectralObject.__init__(self, intensity, spectral_grid, direction=direction, units=units)
        self.intensity = self.spectrum

    INTENSITY_TO_WM2 = dict(Wm2=1.0, ergscm2=1.e-3, nWcm2=1.e-5)     # (:1208-1237)

    def convertto(self, new_units):
        """Intensity units 'Wm2', 'ergscm2', 'nWcm2' (:1200-1237)."""
        if new_units not in self.INTENSITY_TO_WM2:
            raise ValueError('No method for units ' + str(new_units))
        if new_units != self.units:
            self.spectrum = self.spectrum * (self.INTENSITY_TO_WM2[self.units] /
                                             self.INTENSITY_TO_WM2[new_units])
            self.intensity = self.spectrum
            self.units = new_units
        return self.spectrum

    def convertto_Wm2(self):
        return self.convertto('Wm2')

    def convertto_ergscm2(self):
        return self.convertto('ergscm2')

    def convertto_nWcm2(self):
        return self.convertto('nWcm2')

    def add_noise(self, noise):
        self.noise = copy.deepcopy(noise)

    def add_bands(self, bands):
        self.bands = copy.deepcopy(bands)

    def hires_to_lowres_old(self, lowres_obs, spectral_widths=None):
        """The older chain (:1193-1198): axis conversion in place, regular grid, host
        convolve_to_grid, intensity units of the observation."""
        self.convert_grid_to(lowres_obs.spectral_grid.units)
        self.interp_to_regular_grid()
        low = self.convolve_to_grid(lowres_obs.spectral_grid, spectral_widths=spectral_widths)
        low.convertto(lowres_obs.units)
        return low

    def hires_to_lowres(self, lowres_obs, spectral_widths=None, keep_original_hires=True):
        """Low-res spectrum on the observation's grid, in the observation's axis and intensity
        units (:1180-1191): the spectrum is converted to the axis of the observation
        (convert_grid_to), convolved there, and its intensity units follow lowres_obs.units.  A
        cm-1 spectrum seen by an nm observation (the VIMS case) is converted and convolved in one
        device pass."""
        gigi = copy.deepcopy(self) if keep_original_hires else self
        want = lowres_obs.spectral_grid.units
        if gigi.spectral_grid.units == 'cm_1' and want == 'nm':
            if spectral_widths is None:
                spectral_widths = [lowres_obs.spectral_grid.step()] * len(lowres_obs.spectral_grid.grid)
            elif isinstance(spectral_widths, (int, float)):
                spectral_widths = [spectral_widths] * len(lowres_obs.spectral_grid.grid)
            low = copy.deepcopy(gigi)
            low.spectral_grid = copy.deepcopy(lowres_obs.spectral_grid)
            low.spectrum = engine.convolve_lowres_host(gigi.spectral_grid.grid, gigi.spectrum,
                                                       lowres_obs.spectral_grid.grid,
                                                       spectral_widths, 5., units='nm')[0]
            low.intensity = low.spectrum
        else:
            gigi.convert_grid_to(want)
            low = gigi.convolve_to_grid_from_irregular(lowres_obs.spectral_grid,
                                                       spectral_widths=spectral_widths)
        low.convertto(getattr(lowres_obs, 'units', low.units))
        return low


class SpectralGcoeff(SpectralObject):
    """G-coefficient spectrum of one (level, ctype) at one (P, T) (:1247-1375)."""

    def __init__(self, ctype, spectral_grid, mol, iso, MM, minimal_level_string,
                 unidentified_lines=False, spectrum=None, Pres=None, Temp=None):
        if spectrum is None:
            spectrum = np.zeros(len(spectral_grid.grid))
        SpectralObject.__init__(self, spectrum, spectral_grid, link_grid=True)
        self.ctype, self.mol, self.iso, self.MM = ctype, mol, iso, MM
        self.lev_string = minimal_level_string
        self.unidentified_lines = unidentified_lines
        self.pres, self.temp = Pres, Temp

    def BuildCoeff(self, lines, Temp, Pres, n_threads=n_threads, preCalc_shapes=False,
                   debug=False, isomolec=None):
        """Sum of G*shape over the lines that have this level as upper (emission ctypes) or lower
        (absorption) level, or over all lines of (mol, iso) when unidentified (:1277-1337)."""
        self.temp, self.pres = Temp, Pres
        if len(lines) == 0:
            return self.spectrum
        if not preCalc_shapes:
            if isomolec is None:
                raise ValueError('BuildCoeff without precalculated shapes needs the isomolec')
            lines = calc_shapes_lines(self.spectral_grid, lines, Temp, Pres, isomolec)
        elif not hasattr(lines[0], 'shape') or not hasattr(lines[0], 'G_coeffs'):
            raise ValueError('preCalc_shapes is True but the lines carry no shapes: run '
                             'calc_shapes_lines first')
        if self.ctype not in CTYPES:
            raise ValueError('ctype {} not recognized'.format(self.ctype))
        mine = [lin for lin in lines if lin.Mol == self.mol and lin.Iso == self.iso]
        if not self.unidentified_lines:
            if self.ctype == 'absorption':
                mine = [lin for lin in mine if self.lev_string == lin.minimal_level_string_lo()]
            else:
                mine = [lin for lin in mine if self.lev_string == lin.minimal_level_string_up()]
        if mine:
            self.add_lines_to_spectrum([lin.shape for lin in mine],
                                       Strengths=[lin.G_coeffs[self.ctype] for lin in mine])
        return self.spectrum

    def calc_shapes(self, lines, Temp, Pres, isomolec):
        """Shapes and G coefficients of `lines` on this coefficient's grid (:1340-1347)."""
        return calc_shapes_lines(self.spectral_grid, lines, Temp, Pres, isomolec)

    def interpolate(self, coeff2, Pres=None, Temp=None):
        """Linear blend with another coefficient that differs in P or in T (:1349-1375)."""
        if coeff2 is None:
            return None
        same_t, same_p = sbm.isclose(self.temp, coeff2.temp), sbm.isclose(self.pres, coeff2.pres)
        if not same_t and not same_p:
            raise ValueError('The two coeffs have both different temperatures and pressures')
        if Pres is not None:
            if not same_t:
                raise ValueError('The two coeffs have different temperatures')
            w1, w2 = sbm.weight(Pres, self.pres, coeff2.pres, itype='lin')
            p_new, t_new = Pres, self.temp
        elif Temp is not None:
            if not same_p:
                raise ValueError('The two coeffs have different pressures')
            w1, w2 = sbm.weight(Temp, self.temp, coeff2.temp, itype='lin')
            p_new, t_new = self.pres, Temp
        else:
            raise ValueError('give Pres or Temp')
        return SpectralGcoeff(self.ctype, self.spectral_grid, self.mol, self.iso, self.MM,
                              self.lev_string, spectrum=w1 * self.spectrum + w2 * coeff2.spectrum,
                              Pres=p_new, Temp=t_new)
