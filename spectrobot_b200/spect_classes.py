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
3).  Unit conversions, plotting and `degrade_grid*` are out of scope.
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

    def CalcStrength(self, T):
        return CalcStrength_at_T(self.Mol, self.Iso, self.Strength, self.Freq, self.E_lower, T)

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

    def __add__(self, obj2):
        out = copy.deepcopy(self)
        out.spectrum = self.spectrum + self._other(obj2)
        return out

    def __sub__(self, obj2):
        out = copy.deepcopy(self)
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

    def multiply(self, factor, save=True):
        if save:
            self.spectrum = self.spectrum * factor
            return self
        out = copy.deepcopy(self)
        out.spectrum = self.spectrum * factor
        return out

    def erase_grid(self):
        self.spectral_grid = None

    def restore_grid(self, spectral_grid, link_grid=False):
        self.spectral_grid = spectral_grid if link_grid else copy.deepcopy(spectral_grid)

    def half_precision(self):
        self.spectrum = self.spectrum.astype(np.float32)

    def double_precision(self):
        self.spectrum = self.spectrum.astype(np.float64)

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

    def add_lines_to_spectrum(self, lines, Strengths=None, fix_length=imxsig, n_threads=n_threads):
        """Adds line shapes (SpectralObjects on their own 13010-point windows) to this spectrum
        through the sum_all_lines drop-in (:1016-1097).  Window clipping follows
        prepare_fortran_sum (:1100-1147): points of the line inside the spectrum range (tolerance
        step/10) are kept, the rest of the 13010-wide row is zero padding on the right, or on the
        left with a shifted start when the window sticks out at the high end."""
        n_lines = len(lines)
        if n_lines == 0:
            return self.spectrum
        if n_lines > imxlines:
            raise ValueError('{} are too many lines (imxlines = {})'.format(n_lines, imxlines))
        if self.n_points() > imxsig_long:
            raise ValueError('The input spectrum is too long (imxsig_long)')
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
        self.spectrum = lineshape.sum_all_lines(self.spectrum, matrix, init, fin, n_lines,
                                                self.n_points())
        return self.spectrum


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
