"""Stand-in for the reference's missing `spect_base_module`, used ONLY to execute the reference's
own spect_classes.py / spect_main_module.py when the golden fixtures are generated (ref_exec.py).

The module is not part of the reference tree, so these few names carry the semantics
DESIGN.md section 6 publishes; everything else that runs during fixture generation is the
reference's own Python.  Loaded under the module name `spect_base_module`, so objects the
reference pickles (Level inside LutSet) resolve to the product's classes of the same name.
"""
import math as mt
import os

import numpy as np

REFERENCE = os.environ.get('SPECTROBOT_REFERENCE', '/root/reference')
_MOLPARAM = None


def isclose(a, b, rtol=1.e-9, atol=0.0):
    return np.isclose(a, b, rtol=rtol, atol=atol)


def weight(v, v1, v2, itype='lin'):
    if itype != 'lin':
        raise ValueError(itype)
    return (v2 - v) / (v2 - v1), (v - v1) / (v2 - v1)


def rad(deg):
    return deg * mt.pi / 180.0


def _molparam():
    """molparam.txt of the reference: 'Molecule # Iso Abundance Q(296K) gj Molar Mass(g)'."""
    global _MOLPARAM
    if _MOLPARAM is None:
        tab, mol, name = dict(), None, None
        for ln in open(os.path.join(REFERENCE, 'molparam.txt')):
            tok = ln.split()
            if len(tok) == 2 and tok[1].startswith('(') and tok[1].endswith(')'):
                name, mol = tok[0], int(tok[1][1:-1])
                tab[mol] = dict(name=name, isos=[])
            elif len(tok) == 5 and mol is not None:
                try:
                    tab[mol]['isos'].append(dict(code=tok[0], iso_ratio=float(tok[1]),
                                                 Q296=float(tok[2]), gj=int(tok[3]),
                                                 iso_MM=float(tok[4])))
                except ValueError:
                    pass
        _MOLPARAM = tab
    return _MOLPARAM


def find_molec_metadata(mol, iso):
    m = _molparam()[int(mol)]
    e = m['isos'][int(iso) - 1]
    return {'mol_name': m['name'], 'iso_name': e['code'], 'iso_MM': e['iso_MM'],
            'iso_ratio': e['iso_ratio'], 'Q296': e['Q296'], 'gj': e['gj']}


def extract_quanta_HITRAN(mol, iso, lev_string):
    if isinstance(lev_string, bytes):
        lev_string = lev_string.decode()
    quanta, sym = [], ''
    for t in str(lev_string).split():
        try:
            quanta.append(int(t))
        except ValueError:
            sym = t
            break
    return ' '.join(str(q) for q in quanta), quanta, sym


def vibtemp_to_ratio(energy, T_vib, T):
    import scipy.constants as const
    c2 = const.h * const.c * 100.0 / const.k
    return np.exp(-c2 * energy * (1.0 / T_vib - 1.0 / T))


def trova_spip(ifile, hasha='#', read_past=False):
    while True:
        line = ifile.readline()
        if not line:
            return False
        if line.lstrip().startswith(hasha):
            if read_past:
                ifile.readline()
            return True


find_spip = trova_spip


class Level(object):
    def __init__(self, levstring, energy, degen=None, simmetry=None):
        self.lev_string = levstring
        self.energy = float(energy)
        self.degen = degen
        self.simmetry = simmetry or []
        self.vibtemp = None
        self.local_vibtemp = []

    def minimal_level_string(self):
        return extract_quanta_HITRAN(None, None, self.lev_string)[0]

    def get_quanta(self):
        _, q, s = extract_quanta_HITRAN(None, None, self.lev_string)
        return q, s

    def equiv(self, string):
        return extract_quanta_HITRAN(None, None, string)[0] == self.minimal_level_string()

    def add_local_vibtemp(self, temp):
        self.local_vibtemp.append(temp)


class IsoMolec(object):
    def __init__(self, mol, iso, MM=None, ratio=None, LTE=True):
        self.mol, self.iso = int(mol), int(iso)
        md = find_molec_metadata(mol, iso)
        self.MM = md['iso_MM'] if MM is None else MM
        self.ratio = md['iso_ratio'] if ratio is None else ratio
        self.mol_name = md['mol_name']
        self.is_in_LTE = LTE
        self.levels = []
        self.n_lev = 0

    def add_levels(self, lev_strings, energies, vibtemps=None, degeneracies=None, simmetries=None):
        for ls_, en in zip(lev_strings, energies):
            name = 'lev_{:02d}'.format(self.n_lev)
            setattr(self, name, Level(ls_, en))
            self.levels.append(name)
            self.n_lev += 1
        if lev_strings:
            self.is_in_LTE = False

    def has_level(self, lev_string):
        for lev in self.levels:
            if getattr(self, lev).equiv(lev_string):
                return True, lev
        return False, None

    def erase_level(self, lev):
        self.levels.remove(lev)
        delattr(self, lev)


class Molec(object):
    def __init__(self, mol, name, MM=None):
        self.mol, self.name, self.MM = int(mol), name, MM
        self.all_iso = []
        self.iso_N = 0

    def add_iso(self, num, MM=None, ratio=None, LTE=True):
        nam = 'iso_{:1d}'.format(num)
        setattr(self, nam, IsoMolec(self.mol, num, MM=MM, ratio=ratio, LTE=LTE))
        self.all_iso.append(nam)
        self.iso_N += 1
        return getattr(self, nam)


class AtmGrid(object):
    def __init__(self, names, coords):
        if isinstance(names, str):
            names, coords = [names], [coords]
        self.names = list(names)
        self.coords = dict((n, np.asarray(c, dtype=float)) for n, c in zip(names, coords))
        self.grid = [self.coords[n] for n in self.names]
        self.n_dim = len(self.names)


class AtmGridMask(object):
    def __init__(self, grid, mask, interp='lin'):
        self.grid = grid
        self.mask = np.asarray(mask, dtype=float)
        self.interp = {'mask': interp if isinstance(interp, str) else interp[-1]}


class AtmProfile(object):
    """Only what calc_PT_couples_atmosphere touches: .pres and .temp arrays."""

    def __init__(self, grid, values, profname, interp):
        self.grid = grid
        self.names = []
        self.values = dict()
        self.add_profile(values, profname, interp)

    def add_profile(self, values, profname, interp='lin'):
        self.names.append(profname)
        self.values[profname] = np.asarray(values, dtype=float)
        setattr(self, profname, self.values[profname])
