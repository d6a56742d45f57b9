"""Minimal `spect_base_module` (sbm) for the hot path.

The reference imports `spect_base_module as sbm` everywhere, but the module is NOT part of the
reference tree, so nothing here can be checked against the original.  This file provides
the surface the hot path needs, inferred from the call sites, with the
semantics published in DESIGN.md section 6:

  helpers      isclose, weight, rad, find_molec_metadata, extract_quanta_HITRAN, vibtemp_to_ratio,
               hydro_P, read_inputs, check_free_space
  atmosphere   AtmGrid, AtmProfile (1-D in altitude or 2-D latitude-band x altitude), Titan
  molecules    Molec, IsoMolec, Level
  geometry     Coords, LineOfSight (ray / spherical-shell intersections, radtran steps with
               Curtis-Godson columns through the `curgods` drop-in, radtran_fast on the GPU),
               VIMSPixel
Geometry and step construction are host-side set-up; the per-point arithmetic of
the LOS integral runs in libspectrobot.so.
"""
import copy
import json
import math as mt
import os

import numpy as np

from . import curgods

_HERE = os.path.dirname(os.path.abspath(__file__))
_MOLPARAM = None

kb_hpa = 1.38065e-19      # spect_classes.py:34 (P in hPa, n in cm-3)
c_R = 8.31446             # spect_classes.py:35
c_G = 6.67408e-11         # spect_classes.py:36
T_ref = 296.0


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------
def isclose(a, b, rtol=1.e-9, atol=0.0):
    """Scalar/array closeness (call sites spect_classes.py:933,1359; spect_main_module.py:1425)."""
    return np.isclose(a, b, rtol=rtol, atol=atol)


def weight(v, v1, v2, itype='lin'):
    """(w1, w2) such that value(v) = w1*value(v1) + w2*value(v2) (spect_classes.py:1365)."""
    if itype == 'lin':
        return (v2 - v) / (v2 - v1), (v - v1) / (v2 - v1)
    if itype == 'exp':
        return weight(mt.log(v), mt.log(v1), mt.log(v2), 'lin')
    raise ValueError('unknown interpolation type ' + str(itype))


def rad(deg):
    return deg * mt.pi / 180.0


def _molparam():
    global _MOLPARAM
    if _MOLPARAM is None:
        _MOLPARAM = json.load(open(os.path.join(_HERE, "data", "molparam.json")))
    return _MOLPARAM


def find_molec_metadata(mol, iso):
    """HITRAN isotopologue metadata from molparam.txt: iso_MM, iso_ratio, Q296, gj, mol_name
    (call sites spect_classes.py:179,233,270,301)."""
    m = _molparam()[str(int(mol))]
    e = m["isos"][int(iso) - 1]
    return {'mol_name': m["name"], 'iso_name': str(e["code"]), 'iso_MM': e["iso_MM"],
            'iso_ratio': e["iso_ratio"], 'Q296': e["Q296"], 'gj': e["gj"]}


def extract_quanta_HITRAN(mol, iso, lev_string):
    """(minimal_string, quanta, symmetry): the vibrational quanta part of a HITRAN global-quanta
    string, without the symmetry label (spect_classes.py:346-351, 1251)."""
    toks = str(lev_string).split()
    quanta = []
    sym = ''
    for t in toks:
        try:
            quanta.append(int(t))
        except ValueError:
            sym = t
            break
    return ' '.join(str(q) for q in quanta), quanta, sym


def vibtemp_to_ratio(energy, T_vib, T):
    """Ratio of the non-LTE to the LTE population of a level (spect_classes.py:280-281)."""
    from .spect_classes import c2
    return np.exp(-c2 * energy * (1.0 / T_vib - 1.0 / T))


def hydro_P(z_km, temp, MM, P0=1467.0, R_km=2575.0, M_kg=1.3452e23):
    """Hydrostatic pressure profile (titanatm.py:32)."""
    z = np.asarray(z_km, dtype=float)
    g = c_G * M_kg / ((R_km + z) * 1e3) ** 2
    H = c_R * np.asarray(temp, dtype=float) / (MM * 1e-3 * g) / 1e3
    integ = np.concatenate([[0.0], np.cumsum(0.5 * (1 / H[1:] + 1 / H[:-1]) * np.diff(z))])
    return P0 * np.exp(-integ)


def read_inputs(nomefile, key_strings, n_lines=None, itype=None, defaults=None, verbose=False):
    """Parser of the `[key]\\nvalue` input files (inputs_spect_robot_SAMPLE.in;
    radtran_3D_ch4.py:45-49)."""
    txt = [ln.rstrip('\n') for ln in open(nomefile)]
    out = dict()
    for i, key in enumerate(key_strings):
        val = None
        for j, ln in enumerate(txt):
            if ln.strip() == '[' + key + ']' and j + 1 < len(txt):
                val = txt[j + 1].strip()
        if val is None and defaults is not None:
            val = defaults.get(key) if isinstance(defaults, dict) else defaults[i]
        if val is not None and itype is not None and itype[i] is not None and isinstance(val, str):
            cast = itype[i]
            val = (val.lower() in ('true', '1', 't')) if cast is bool else cast(val)
        out[key] = val
        if verbose:
            print(key, val)
    return out


def trova_spip(ifile, hasha='#', read_past=False):
    """Advances an open text file to the line after the next one starting with `hasha` (the
    "#"-delimited headers of the gbb-style input files; surface inferred from the call sites
    spect_classes.py:1558 and Appendix C of SURVEY.md).  Returns False at end of file."""
    while True:
        line = ifile.readline()
        if not line:
            return False
        if line.lstrip().startswith(hasha):
            if read_past:
                ifile.readline()
            return True


find_spip = trova_spip


def check_free_space(path):
    st = os.statvfs(path)
    return st.f_bavail * st.f_frsize / 1.e9


# ---------------------------------------------------------------------------------------------
# atmosphere
# ---------------------------------------------------------------------------------------------
class AtmGrid(object):
    def __init__(self, names, coords):
        if isinstance(names, str):
            names, coords = [names], [coords]
        self.names = list(names)
        self.coords = dict((n, np.asarray(c, dtype=float)) for n, c in zip(names, coords))
        self.grid = [self.coords[n] for n in self.names]
        self.n_dim = len(self.names)


class AtmProfile(object):
    """Profiles on an AtmGrid: 'alt', ('lat', 'alt'), ('sza', 'alt') or ('lat', 'sza', 'alt') - the
    last is the reference's 3-D vibrational-temperature profile, `prof.calc([lat, sza, alt])`
    (radtran_3D_ch4.py:249-250).  Latitude with interp 'box' (radtran_3D_ch4.py:83-88): the
    coordinates are the band START latitudes, one per row of values, the last band ending at the
    pole - or the n+1 band EDGES; with interp 'lin' (radtran_3Dvs2D_radtrans_new.py:82-87) they are
    band centres and values are interpolated linearly between them, constant outside (host
    evaluation only: the device step builder works on bands and refuses such a profile).  sza
    coordinates are nodes in degrees (linear in between, clamped outside); altitude is linear,
    log-linear ('exp') or box."""

    def __init__(self, grid, values, profname, interp):
        self.grid = grid
        self.names = []
        self.interp = dict()
        self.lat_interp = dict()
        self.values = dict()
        self.add_profile(values, profname, interp)

    def add_profile(self, values, profname, interp='lin'):
        self.names.append(profname)
        self.values[profname] = np.asarray(values, dtype=float)
        per_axis = [interp] if isinstance(interp, str) else list(interp)
        self.interp[profname] = per_axis[-1]
        has_lat = 'lat' in self.grid.names and len(per_axis) > 1
        self.lat_interp[profname] = per_axis[0] if has_lat else 'box'
        if 'lat' in self.grid.names:
            self.lat_edges()                                     # checks rows against coordinates
        setattr(self, profname, self.values[profname])

    def get(self, name):
        return self.values[name]

    def n_band(self):
        return self.values[self.names[0]].shape[0] if 'lat' in self.grid.names else 1

    def lat_edges(self):
        """The n_band + 1 band edges, from band starts (the reference's own call sites) or edges."""
        lat, rows = self.grid.coords['lat'], self.n_band()
        if len(lat) == rows + 1:
            return lat
        if len(lat) == rows:
            return np.append(lat, 90.0 if lat[-1] < 90.0 else lat[-1] + 1.0)
        raise ValueError('{} latitude coordinates for {} rows of values'.format(len(lat), rows))

    def _band(self, lat):
        if 'lat' not in self.grid.names:
            return None
        edges = self.lat_edges()
        return int(np.clip(np.searchsorted(edges, lat, side='right') - 1, 0, len(edges) - 2))

    def _coords(self, point, sza):
        """(lat, sza, alt) of a Coords point (+ sza keyword) or of a plain coordinate list given in
        the order of the grid's own dimensions ([lat, alt], [lat, sza, alt], ...); a bare number is
        an altitude (radtran_3Dvs2D_radtrans_new.py:243)."""
        if hasattr(point, 'Spherical'):
            lat, lon, alt = point.Spherical()
            return lat, sza, alt
        if np.isscalar(point):
            return 0.0, sza, float(point)
        point = list(point)
        if len(point) == self.grid.n_dim and self.grid.names != ['lat', 'lon', 'alt']:
            c = dict(zip(self.grid.names, point))
            return c.get('lat', 0.0), c.get('sza', sza), c['alt']
        lat, lon, alt = point
        return lat, sza, alt

    def _alt_value(self, v, n, alt):
        z = self.grid.coords['alt']
        if self.interp[n] == 'exp':
            return float(np.exp(np.interp(alt, z, np.log(v))))
        if self.interp[n] == 'box':
            return float(v[int(np.clip(np.searchsorted(z, alt, side='right') - 1, 0, len(z) - 1))])
        return float(np.interp(alt, z, v))

    def _row_value(self, v, n, sza, alt):
        """Value of one latitude row (or of a profile without latitude) at (sza, alt)."""
        if 'sza' in self.grid.names:
            nodes = self.grid.coords['sza']
            if sza is None:
                raise ValueError('profile %s depends on the solar zenith angle: pass sza' % n)
            return float(np.interp(sza, nodes, [self._alt_value(v[j], n, alt) for j in range(len(nodes))]))
        return self._alt_value(v, n, alt)

    def calc(self, point, profname=None, sza=None):
        lat, sza, alt = self._coords(point, sza)
        names = [profname] if profname is not None else self.names
        res = dict()
        for n in names:
            v = self.values[n]
            if 'lat' not in self.grid.names:
                res[n] = self._row_value(v, n, sza, alt)
            elif self.lat_interp.get(n, 'box') == 'lin':
                cen = self.grid.coords['lat'][:v.shape[0]]
                j = int(np.clip(np.searchsorted(cen, lat, side='right') - 1, 0, max(len(cen) - 2, 0)))
                if len(cen) == 1:
                    res[n] = self._row_value(v[0], n, sza, alt)
                else:
                    w = float(np.clip((lat - cen[j]) / (cen[j + 1] - cen[j]), 0.0, 1.0))
                    res[n] = (1.0 - w) * self._row_value(v[j], n, sza, alt) + \
                        w * self._row_value(v[j + 1], n, sza, alt)
            else:
                res[n] = self._row_value(v[self._band(lat)], n, sza, alt)
        if profname is None and len(self.names) == 1:
            return res[self.names[0]]           # single-field profile: the number itself, as the
            # reference's drivers use it (radtran_3D_ch4.py:139, radtran_3Dvs2D_radtrans_new.py:243)
        return res[profname] if profname is not None else res

    def __add__(self, other):
        """New profile: every field plus a number or plus the other profile's (single / same-named)
        field."""
        out = copy.deepcopy(self)
        for n in out.names:
            add = other if np.isscalar(other) else other.values[n if n in other.values else other.names[0]]
            out.values[n] = out.values[n] + add
            setattr(out, n, out.values[n])
        return out

    __radd__ = __add__

    def __mul__(self, factor):
        """New profile: every field times a number or times the other profile's field."""
        out = copy.deepcopy(self)
        for n in out.names:
            f = factor if np.isscalar(factor) else factor.values[n if n in factor.values else factor.names[0]]
            out.values[n] = out.values[n] * f
            setattr(out, n, out.values[n])
        return out

    __rmul__ = __mul__

    def __iadd__(self, other):
        """prof += maskgrid*value (RetSet.profile, smm:483-489): adds the other profile's single
        field to every field of this one."""
        add = other.values[other.names[0]]
        for n in self.names:
            self.values[n] = self.values[n] + add
            setattr(self, n, self.values[n])
        return self

    def profile_1d(self, name, lat=0.0):
        b = self._band(lat)
        return self.values[name] if b is None else self.values[name][b]


class AtmGridMask(object):
    """Weight of one retrieval parameter over the atmosphere grid (smm:348-350, 355-374): a
    triangle in altitude (alt_triangle, linear), a latitude box (lat_box: coords are the box START
    latitudes, ascending; the last box is open-ended) or their product (merge)."""

    def __init__(self, grid, mask, interp='lin'):
        self.grid = grid
        self.mask = np.asarray(mask, dtype=float)
        self.interp = {'mask': interp if isinstance(interp, str) else interp[-1]}

    def _lat_box(self, lat):
        starts = self.grid.coords['lat']
        return int(np.clip(np.searchsorted(starts, lat, side='right') - 1, 0, len(starts) - 1))

    def calc(self, point, profname=None):
        lat, lon, alt = point.Spherical() if hasattr(point, 'Spherical') else point
        if self.grid.names == ['lat']:
            return float(self.mask[self._lat_box(lat)])
        m = self.mask[self._lat_box(lat)] if self.grid.n_dim > 1 else self.mask
        z = self.grid.coords['alt']
        if self.interp['mask'] == 'box':
            return float(m[int(np.clip(np.searchsorted(z, alt, side='right') - 1, 0, len(z) - 1))])
        return float(np.interp(alt, z, m))

    def merge(self, mask2):
        """Product of an altitude mask and a latitude-box mask -> mask on ('lat', 'alt')
        (LinearProfile_2D, smm:392-396)."""
        alt_m, lat_m = (self, mask2) if self.grid.names == ['alt'] else (mask2, self)
        if alt_m.grid.names != ['alt'] or lat_m.grid.names != ['lat']:
            raise ValueError('merge needs one altitude mask and one latitude mask')
        grid = AtmGrid(['lat', 'alt'], [lat_m.grid.coords['lat'], alt_m.grid.coords['alt']])
        return AtmGridMask(grid, np.outer(lat_m.mask, alt_m.mask), alt_m.interp['mask'])

    def table(self, z, lat_edges=None):
        """[n_band][n_z] (or [n_z]) weights on the atmosphere's altitude grid and latitude bands:
        what the device step builder takes.  The latitude boxes must be the atmosphere's bands."""
        if not np.array_equal(self.grid.coords['alt'], z) or self.interp['mask'] != 'lin':
            raise ValueError('parameter masks must be linear on the atmosphere altitude grid')
        if self.grid.n_dim == 1:
            return self.mask
        if lat_edges is None or not np.array_equal(self.grid.coords['lat'], np.asarray(lat_edges)[:-1]):
            raise ValueError('latitude boxes of the parameter masks must start at the atmosphere band edges')
        return self.mask

    def __mul__(self, value):
        grid = self.grid
        if grid.n_dim > 1:   # AtmProfile keeps band EDGES as latitude coordinates
            lat = grid.coords['lat']
            grid = AtmGrid(['lat', 'alt'], [np.append(lat, 90.0 if lat[-1] < 90.0 else lat[-1] + 1.0),
                                            grid.coords['alt']])
        return AtmProfile(grid, self.mask * float(value), 'mask', self.interp['mask'])

    __rmul__ = __mul__


def AtmProfZeros(grid, profname, interp):
    shape = tuple(len(g) - (1 if n == 'lat' else 0) for n, g in zip(grid.names, grid.grid))
    return AtmProfile(grid, np.zeros(shape), profname, interp)


def mask_profile_grid(maskgrid):
    """AtmGrid of the profile a parameter mask spans (latitude coordinates as band edges)."""
    return (maskgrid * 0.0).grid


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
        m, q, s = extract_quanta_HITRAN(None, None, self.lev_string)
        return q, s

    def equiv(self, string):
        return extract_quanta_HITRAN(None, None, string)[0] == self.minimal_level_string()

    def add_vibtemp(self, profile):
        self.vibtemp = profile

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

    def add_levels(self, lev_strings, energies, vibtemps=None, degeneracies=None, simmetries=None,
                   add_fundamental=False):
        """Levels `lev_NN` in the order given (call sites radtran_3D_ch4.py:236-255,
        run_0607_lut.py:94).  add_fundamental puts the ground state (all quanta zero, energy 0, no
        vibrational temperature: LTE population) in front when the list does not hold it."""
        lev_strings, energies = list(lev_strings), list(energies)
        extra = 0
        if add_fundamental and len(lev_strings) > 0:
            quanta = extract_quanta_HITRAN(self.mol, self.iso, lev_strings[0])[1]
            ground = ' '.join('0' for _ in quanta)
            known = [extract_quanta_HITRAN(self.mol, self.iso, ls_)[0] for ls_ in lev_strings] + \
                [getattr(self, lev).minimal_level_string() for lev in self.levels]
            if ground not in known:
                lev_strings, energies, extra = [ground] + lev_strings, [0.0] + energies, 1
        for i, (ls_, en) in enumerate(zip(lev_strings, energies)):
            j = i - extra                      # index into the caller's optional lists
            name = 'lev_{:02d}'.format(self.n_lev)
            lev = Level(ls_, en,
                        degen=None if degeneracies is None or j < 0 else degeneracies[j],
                        simmetry=None if simmetries is None or j < 0 else simmetries[j])
            if vibtemps is not None and j >= 0 and vibtemps[j] is not None:
                lev.add_vibtemp(vibtemps[j])
                self.is_in_LTE = False
            setattr(self, name, lev)
            self.levels.append(name)
            self.n_lev += 1

    def add_simmetries_levels(self, lines):
        """Collects, per level, the symmetry labels its vibrational quanta appear with in the
        global-quanta strings of `lines` (CH4: '0 0 1 0 1F2' -> '1F2'; call sites
        run_0607_lut.py:95,100).  Levels are matched by quanta only (Level.equiv), so the labels
        are book-keeping: they do not change which lines feed a level.  Returns {level: labels}."""
        found = dict((lev, []) for lev in self.levels)
        for lin in lines:
            if lin.Mol != self.mol or lin.Iso != self.iso:
                continue
            for string in (lin.Up_lev_str, lin.Lo_lev_str):
                ms, _, sym = extract_quanta_HITRAN(self.mol, self.iso, string)
                if not sym:
                    continue
                for lev in self.levels:
                    if getattr(self, lev).minimal_level_string() == ms and sym not in found[lev]:
                        found[lev].append(sym)
        for lev, syms in found.items():
            L = getattr(self, lev)
            L.simmetry = sorted(set(list(L.simmetry) + syms))
        return found

    def has_level(self, lev_string):
        for lev in self.levels:
            if getattr(self, lev).equiv(lev_string):
                return True, lev
        return False, None

    def erase_level(self, lev):
        self.levels.remove(lev)
        delattr(self, lev)

    def level_energies(self):
        return np.array([getattr(self, lev).energy for lev in self.levels])


class Molec(object):
    def __init__(self, mol, name, MM=None):
        self.mol, self.name, self.MM = int(mol), name, MM
        self.all_iso = []
        self.iso_N = 0
        self.abundance = None

    def add_iso(self, num, MM=None, ratio=None, LTE=True):
        nam = 'iso_{:1d}'.format(num)
        setattr(self, nam, IsoMolec(self.mol, num, MM=MM, ratio=ratio, LTE=LTE))
        self.all_iso.append(nam)
        self.iso_N += 1
        return getattr(self, nam)

    def add_all_iso_from_HITRAN(self, lines, add_levels=False, n_max=None):
        """One IsoMolec (LTE, no levels) per isotopologue number of this molecule that occurs in
        `lines` and is not there yet (callers radtran_test_CO.py:118, run_0607_lut.py:107)."""
        isos = sorted(set(int(lin.Iso) for lin in lines if lin.Mol == self.mol))
        for num in isos[:n_max]:
            if 'iso_{:1d}'.format(num) not in self.all_iso:
                self.add_iso(num, LTE=True)

    def add_clim(self, profile):
        self.abundance = profile

    def link_to_atmos(self, atmosphere):
        self.atmosphere = atmosphere


class Planet(object):
    def __init__(self, name, radius, atm_extension):
        self.name, self.radius, self.atm_extension = name, float(radius), float(atm_extension)
        self.gases = dict()
        self.atmosphere = None

    def add_atmosphere(self, atmosphere):
        self.atmosphere = atmosphere

    def add_gas(self, gas):
        self.gases[gas.name] = gas
        gas.link_to_atmos(self.atmosphere)


class Titan(Planet):
    def __init__(self, atm_extension=1500.0):
        Planet.__init__(self, 'Titan', 2575.0, atm_extension)
        self.mass = 1.3452e23

    def add_default_atm(self):
        """A smooth single-band Titan-like atmosphere up to atm_extension (temp linear, pres
        log-linear in altitude): the analytic profiles of spectrobot_b200.synthetic.  (The
        reference reads its default climatology from files that are not shipped;
        spect_robot.py:22.)"""
        from . import synthetic as S
        atm = S.titan_atmosphere(n_bands=1, z_top=self.atm_extension)
        prof = AtmProfile(AtmGrid('alt', atm["z"]), atm["temp"][0], 'temp', 'lin')
        prof.add_profile(atm["pres"][0], 'pres', 'exp')
        self.add_atmosphere(prof)
        return prof


# ---------------------------------------------------------------------------------------------
# geometry
# ---------------------------------------------------------------------------------------------
class Coords(object):
    """Point in planetocentric coordinates: 'Spherical' = (lat deg, lon deg, alt km above R) or
    'Cartesian' = (x, y, z) km (spect_robot.py:24-27)."""

    def __init__(self, vec, s_ref='Spherical', R=2575.0):
        self.R = R
        v = np.asarray(vec, dtype=float)
        if s_ref == 'Spherical':
            lat, lon, alt = v
            r = R + alt
            self.cart = np.array([r * mt.cos(rad(lat)) * mt.cos(rad(lon)),
                                  r * mt.cos(rad(lat)) * mt.sin(rad(lon)), r * mt.sin(rad(lat))])
        else:
            self.cart = v.copy()

    def Cartesian(self):
        return self.cart.copy()

    def Spherical(self):
        x, y, z = self.cart
        r = mt.sqrt(x * x + y * y + z * z)
        return np.array([mt.degrees(mt.asin(z / r)), mt.degrees(mt.atan2(y, x)), r - self.R])

    def distance(self, other):
        return float(np.linalg.norm(self.cart - other.cart))


class LineOfSight(object):
    """Straight ray from `start` (observer) through `second_point`."""

    def __init__(self, start, second_point, tag=None):
        self.starting_point = start
        self.second_point = second_point
        d = second_point.Cartesian() - start.Cartesian()
        self.direction = d / np.linalg.norm(d)
        self.tag = tag
        self.intersections = None
        self.szas = None
        self.radtran_steps = None
        self.involved_retparams = dict()

    def details(self):
        print('LOS from {} towards {}'.format(self.starting_point.Spherical(),
                                              self.second_point.Spherical()))

    def _s_tangent(self):
        return float(-np.dot(self.starting_point.Cartesian(), self.direction))

    def get_tangent_point(self):
        p = self.starting_point.Cartesian() + self._s_tangent() * self.direction
        return Coords(p, s_ref='Cartesian', R=self.starting_point.R)

    def get_tangent_altitude(self):
        return float(self.get_tangent_point().Spherical()[2])

    @property
    def tangent_altitude(self):
        return self.get_tangent_altitude()

    def calc_atm_intersections(self, planet, delta_x=5.0, start_from_TOA=True,
                               stop_at_second_point=False, LOS_order='radtran', verbose=False):
        """Points along the ray inside the atmosphere, every delta_x km.  The order
        of the points is the order of the layer recursion.  LOS_order 'radtran' (default): from the
        far end of the ray towards the starting point (the observer).  LOS_order 'photon': from the
        starting point's side towards the far end - photons that enter where the ray starts (the
        solar-ray test, radtran_3D_ch4.py:311; `invert_LOS_direction`, smm:3135-3137).
        start_from_TOA=False starts at the starting point when it lies inside the atmosphere,
        stop_at_second_point=True ends at the second point."""
        if LOS_order not in ('radtran', 'photon'):
            raise ValueError("LOS_order must be 'radtran' or 'photon'")
        self.LOS_order = LOS_order
        o = self.starting_point.Cartesian()
        st = self._s_tangent()
        rt = float(np.linalg.norm(o + st * self.direction))
        r_top = planet.radius + planet.atm_extension
        if rt >= r_top:
            self.intersections = []
            self._s = np.zeros(0)
            return self.intersections
        half = mt.sqrt(r_top ** 2 - rt ** 2)
        s_near, s_far = st - half, st + half
        if rt < planet.radius:                       # ray hits the surface: stop there
            s_far = st - mt.sqrt(planet.radius ** 2 - rt ** 2)
        if not start_from_TOA:
            s_near = max(s_near, 0.0)
        if stop_at_second_point:
            s_far = min(s_far, float(np.linalg.norm(self.second_point.Cartesian() - o)))
        if s_far <= s_near:
            self.intersections = []
            self._s = np.zeros(0)
            return self.intersections
        # samples anchored on the tangent point, so that no two ADJACENT samples sit at the same
        # altitude (curgod_fort_* divide by log(nd2/nd1), curgods.f:17)
        kmax = int(mt.floor(half / delta_x - 1e-9))
        inner = st + delta_x * np.arange(kmax, -kmax - 1, -1)
        inner = inner[(inner < s_far - 1e-6) & (inner > s_near + 1e-6)]
        s = np.concatenate([[s_far], inner, [s_near]])
        if LOS_order == 'photon':
            s = s[::-1]
        self._s = s
        self.intersections = [Coords(o + si * self.direction, s_ref='Cartesian', R=planet.radius)
                              for si in s]
        return self.intersections

    def calc_SZA_along_los(self, planet, sub_solar_point):
        sun = sub_solar_point.Cartesian()
        sun = sun / np.linalg.norm(sun)
        self.szas = np.array([mt.degrees(mt.acos(np.clip(np.dot(p.Cartesian(), sun) /
                                                         np.linalg.norm(p.Cartesian()), -1, 1)))
                              for p in self.intersections])
        return self.szas

    def calc_along_LOS(self, atmosphere, profname=None, set_attr=False, set_attr_name=None):
        vals = np.array([atmosphere.calc(p, profname=profname) for p in self.intersections])
        if set_attr:
            setattr(self, set_attr_name or profname, vals)
        return vals

    def calc_radtran_steps(self, planet, lines=None, queue=None, calc_derivatives=False,
                           bayes_set=None, max_T_variation=5.0, max_Plog_variation=1.0,
                           max_opt_depth=None):
        """Merge the intersection points into radtran steps: consecutive points are
        merged while T varies by < max_T_variation, ln P by < max_Plog_variation and (with
        max_opt_depth and `lines`) the estimated optical depth sum_gas sigma_peak*column stays below
        max_opt_depth; every step gets air-weighted Curtis-Godson T and P (curgod_fort_1/4), the
        gas columns (curgod_fort_2) and the column-weighted vibrational temperature of every level
        (curgod_fort_3), evaluated at the solar zenith angles self.szas when the level's profile
        depends on the SZA (calc_SZA_along_los, or a constant with use_tangent_sza)."""
        pts = self.intersections
        n = len(pts)
        atm = planet.atmosphere
        if n == 0:                       # the ray misses the atmosphere: no steps
            self.radtran_steps = {'step': [], 'gas_isos': [(g, iso) for g in sorted(planet.gases)
                                                            for iso in planet.gases[g].all_iso],
                                  'opt': dict(max_T_variation=max_T_variation,
                                              max_Plog_variation=max_Plog_variation,
                                              max_opt_depth=max_opt_depth)}
            if calc_derivatives and bayes_set is not None:
                for par in bayes_set.params():
                    self.involved_retparams[(par.nameset, par.key)] = False
            if queue is not None:
                queue.put(self)
            return self
        T = np.array([atm.calc(p, 'temp') for p in pts])
        P = np.array([atm.calc(p, 'pres') for p in pts])
        nd = P / (kb_hpa * T)
        x = np.abs(self._s[0] - self._s) * 1.e5                 # path length from the first point, cm
        lnP = np.log(P)
        sigma = None
        if max_opt_depth is not None and max_opt_depth > 0.0:
            if not lines:
                raise ValueError('max_opt_depth needs the line list (peak cross-sections)')
            sigma = peak_cross_sections(planet, lines)
            vmr_all = dict((g, np.array([planet.gases[g].abundance.calc(p, 'vmr') for p in pts]))
                           for g in planet.gases)
        bounds, i0 = [], 0
        for i in range(1, n):
            seg = slice(i0, i + 1)
            too_thick = False
            if sigma is not None:
                tau = sum(sigma[g] * curgods.curgod_fort_2(nd[seg], vmr_all[g][seg], x[seg], i + 1 - i0)
                          for g in planet.gases)
                too_thick = tau > max_opt_depth
            if (T[seg].max() - T[seg].min() > max_T_variation or
                    lnP[seg].max() - lnP[seg].min() > max_Plog_variation or too_thick) and i - i0 >= 2:
                bounds.append((i0, i - 1))
                i0 = i - 1
        if n >= 2:
            bounds.append((i0, n - 1))
        steps = []
        gas_isos = [(g, iso) for g in sorted(planet.gases) for iso in planet.gases[g].all_iso]
        # retrieval parameters that are VMR nodes of a gas of this planet: the
        # column of a step depends on parameter p through d u/dp = int n * mask_p ds
        jac_pars = []
        if calc_derivatives and bayes_set is not None:
            for par in bayes_set.params():
                self.involved_retparams[(par.nameset, par.key)] = False
                if par.nameset in planet.gases:
                    jac_pars.append((par, np.array([par.maskgrid.calc(p) for p in pts])))
        for a, e in bounds:
            sl = slice(a, e + 1)
            npnt = e - a + 1
            ones = np.ones(npnt)
            air = curgods.curgod_fort_1(nd[sl], x[sl], npnt)
            st = {'range': (a, e), 'air_column': air,
                  'temp': curgods.curgod_fort_4(nd[sl], ones, T[sl], x[sl], npnt) / air,
                  'pres': curgods.curgod_fort_4(nd[sl], ones, P[sl], x[sl], npnt) / air,
                  'columns': dict(), 'vibtemps': dict(),
                  'sza': None if self.szas is None else float(np.mean(self.szas[sl]))}
            for g in sorted(planet.gases):
                gas = planet.gases[g]
                vmr = np.array([gas.abundance.calc(p, 'vmr') for p in pts[a:e + 1]])
                col = curgods.curgod_fort_2(nd[sl], vmr, x[sl], npnt)
                st['columns'][g] = col
                for iso in gas.all_iso:
                    im = getattr(gas, iso)
                    for lev in im.levels:
                        L = getattr(im, lev)
                        if L.vibtemp is None:
                            tv = st['temp']
                        else:
                            if 'sza' in L.vibtemp.grid.names:
                                if self.szas is None:
                                    raise ValueError('SZA-dependent vibrational temperatures: call '
                                                     'calc_SZA_along_los (or set los.szas) first')
                                tvp = np.array([L.vibtemp.calc(p, 'vibtemp', sza=sz)
                                                for p, sz in zip(pts[a:e + 1], self.szas[a:e + 1])])
                            else:
                                tvp = np.array([L.vibtemp.calc(p, 'vibtemp') for p in pts[a:e + 1]])
                            tv = curgods.curgod_fort_3(nd[sl], vmr, tvp, x[sl], npnt) / col
                        st['vibtemps'][(g, iso, lev)] = tv
            if jac_pars:
                st['dcolumns'] = dict()
                for par, mvals in jac_pars:
                    m = mvals[sl]
                    d = curgods.curgod_fort_2(nd[sl], m, x[sl], npnt) if np.any(m != 0.0) else 0.0
                    st['dcolumns'][(par.nameset, par.key)] = d
                    if d != 0.0:
                        self.involved_retparams[(par.nameset, par.key)] = True
            steps.append(st)
        self.radtran_steps = {'step': steps, 'gas_isos': gas_isos,
                              'opt': dict(max_T_variation=max_T_variation,
                                          max_Plog_variation=max_Plog_variation,
                                          max_opt_depth=max_opt_depth)}
        if queue is not None:
            queue.put(self)
        return self

    def step_tables(self, planet, n_steps_max=None):
        """Arrays for the C ABI (sr_los_steps) from self.radtran_steps, one gas-iso per LUT."""
        steps = self.radtran_steps['step']
        gi = self.radtran_steps['gas_isos']
        ns = len(steps)
        nmax = ns if n_steps_max is None else n_steps_max
        n_lev = max([len(getattr(planet.gases[g], iso).levels) for g, iso in gi] + [1])
        temp = np.full(nmax, 100.0)
        pres = np.full(nmax, 1.e-6)
        col = np.zeros((len(gi), nmax))
        tvib = np.full((len(gi), n_lev, nmax), 100.0)
        for k, st in enumerate(steps):
            temp[k], pres[k] = st['temp'], st['pres']
            for m, (g, iso) in enumerate(gi):
                col[m, k] = st['columns'][g]
                im = getattr(planet.gases[g], iso)
                for j, lev in enumerate(im.levels):
                    tvib[m, j, k] = st['vibtemps'][(g, iso, lev)]
        return ns, temp, pres, col, tvib

    def radtran_fast(self, sp_grid, planet, queue=None, cartLUTs=None, cartDROP=None,
                     calc_derivatives=False, bayes_set=None, LUTS=None, radtran_opt=None,
                     store_abscoeff=False, track_levels=None, solo_absorption=False,
                     initial_intensity=None):
        """Hi-res radiance of this LOS on sp_grid from device-resident LUTs (one-LOS batch through
        the same C ABI entry point smm.radtrans uses for whole batches).  Returns
        [SpectralIntensity, single_rads, bayes_set copy] like the reference
        (spect_main_module.py:3214-3228): single_rads = {(gas, iso[, lev]): hi-res contribution of
        that emitter} (smm.single_radiances); with calc_derivatives the hi-res derivative spectra
        are attached to the parameters of the returned copy (par.hires_deriv, :2867-2874)."""
        from . import spect_main_module as smm
        rads = smm.los_batch_radiances([self], sp_grid, planet, LUTS,
                                       solo_absorption=solo_absorption,
                                       initial_intensity=initial_intensity)
        single = dict()
        if not solo_absorption:
            was = self.tag
            self.tag = was if was is not None else 'LOS'
            sr_ = smm.single_radiances([self], sp_grid, planet, LUTS, None, None, 'same', 1.0, None,
                                       track_levels, solo_absorption=solo_absorption,
                                       initial_intensity=None)
            single = dict((k, v[self.tag]) for k, v in sr_.items())
            self.tag = was
        out = [rads[0], single, copy.deepcopy(bayes_set)]
        if calc_derivatives:
            rad, jac = smm.los_batch_jacobians([self], sp_grid, planet, LUTS, bayes_set,
                                               solo_absorption=solo_absorption,
                                               initial_intensity=initial_intensity)
            out[0] = rad[0]
            for par, der in zip(out[2].params(), jac[0]):
                par.add_hires_deriv(der)
        if queue is not None:
            queue.put(out)
        return out

    def radtran(self, wn_range, planet, lines, cartLUTs=None, cartDROP=None, calc_derivatives=False,
                bayes_set=None, LUTS=None, useLUTs=True, radtran_opt=None, g3D=False,
                sub_solar_point=None, solo_absorption=False, initial_intensity=None,
                track_levels=None, sp_grid=None, LUTopt=None):
        """The single-LOS forward model of the reference's slow path (callers smm:2506-2545,
        radtran_test_CO.py:194 through smm.inversion, radtran_3D_ch4.py:312, 336): intersections
        (kept if already computed, e.g. with LOS_order='photon'), SZA along the LOS when g3D,
        radtran steps, then the hi-res radiance over wn_range.  useLUTs=True interpolates the
        look-up tables (built for this atmosphere when LUTS is None); useLUTs=False evaluates the
        cross-sections line by line at every step's own Curtis-Godson (P, T) - no table, no
        interpolation.  Returns [SpectralIntensity, single_rads, bayes_set copy]; with
        calc_derivatives the hi-res derivative spectra are ALSO attached to the parameters of the
        bayes_set that was passed in, which is where smm.inversion reads them (:2508-2515)."""
        from . import spect_main_module as smm
        radtran_opt = dict(radtran_opt or {})
        if self.intersections is None:
            self.calc_atm_intersections(planet)
        if g3D:
            if sub_solar_point is None:
                raise ValueError('g3D needs the sub-solar point')
            self.calc_SZA_along_los(planet, sub_solar_point)
        self.calc_radtran_steps(planet, lines, calc_derivatives=calc_derivatives, bayes_set=bayes_set,
                                **radtran_opt)
        if sp_grid is None:
            sp_grid = smm.prepare_spe_grid(wn_range).spectral_grid
        if initial_intensity is not None and hasattr(initial_intensity, 'spectrum'):
            initial_intensity = initial_intensity.spectrum
        if not useLUTs:
            out = smm.los_radiance_line_by_line(self, sp_grid, planet, lines,
                                                calc_derivatives=calc_derivatives, bayes_set=bayes_set,
                                                solo_absorption=solo_absorption,
                                                initial_intensity=initial_intensity)
        else:
            if LUTS is None:
                gases = list(planet.gases.values())
                opt = dict(LUTopt or {})
                if 'max_pres' not in opt:
                    opt['max_pres'] = max(st['pres'] for st in self.radtran_steps['step']) * 1.01
                PT = smm.calc_PT_couples_atmosphere(lines, gases, planet.atmosphere, **opt)
                LUTS = smm.check_and_build_allluts(dict(cart_LUTS=cartLUTs), sp_grid, lines, gases,
                                                   PTcouples=PT, LUTopt=opt)
            out = self.radtran_fast(sp_grid, planet, cartLUTs=cartLUTs, cartDROP=cartDROP,
                                    calc_derivatives=calc_derivatives, bayes_set=bayes_set, LUTS=LUTS,
                                    radtran_opt=radtran_opt, track_levels=track_levels,
                                    solo_absorption=solo_absorption,
                                    initial_intensity=initial_intensity)
        if calc_derivatives and bayes_set is not None:
            for par, par_mod in zip(bayes_set.params(), out[2].params()):
                par.hires_deriv = par_mod.hires_deriv
                par.not_involved = not self.involved_retparams.get((par.nameset, par.key), False)
        return out


def peak_cross_sections(planet, lines):
    """{gas name: largest line-centre absorption cross-section estimate, cm2 per molecule of the
    gas}: max over the gas' lines of Strength(296 K, abundance included) / (Doppler HWHM at 296 K *
    sqrt(pi / ln 2)) - the scale the `max_opt_depth` step limit multiplies the gas columns with
    (DESIGN.md 6.1)."""
    from . import spect_classes as spcl
    out = dict()
    for g, gas in planet.gases.items():
        best = 0.0
        for iso in gas.all_iso:
            im = getattr(gas, iso)
            for lin in lines:
                if lin.Mol == im.mol and lin.Iso == im.iso:
                    dw = spcl.Doppler_width(T_ref, im.MM, lin.Freq)
                    best = max(best, lin.Strength / (dw * mt.sqrt(mt.pi / mt.log(2.0))))
        out[g] = best
    return out


class VIMSPixel(object):
    """Synthetic VIMS limb pixel: observer position + tangent geometry; three LOS per pixel
    (low, centre, up: spect_main_module.py:3091-3096)."""

    def __init__(self, keys, values):
        for k, v in zip(keys, values):
            setattr(self, k, v)
        self.pixel_rot = getattr(self, 'pixel_rot', 0.0)
        self.fov_km = getattr(self, 'fov_km', 24.0)       # full vertical extent of the pixel
        self.observation = getattr(self, 'observation', None)

    def Spacecraft(self):
        return Coords([self.sub_obs_lat, self.sub_obs_lon, self.dist], s_ref='Spherical')

    def _los_at(self, alt):
        tg = Coords([self.limb_tg_lat, self.limb_tg_lon, alt], s_ref='Spherical')
        return LineOfSight(self.Spacecraft(), tg)

    def LOS(self):
        return self._los_at(self.limb_tg_alt)

    def low_LOS(self):
        return self._los_at(self.limb_tg_alt - self.fov_km / 2.0)

    def up_LOS(self):
        return self._los_at(self.limb_tg_alt + self.fov_km / 2.0)

    def sub_solar_point(self):
        return Coords([self.sub_solar_lat, self.sub_solar_lon, 0.0], s_ref='Spherical')
