"""Synthetic stand-ins for the INPUT FILES of the reference's drivers (HITRAN line lists,
p-T / VMR / vibrational-temperature profiles, VIMS pixel cubes): none of them is shipped with the
reference.  Everything is built through the same object API the file readers of the
missing spect_base_module would feed (Molec / IsoMolec / Level / AtmProfile / VIMSPixel), so the
driver-shaped scripts next to this file keep the reference's flow from "LOADING PLANET" on."""
import os

import numpy as np

import spect_base_module as sbm
import spect_classes as spcl
from spectrobot_b200 import synthetic as S

LAT_EXT = [-90., -75., -60., -30., 30., 60., 75., 90.]          # radtran_3D_ch4.py:70-72


def write_hitran_file(path, wn_range, specs, seed=20067):
    """A HITRAN2012-format (160 column) line file.  specs: list of dict(mol, iso, n_lines,
    level_energies or None (LTE, no level assignment), q296, ratio)."""
    lines = []
    for k, sp in enumerate(specs):
        n_lev = 1 if sp.get('level_energies') is None else len(sp['level_energies'])
        tab = S.line_table(sp['n_lines'], wn_range[0] - 3.3, wn_range[1] + 3.3, n_levels=n_lev,
                           q296=sp['q296'], iso_ratio=sp['ratio'], seed=seed + k,
                           level_energies=sp.get('level_energies'))
        lines += S.spect_lines(tab, mol=sp['mol'], iso=sp['iso'])
    lines.sort(key=lambda l: l.Freq)
    with open(path, 'w') as f:
        for lin in lines:
            f.write(spcl.format_line_record(lin) + '\n')
    return path


def lat_coords(n_bands, lat_interp='box'):
    """Latitude coordinates of the profiles: the band edges for 'box' (radtran_3D_ch4.py:83), the
    band centres for 'lin' (radtran_3Dvs2D_radtrans_new.py:72,82)."""
    ext = LAT_EXT[:n_bands + 1]
    return ext if lat_interp == 'box' else [(a + b) / 2.0 for a, b in zip(ext[:-1], ext[1:])]


def atmosphere(n_bands=7, z_top=1500.0, lat_interp='box'):
    """(AtmGrid, AtmProfile 'temp'+'pres', altitude grid): the p-T climatology files."""
    atm = S.titan_atmosphere(n_bands=n_bands, z_top=z_top)
    if n_bands == 1:
        grid = sbm.AtmGrid('alt', atm['z'])
        prof = sbm.AtmProfile(grid, atm['temp'][0], 'temp', 'lin')
        prof.add_profile(atm['pres'][0], 'pres', 'exp')
    else:
        grid = sbm.AtmGrid(['lat', 'alt'], [lat_coords(n_bands, lat_interp), atm['z']])
        prof = sbm.AtmProfile(grid, atm['temp'], 'temp', [lat_interp, 'lin'])
        prof.add_profile(atm['pres'], 'pres', [lat_interp, 'exp'])
    return grid, prof, atm


def vmr_profile(grid, atm, value, n_bands, lat_interp='box'):
    shape = (len(atm['z']),) if n_bands == 1 else (n_bands, len(atm['z']))
    return sbm.AtmProfile(grid, np.full(shape, value), 'vmr', 'lin' if n_bands == 1 else [lat_interp, 'lin'])


def nlte_molec(mol, name, atm, level_energies, n_bands, sza_nodes=None, lat_interp='box'):
    """A Molec whose iso_1 carries vibrational levels with T_vib profiles (the vt_* files read by
    add_nLTE_molecs_from_tvibmanuel[_3D]); 3-D (lat, SZA, alt) with sza_nodes; lat_interp as the
    reader's keyword of the same name."""
    gas = sbm.Molec(mol, name)
    im = gas.add_iso(1, LTE=False)
    z = atm['z']
    lats = lat_coords(n_bands, lat_interp)
    if sza_nodes is None:
        tv = np.stack([S.vib_temperatures(z, atm['temp'][b], level_energies, 60.0)
                       for b in range(n_bands)], axis=1)                  # [lev][band][z]
        g = sbm.AtmGrid('alt', z) if n_bands == 1 else sbm.AtmGrid(['lat', 'alt'], [lats, z])
        interp = 'lin' if n_bands == 1 else [lat_interp, 'lin']
    else:
        tv = S.vib_temperatures_3d(z, atm['temp'][:n_bands], level_energies, sza_nodes)
        g = (sbm.AtmGrid(['sza', 'alt'], [sza_nodes, z]) if n_bands == 1 else
             sbm.AtmGrid(['lat', 'sza', 'alt'], [lats, sza_nodes, z]))
        interp = 'lin' if n_bands == 1 else [lat_interp, 'lin', 'lin']
    profs = [sbm.AtmProfile(g, t[0] if n_bands == 1 else t, 'vibtemp', interp) for t in tv]
    im.add_levels(S.level_strings(len(level_energies)), level_energies, vibtemps=profs)
    return gas


def observed_pixels(tangent_km, wn_range, n_chan, lat=10.0, sza=60.0, units='nm', obs_units='Wm2'):
    """VIMS-like pixels (read_input_observed): channels across the range on a wavelength axis in
    nm (smm:2372), Gaussian widths of one channel spacing, unit mask, noise set later."""
    c_cm = np.linspace(wn_range[0] + 1.0, wn_range[1] - 1.0, n_chan)
    w_cm = np.full(n_chan, (c_cm[1] - c_cm[0]))
    if units == 'nm':
        centres = np.sort(1.e7 / c_cm)
        widths = (w_cm * 1.e7 / c_cm ** 2)[::-1].copy()
    else:
        centres, widths = c_cm, w_cm
    return S.vims_pixels(tangent_km, lat=lat, channels=centres, widths=widths, units=units,
                         sza=sza, obs_units=obs_units)


def set_observations(pixels, sims, rel_noise=0.01):
    """Use simulated spectra as the 'observed' ones, with a flat noise level."""
    for pix, sim in zip(sorted(pixels, key=lambda p: p.limb_tg_alt), sims):
        pix.observation.spectrum = np.array(sim.spectrum, dtype=float)
        pix.observation.intensity = pix.observation.spectrum
        pix.observation.noise = spcl.SpectralObject(
            np.full(len(sim.spectrum), rel_noise * np.max(sim.spectrum)), pix.observation.spectral_grid)
        pix.observation.mask = np.ones(len(sim.spectrum))


def work_dirs(tag):
    base = os.environ.get('SR_EXAMPLE_DIR', os.path.join('/tmp', 'spectrobot_examples', tag))
    luts, out = os.path.join(base, 'LUTs') + '/', os.path.join(base, 'out') + '/'
    for d in (luts, out):
        os.makedirs(d, exist_ok=True)
    return base, luts, out
