#!/usr/bin/python
# -*- coding: utf-8 -*-
"""Python-3 copy of the reference driver radtran_3D_ch4.py (BASELINE.json configs[1]): CH4 3.3 um
non-LTE limb radiances of a 3-D Titan (7 latitude bands, vibrational temperatures on (lat, SZA,
alt)), a 2-D linear VMR parameter space, LUTs, and the fast limb retrieval
smm.inversion_fast_limb(..., g3D=True, group_observations=False) (radtran_3D_ch4.py:349).
What differs from the original: Python 3, no absolute paths, synthetic inputs
(examples/synthetic_inputs.py) for the un-shipped climatology / T_vib / HITRAN / VIMS files; the
"observed" spectra are simulated with a known CH4 profile so that the retrieval has a truth.
SR_EXAMPLE_SMALL=1 shrinks the spectral range and line / pixel counts (used by the tests)."""
import copy
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectrobot_b200  # noqa: E402
spectrobot_b200.install_reference_names()

import spect_base_module as sbm  # noqa: E402
import spect_classes as spcl  # noqa: E402
import lineshape  # noqa: E402,F401
import spect_main_module as smm  # noqa: E402
import synthetic_inputs as syn  # noqa: E402


def main(small=bool(int(os.environ.get('SR_EXAMPLE_SMALL', '0')))):
    time0 = time.time()
    base, cart_LUTS, out_dir = syn.work_dirs('radtran_3D_ch4')
    inputs = dict(cart_LUTS=cart_LUTS, out_dir=out_dir, n_threads=8, test=False, n_split=None)

    ### LOADING PLANET
    print('Loading planet...')
    planet = sbm.Titan(1500.)
    lat_ext = syn.LAT_EXT
    grid, Atm, atm = syn.atmosphere(n_bands=7)
    planet.add_atmosphere(Atm)

    ### LOADING MOLECULES
    print('Loading molecules...')
    sza_nodes = syn.S.SZA_NODES
    n_lev_ch4, n_lev_hcn = (5, 3) if small else (15, 11)
    e_ch4 = np.concatenate([syn.S.CH4_LEVEL_ENERGIES, [3130., 3160., 3190.]])[:n_lev_ch4]
    e_hcn = np.array([0.0, 712.0, 1412.0, 2097.0, 2113.5, 2806.0, 3311.5, 3502.0, 4004.0, 4173.0, 4699.0])[:n_lev_hcn]
    nlte_molecs = dict()
    nlte_molecs['CH4'] = syn.nlte_molec(6, 'CH4', atm, e_ch4, 7, sza_nodes=sza_nodes)
    nlte_molecs['CH4'].add_iso(2, LTE=True)
    nlte_molecs['HCN'] = syn.nlte_molec(23, 'HCN', atm, e_hcn, 7, sza_nodes=sza_nodes)
    atm_gases_old = dict(CH4=syn.vmr_profile(grid, atm, 0.015, 7), HCN=syn.vmr_profile(grid, atm, 2.e-6, 7))
    for molec in nlte_molecs.values():
        molec.link_to_atmos(Atm)
        molec.add_clim(atm_gases_old[molec.name])
        planet.add_gas(molec)

    ##### SETTING THE BAYESSET:
    baybau = smm.BayesSet(tag='test_CH4_HCN_3D')
    alt_nodes = np.arange(350., 1050., 100.)
    lat_limits = lat_ext[:-1]
    for gasname in (['CH4'] if small else ['CH4', 'HCN']):
        cososo = atm_gases_old[gasname]
        apriori_profs = np.array([[cososo.calc([lat, alt], 'vmr') for alt in alt_nodes] for lat in lat_limits])
        apriori_prof_errs = 0.7 * apriori_profs
        set_ = smm.LinearProfile_2D(gasname, planet.atmosphere, alt_nodes, lat_limits, apriori_profs, apriori_prof_errs)
        baybau.add_set(set_)

    ### updating the profile of gases in bayesset
    for gas in baybau.sets.keys():
        planet.gases[gas].add_clim(baybau.sets[gas].profile())

    wn_range = [2996., 3004.] if small else [2850., 3450.]
    wn_range_obs = [spcl.convertto_nm(wn_range[1], 'cm_1') + 10., spcl.convertto_nm(wn_range[0], 'cm_1') - 10.]
    print(wn_range_obs)

    radtran_opt = dict()
    radtran_opt['max_T_variation'] = 5.
    radtran_opt['max_Plog_variation'] = 1.

    tangents = [420., 610., 800.] if small else list(np.arange(360., 1041., 40.))
    pixels = syn.observed_pixels(tangents, wn_range, 12 if small else 36, lat=12.0,
                                 sza=np.linspace(35., 75., len(tangents)))
    pixels = [pix for pix in pixels if pix.limb_tg_alt > 350.]
    pix_ok = pixels

    print('Loading lines...')
    n_ch4, n_hcn = (160, 40) if small else (30000, 2000)
    db_file = syn.write_hitran_file(os.path.join(base, 'synthetic_hitran.par'), wn_range, [
        dict(mol=6, iso=1, n_lines=n_ch4, level_energies=e_ch4, q296=590.52, ratio=nlte_molecs['CH4'].iso_1.ratio),
        dict(mol=6, iso=2, n_lines=n_ch4 // 10, level_energies=None, q296=1180.8, ratio=nlte_molecs['CH4'].iso_2.ratio),
        dict(mol=23, iso=1, n_lines=n_hcn, level_energies=e_hcn, q296=892.2, ratio=nlte_molecs['HCN'].iso_1.ratio)])
    linee = spcl.read_line_database(db_file, freq_range=wn_range)
    linee = [lin for lin in linee if lin.Freq >= wn_range[0] and lin.Freq <= wn_range[1]]
    linee = smm.check_lines_mols(linee, planet.gases.values())
    smm.keep_levels_wlines(planet, linee)

    keep_levels = dict()
    keep_levels[('CH4', 'iso_1')] = ['lev_00', 'lev_01', 'lev_02', 'lev_09', 'lev_07', 'lev_14', 'lev_08', 'lev_06', 'lev_03', 'lev_05', 'lev_04', 'lev_10']
    keep_levels[('CH4', 'iso_2')] = []
    keep_levels[('HCN', 'iso_1')] = ['lev_00', 'lev_01', 'lev_02', 'lev_04', 'lev_10', 'lev_07']
    smm.keep_levels(planet, keep_levels)
    linee = smm.check_lines_mols(linee, planet.gases.values())
    for gas in planet.gases:
        for iso in planet.gases[gas].all_iso:
            print([gas, iso], getattr(planet.gases[gas], iso).levels)

    LUTopt = dict()
    LUTopt['max_pres'] = 2.0   # hPa circa 200 km
    LUTopt['temp_step'] = 5.
    LUTopt['pres_step_log'] = 1.0

    sp_gri = smm.prepare_spe_grid(wn_range).spectral_grid
    PTcoup_needed = smm.calc_PT_couples_atmosphere(linee, planet.gases.values(), planet.atmosphere, **LUTopt)
    LUTS = smm.check_and_build_allluts(inputs, sp_gri, linee, planet.gases.values(), PTcouples=PTcoup_needed, LUTopt=LUTopt)
    print('{} PT couples, LUTs ready after {:6.2f} s'.format(len(PTcoup_needed), time.time() - time0))

    # "observations": the forward model with the true profile (a-priori scaled by 1.25 in CH4)
    truth = copy.deepcopy(baybau)
    for par in truth.sets['CH4'].set:
        par.value = 1.25 * par.apriori
    sims_true, _, _ = smm.radtrans(inputs, planet_with(planet, truth), linee, copy.deepcopy(pix_ok), wn_range=wn_range,
                                   radtran_opt=radtran_opt, LUTopt=LUTopt, save_hires=False, nome_inv='truth')
    syn.set_observations(pix_ok, sims_true)

    dampa = open(os.path.join(out_dir, 'out_3D_inversion_test_fast.pic'), 'wb')
    print(len(pix_ok))
    print([pix.limb_tg_alt for pix in pix_ok])

    result = smm.inversion_fast_limb(inputs, planet, linee, baybau, pix_ok, wn_range=wn_range, radtran_opt=radtran_opt, debugfile=dampa, LUTopt=LUTopt, g3D=True, group_observations=False, max_it=3 if small else 10)

    dampa.close()
    tot_time = time.time() - time0
    print('Tempo totale: {} min'.format(tot_time / 60.))
    print('Tempo una LOS: {} min'.format(tot_time / (3. * len(pix_ok)) / 60.))
    return result, truth, sims_true, planet, linee, pix_ok


def planet_with(planet, bayes_set):
    """A copy of the planet carrying the VMR profiles of bayes_set."""
    pl = copy.deepcopy(planet)
    for gas in bayes_set.sets.keys():
        pl.gases[gas].add_clim(bayes_set.sets[gas].profile())
    return pl


if __name__ == '__main__':
    main()
