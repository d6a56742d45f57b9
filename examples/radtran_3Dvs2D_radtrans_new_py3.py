#!/usr/bin/python
# -*- coding: utf-8 -*-
"""Python-3 copy of the reference driver radtran_3Dvs2D_radtrans_new.py (BASELINE.json
configs[3]): limb radiances and retrievals of a 3-D Titan whose climatology, VMRs and vibrational
temperatures are interpolated LINEARLY in latitude between the band centres
(`sbm.AtmProfile(grid, TT, 'temp', ['lin','lin'])`, `lat_interp='lin'`, :82-111), compared between
the SZA followed along every LOS ("3D"), the tangent-point SZA for the whole LOS ("2D",
use_tangent_sza) and the inverted LOS direction:

    smm.radtrans(..., save_hires=True, group_observations=True, track_levels=...)       (:367-391)
    smm.inversion_fast_limb(..., debugfile=..., group_observations=True, ...)           (:396-430)

with a 1-D BayesSet of smm.LinearProfile_1D_new sets (:239-275).  What differs from the original:
Python 3, no absolute paths, synthetic inputs (examples/synthetic_inputs.py) for the un-shipped
climatology / T_vib / HITRAN / VIMS files, one SZA sequence instead of three, and the "observed"
spectra are simulated with a known CH4 profile so that the retrievals have a truth.
SR_EXAMPLE_SMALL=1 shrinks the spectral range and the line / pixel counts (used by the tests)."""
import copy
import os
import pickle
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
    base, cart_LUTS, out_dir = syn.work_dirs('radtran_3Dvs2D')
    inputs = dict(cart_LUTS=cart_LUTS, out_dir=out_dir, n_threads=8, test=False, n_split=None)

    ### LOADING PLANET
    print('Loading planet...')
    planet = sbm.Titan(1500.)
    lat_ext = syn.LAT_EXT
    grid, Atm, atm = syn.atmosphere(n_bands=7, lat_interp='lin')
    planet.add_atmosphere(Atm)

    ### LOADING MOLECULES
    print('Loading molecules...')
    sza_nodes = syn.S.SZA_NODES
    n_lev_ch4, n_lev_hcn = (4, 3) if small else (12, 6)
    e_ch4 = syn.S.CH4_LEVEL_ENERGIES[:n_lev_ch4]
    e_hcn = np.array([0.0, 712.0, 1412.0, 2097.0, 3311.5, 4004.0])[:n_lev_hcn]
    nlte_molecs = dict()
    nlte_molecs['CH4'] = syn.nlte_molec(6, 'CH4', atm, e_ch4, 7, sza_nodes=sza_nodes, lat_interp='lin')
    nlte_molecs['HCN'] = syn.nlte_molec(23, 'HCN', atm, e_hcn, 7, sza_nodes=sza_nodes, lat_interp='lin')
    atm_gases_old = dict(CH4=syn.vmr_profile(grid, atm, 0.015, 7, lat_interp='lin'),
                         HCN=syn.vmr_profile(grid, atm, 2.e-6, 7, lat_interp='lin'))
    for molec in nlte_molecs.values():
        molec.link_to_atmos(Atm)
        molec.add_clim(atm_gases_old[molec.name])
        planet.add_gas(molec)
    planet3D = planet

    ##### SETTING THE BAYESSET (1D):
    zold = atm['z']
    alt_gri = sbm.AtmGrid('alt', zold)
    vmr_1D = dict(CH4=sbm.AtmProfile(alt_gri, np.full(len(zold), 0.015), profname='vmr', interp='lin'),
                  HCN=sbm.AtmProfile(alt_gri, np.full(len(zold), 2.e-6), profname='vmr', interp='lin'))
    baybau1D = smm.BayesSet(tag='test_CH4_HCN_1D')
    alt_nodes = np.arange(450., 1051., 100.)
    cososo = vmr_1D['CH4']
    prf = []
    for alt in alt_nodes:
        prf.append(cososo.calc(alt))
    apriori_prof = np.array(prf)
    apriori_prof_err = apriori_prof + 0.015
    set_ = smm.LinearProfile_1D_new('CH4', alt_gri, alt_nodes, apriori_prof, apriori_prof_err)
    baybau1D.add_set(set_)
    if not small:
        alt_nodes = np.arange(550., 1051., 100.)
        cososo = vmr_1D['HCN']
        apriori_prof = np.array([cososo.calc(alt) for alt in alt_nodes])
        set_ = smm.LinearProfile_1D_new('HCN', alt_gri, alt_nodes, apriori_prof, apriori_prof + 3.e-4)
        baybau1D.add_set(set_)

    ###############################################################
    wn_range = [2996., 3004.] if small else [2850., 3450.]
    wn_range_obs = [spcl.convertto_nm(wn_range[1], 'cm_1') + 10., spcl.convertto_nm(wn_range[0], 'cm_1') - 10.]
    print(wn_range_obs)

    radtran_opt = dict()
    radtran_opt['max_T_variation'] = 5.
    radtran_opt['max_Plog_variation'] = 1.

    print('Loading lines...')
    n_ch4, n_hcn = (160, 40) if small else (30000, 2000)
    db_file = syn.write_hitran_file(os.path.join(base, 'synthetic_hitran.par'), wn_range, [
        dict(mol=6, iso=1, n_lines=n_ch4, level_energies=e_ch4, q296=590.52, ratio=nlte_molecs['CH4'].iso_1.ratio),
        dict(mol=23, iso=1, n_lines=n_hcn, level_energies=e_hcn, q296=892.2, ratio=nlte_molecs['HCN'].iso_1.ratio)])
    linee = spcl.read_line_database(db_file, freq_range=wn_range)
    linee = smm.check_lines_mols(linee, planet3D.gases.values())
    smm.keep_levels_wlines(planet3D, linee)

    LUTopt = dict()
    LUTopt['max_pres'] = 0.1   # hPa circa 200 km
    LUTopt['temp_step'] = 5.
    LUTopt['pres_step_log'] = 1.0

    sp_gri = smm.prepare_spe_grid(wn_range).spectral_grid
    PTcoup_needed = smm.calc_PT_couples_atmosphere(linee, planet3D.gases.values(), planet3D.atmosphere, **LUTopt)
    LUTS = smm.check_and_build_allluts(inputs, sp_gri, linee, planet3D.gases.values(), PTcouples=PTcoup_needed, LUTopt=LUTopt)
    print('{} PT couples, LUTs ready after {:6.2f} s'.format(len(PTcoup_needed), time.time() - time0))

    ###################################################################
    tangents = [520., 700., 880.] if small else list(np.arange(460., 1021., 40.))
    pixels = syn.observed_pixels(tangents, wn_range, 12 if small else 36, lat=12.0,
                                 sza=np.linspace(55., 75., len(tangents)))
    for pix in pixels:
        pix.pixel_rot = 0.0
    pixels.sort(key=lambda x: x.limb_tg_alt)

    # "observations": the 3-D forward model with the a-priori CH4 scaled by 1.25
    truth = copy.deepcopy(baybau1D)
    for par in truth.sets['CH4'].set:
        par.value = 1.25 * par.apriori
    planet_true = copy.deepcopy(planet3D)
    for gas in truth.sets.keys():
        planet_true.gases[gas].add_clim(truth.sets[gas].profile())
    sims_true, _, _ = smm.radtrans(inputs, planet_true, linee, copy.deepcopy(pixels), wn_range=wn_range,
                                   radtran_opt=radtran_opt, LUTopt=LUTopt, save_hires=False,
                                   group_observations=True, nome_inv='truth')
    syn.set_observations(pixels, sims_true)

    track_levels_all = dict()
    for molnam in ['CH4', 'HCN']:
        mol = planet3D.gases[molnam]
        for iso in mol.all_iso:
            isomol = getattr(mol, iso)
            track_levels_all[(molnam, iso)] = isomol.levels

    track_levels_short = dict()
    track_levels_short[('CH4', 'iso_1')] = planet3D.gases['CH4'].iso_1.levels[1:3]
    track_levels_short[('HCN', 'iso_1')] = planet3D.gases['HCN'].iso_1.levels[1:2]

    # the a-priori VMRs of the parameter space are what the radtrans runs see
    for gas in baybau1D.sets.keys():
        planet3D.gases[gas].add_clim(baybau1D.sets[gas].profile())

    results = dict()
    for teag, kw in (('tracklevels_szavar_all', dict(use_tangent_sza=False, track_levels=track_levels_all)),
                     ('tracklevels_noszavar_short', dict(use_tangent_sza=True, track_levels=track_levels_short)),
                     ('3D_tracklevels_inverseLOS_short', dict(use_tangent_sza=False, invert_LOS_direction=True,
                                                              track_levels=track_levels_short))):
        dampa = open(inputs['out_dir'] + './radtran_' + teag + '.pic', 'wb')
        result = smm.radtrans(inputs, planet3D, linee, pixels, wn_range=wn_range, radtran_opt=radtran_opt,
                              LUTopt=LUTopt, nome_inv=teag, save_hires=True, group_observations=True, **kw)
        pickle.dump(result, dampa)
        dampa.close()
        results['radtran_' + teag] = result
        print('Tempo totale: {} min'.format((time.time() - time0) / 60.))

    print('Faccio le inversions')
    for teag, kw in (('2Dvs3D_szavar_lin', dict()),
                     ('2Dvs3D_noszavar_lin', dict(use_tangent_sza=True)),
                     ('2Dvs3D_inverseLOS_lin', dict(use_tangent_sza=False, invert_LOS_direction=True))):
        bay = copy.deepcopy(baybau1D)
        dampa = open(inputs['out_dir'] + './out_' + teag + '.pic', 'wb')
        result = smm.inversion_fast_limb(inputs, copy.deepcopy(planet3D), linee, bay, pixels, wn_range=wn_range,
                                         radtran_opt=radtran_opt, debugfile=dampa, LUTopt=LUTopt, nome_inv=teag,
                                         group_observations=True, max_it=2 if small else 10, **kw)
        dampa.close()
        results['out_' + teag] = result
        print('Tempo totale: {} min'.format((time.time() - time0) / 60.))

    print(time.ctime())
    return results, truth, sims_true, planet3D, linee, pixels


if __name__ == '__main__':
    main()
