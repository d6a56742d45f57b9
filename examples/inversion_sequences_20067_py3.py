#!/usr/bin/python
# -*- coding: utf-8 -*-
"""Python-3 copy of the reference driver inversion_sequences_20067.py (BASELINE.json configs[4]):
a loop over limb-scan sequences (per year and 10-degree latitude bin), each retrieved with

    smm.inversion_fast_limb(inputs, planet3D, linee, bay1, pixels, wn_range=..., radtran_opt=...,
                            debugfile=dampa, LUTopt=LUTopt, nome_inv=teag, group_observations=True,
                            alt_first_los=300., check_log=check_log)                  (:409)

on a 3-D Titan with latitude BANDS (`['box','lin']`, :112-117), a 1-D BayesSet of
smm.LinearProfile_1D_new sets, the pixel masks of mask_and_check_pixels (:22-50) and the
check_log / results_inversion_0607.pic book-keeping (:367-440).  What differs from the original:
Python 3, no absolute paths, synthetic inputs (examples/synthetic_inputs.py) for the un-shipped
climatology / T_vib / HITRAN files and for the pickled sequences of VIMS pixels, whose "observed"
spectra are simulated with a known CH4 profile.  SR_EXAMPLE_SMALL=1 shrinks the spectral range,
the line / pixel counts and the number of sequences (used by the tests)."""
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


def mask_and_check_pixels(pixels, min_alt=400.):
    """The reference's pixel preparation (:22-50): channels masked where the spectrum is
    unphysical, flat noise, band widths kept, pixel rotation zero, pixels below min_alt dropped."""
    for pix in pixels:
        spe = pix.observation.spectrum
        cond2 = (spe < -3.e-8) | (spe > 3.e-6)
        pix.observation.mask[cond2] = 0
        if len(spe) != len(pix.observation.noise.spectrum) or len(spe) != len(pix.observation.mask):
            raise ValueError('Inconsistent length of mask or noise')
        pix.pixel_rot = 0.0
    pixels = [pix for pix in pixels if pix.limb_tg_alt > min_alt]
    pixels.sort(key=lambda x: x.limb_tg_alt)
    return pixels


def main(small=bool(int(os.environ.get('SR_EXAMPLE_SMALL', '0')))):
    time0 = time.time()
    base, cart_LUTS, out_dir = syn.work_dirs('inversion_sequences')
    inputs = dict(cart_LUTS=cart_LUTS, out_dir=out_dir, n_threads=8, test=False, n_split=None)
    sbm.check_free_space(inputs['cart_LUTS'])
    sbm.check_free_space(inputs['out_dir'])

    ### LOADING PLANET
    print('Loading planet...')
    planet = sbm.Titan(1500.)
    grid, Atm, atm = syn.atmosphere(n_bands=7)
    planet.add_atmosphere(Atm)

    ### LOADING MOLECULES
    print('Loading molecules...')
    sza_nodes = syn.S.SZA_NODES
    e_ch4 = syn.S.CH4_LEVEL_ENERGIES[:4 if small else 12]
    nlte_molecs = dict()
    nlte_molecs['CH4'] = syn.nlte_molec(6, 'CH4', atm, e_ch4, 7, sza_nodes=sza_nodes)
    atm_gases_old = dict(CH4=syn.vmr_profile(grid, atm, 0.015, 7))
    for molec in nlte_molecs.values():
        molec.link_to_atmos(Atm)
        molec.add_clim(atm_gases_old[molec.name])
        planet.add_gas(molec)
    planet3D = planet

    ##### SETTING THE BAYESSET (1D):
    zold = atm['z']
    alt_gri = sbm.AtmGrid('alt', zold)
    baybau1D = smm.BayesSet(tag='test_CH4_1D')
    alt_nodes = np.arange(450., 1051., 100.)
    cososo = sbm.AtmProfile(alt_gri, np.full(len(zold), 0.015), profname='vmr', interp='lin')
    apriori_prof = np.array([cososo.calc(alt) for alt in alt_nodes])
    apriori_prof_err = apriori_prof + 0.015
    set_ = smm.LinearProfile_1D_new('CH4', alt_gri, alt_nodes, apriori_prof, apriori_prof_err)
    baybau1D.add_set(set_)

    wn_range = [2996., 3004.] if small else [2850., 3450.]
    radtran_opt = dict()
    radtran_opt['max_T_variation'] = 5.
    radtran_opt['max_Plog_variation'] = 1.

    print('Loading lines...')
    n_ch4 = 160 if small else 1000000
    db_file = syn.write_hitran_file(os.path.join(base, 'synthetic_hitran.par'), wn_range, [
        dict(mol=6, iso=1, n_lines=n_ch4, level_energies=e_ch4, q296=590.52, ratio=nlte_molecs['CH4'].iso_1.ratio)])
    linee = spcl.read_line_database(db_file, freq_range=wn_range)
    linee = smm.check_lines_mols(linee, planet3D.gases.values())
    smm.keep_levels_wlines(planet3D, linee)

    LUTopt = dict()
    LUTopt['max_pres'] = 0.2   # hPa: below the 300 km of alt_first_los in the synthetic atmosphere
    LUTopt['temp_step'] = 5.
    LUTopt['pres_step_log'] = 1.0

    # "observed" sequences: {(lat1, lat2): [sequence, ...]} per year, simulated with CH4 x 1.25
    truth = copy.deepcopy(baybau1D)
    for par in truth.sets['CH4'].set:
        par.value = 1.25 * par.apriori
    planet_true = copy.deepcopy(planet3D)
    planet_true.gases['CH4'].add_clim(truth.sets['CH4'].profile())
    years = [2006] if small else [2006, 2007]
    # (synthetic_inputs.observed_pixels looks from the equatorial plane: low latitudes keep the
    # nominal tangent altitudes within a few km of the true ones)
    bins = [(-20, -10), (10, 20)] if small else [(-20, -10), (-10, 0), (0, 10), (10, 20), (20, 30)]
    tangents = [430., 610., 790.] if small else list(np.arange(420., 1021., 50.))
    all_seqs_year = dict()
    for yea in years:
        all_seqs = dict()
        for k, (lat1, lat2) in enumerate(bins):
            sza = 40. + 7. * k + (yea - 2006)
            pixels = syn.observed_pixels(tangents, wn_range, 12 if small else 36, lat=0.5 * (lat1 + lat2),
                                         sza=np.linspace(sza, sza + 8., len(tangents)))
            sims_true, _, _ = smm.radtrans(inputs, planet_true, linee, copy.deepcopy(pixels), wn_range=wn_range,
                                           radtran_opt=radtran_opt, LUTopt=LUTopt, save_hires=False,
                                           group_observations=True, alt_first_los=300., nome_inv='truth')
            syn.set_observations(pixels, sims_true)
            all_seqs[(lat1, lat2)] = [dict(n_pixels=len(pixels), szas=[p.limb_tg_sza for p in pixels],
                                           pixels=pixels)]
        all_seqs_year[yea] = all_seqs

    sequences = []
    results_tot = []
    lats = np.arange(-90, 91, 10)
    num = 0
    check_log = open(inputs['out_dir'] + 'check_log_allinv.dat', 'a')
    check_log.write(time.ctime())
    for yea in years:
        print('YEAR ', yea)
        check_log.write('----------  YEAR {} ---------\n'.format(yea))
        all_seqs = all_seqs_year[yea]
        for lat1, lat2 in zip(lats[:-1], lats[1:]):
            latsss = (lat1, lat2)
            if latsss not in all_seqs:
                continue
            print('LATITUDE ---> ', latsss, len(all_seqs[latsss]))
            check_log.write('\n----------  LATITUDE {} - {} ---------\n'.format(lat1, lat2))
            sequa = all_seqs[latsss]
            if len(sequa) > 10:
                sequa.sort(key=lambda x: x['szas'][1])
                sequa = sequa[:10]
            for sequ1 in sequa:
                num += 1
                check_log.write('\nSEQ: n_pix {}, sza {:6.1f} \n'.format(sequ1['n_pixels'], np.mean(sequ1['szas'])))
                pixels = mask_and_check_pixels(sequ1['pixels'])
                sequ1['pixels'] = pixels
                teag = 'inversion_0607_seq_{:03d}'.format(num)
                bay1 = copy.deepcopy(baybau1D)
                time1 = time.time()
                dampa = open(inputs['out_dir'] + './out_' + teag + '.pic', 'wb')
                result = smm.inversion_fast_limb(inputs, planet3D, linee, bay1, pixels, wn_range=wn_range,
                                                 radtran_opt=radtran_opt, debugfile=dampa, LUTopt=LUTopt,
                                                 nome_inv=teag, group_observations=True, alt_first_los=300.,
                                                 check_log=check_log, max_it=2 if small else 10)
                dampa.close()
                results_tot.append(result[3])
                sequences.append(sequ1)
                tot_time = time.time() - time1
                print('Tempo totale: {} min'.format(tot_time / 60.))
                check_log.write('Tempo inversione: {:6.0f} min\n'.format(tot_time / 60.))

    dampa = open(inputs['out_dir'] + 'results_inversion_0607.pic', 'wb')
    pickle.dump([num, sequences, results_tot], dampa)
    dampa.close()
    print(time.ctime())
    print('Fine! {:6.1f} s'.format(time.time() - time0))
    check_log.write('\n' + time.ctime() + 'Fine!\n')
    check_log.close()
    return num, sequences, results_tot, truth, planet3D


if __name__ == '__main__':
    main()
