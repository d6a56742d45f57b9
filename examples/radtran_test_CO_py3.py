#!/usr/bin/python
# -*- coding: utf-8 -*-
"""Python-3 copy of the reference driver radtran_test_CO.py (BASELINE.json configs[0]): CO 1-0 band
(~4.7 um) line-by-line cross-sections and single-LOS radiances through a 1-D Titan profile with a
non-LTE CO, a 1-D linear VMR parameter space and the per-LOS retrieval smm.inversion(...)
(radtran_test_CO.py:194) -> LineOfSight.radtran -> hires_to_lowres -> FOV_integr_1D.
What differs from the original: Python 3, no absolute paths, synthetic inputs
(examples/synthetic_inputs.py) for the un-shipped profile / T_vib / HITRAN / VIMS files.
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


def main(small=bool(int(os.environ.get('SR_EXAMPLE_SMALL', '0'))), useLUTs=True):
    time0 = time.time()
    base, cart_LUTS, out_dir = syn.work_dirs('radtran_test_CO')
    inputs = dict(cart_LUTS=cart_LUTS, out_dir=out_dir, n_threads=8, test=False)

    wn_range = [2140., 2150.] if small else [2050., 2250.]
    wn_range_obs = [spcl.convertto_nm(wn_range[1], 'cm_1') + 10., spcl.convertto_nm(wn_range[0], 'cm_1') - 10.]
    print(wn_range_obs)

    ### LOADING PLANET
    print('Loading planet...')
    planet = sbm.Titan(1000.)
    alt_gri, atm_old, atm = syn.atmosphere(n_bands=1, z_top=1000.0)
    planet.add_atmosphere(atm_old)

    ### LOADING MOLECULES
    print('Loading molecules...')
    e_co = np.array([0.0, 2143.27, 4260.06])                  # CO v = 0, 1, 2
    co = syn.nlte_molec(5, 'CO', atm, e_co, 1)
    co.link_to_atmos(atm_old)
    co.add_clim(syn.vmr_profile(alt_gri, atm, 50.e-6, 1))
    planet.add_gas(co)

    ### LOADING LINES
    print('Loading lines...')
    db_file = syn.write_hitran_file(os.path.join(base, 'synthetic_hitran.par'), wn_range, [
        dict(mol=5, iso=1, n_lines=40 if small else 400, level_energies=e_co, q296=107.42, ratio=co.iso_1.ratio)])
    linee = spcl.read_line_database(db_file, freq_range=wn_range)
    planetmols = [gas.mol for gas in planet.gases.values()]
    linee = [lin for lin in linee if lin.Freq >= wn_range[0] and lin.Freq <= wn_range[1] and lin.Mol in planetmols]
    print(len(linee))
    print(planet.gases)

    ##### SETTING THE BAYESSET:
    baybau = smm.BayesSet(tag='test_CO_vero')
    alt_nodes = np.arange(200., 501., 50.)
    apriori_prof = np.ones(7) * 50.0 * 1.e-6
    apriori_prof_err = 0.7 * apriori_prof
    set_ = smm.LinearProfile_1D('CO', planet.atmosphere, alt_nodes, apriori_prof, apriori_prof_err)
    baybau.add_set(set_)

    ### updating the profile of gases in bayesset
    for gas in baybau.sets.keys():
        planet.gases[gas].add_clim(baybau.sets[gas].profile())

    pixels = syn.observed_pixels([250., 400.] if small else list(np.arange(220., 481., 20.)), wn_range,
                                 10 if small else 40, lat=-25.0, sza=30.0)
    pixels = pixels[::1 if small else 5]

    radtran_opt = dict()
    radtran_opt['max_T_variation'] = 5.
    radtran_opt['max_Plog_variation'] = 1.0

    LUTopt = dict()
    LUTopt['temp_step'] = 5.
    LUTopt['pres_step_log'] = 1.0
    LUTopt['max_pres'] = 2.5

    # "observations": the forward model with 1.3 x the a-priori CO
    truth = copy.deepcopy(baybau)
    for par in truth.sets['CO'].set:
        par.value = 1.3 * par.apriori
    pl_true = copy.deepcopy(planet)
    pl_true.gases['CO'].add_clim(truth.sets['CO'].profile())
    sims_true, _, _ = smm.radtrans(inputs, pl_true, linee, copy.deepcopy(pixels), wn_range=wn_range, radtran_opt=radtran_opt,
                                   LUTopt=LUTopt, save_hires=False, nome_inv='truth')
    syn.set_observations(pixels, sims_true)

    dampa = open(os.path.join(out_dir, 'debuh_yeah.pic'), 'wb')
    result = smm.inversion(inputs, planet, linee, baybau, pixels, wn_range=wn_range, radtran_opt=radtran_opt, debugfile=dampa, useLUTs=useLUTs, test=inputs['test'], LUTopt=LUTopt, max_it=3 if small else 10)
    dampa.close()

    tot_time = time.time() - time0
    print('Tempo totale: {} min'.format(tot_time / 60.))
    print('Tempo una LOS: {} min'.format(tot_time / (3. * len(pixels)) / 60.))
    return result, truth, sims_true, planet, linee, pixels


if __name__ == '__main__':
    main()
