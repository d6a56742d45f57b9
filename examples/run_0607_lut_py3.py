#!/usr/bin/python
# -*- coding: utf-8 -*-
"""Python-3 copy of the reference's LUT drivers run_0607_lut.py / run_0607_lut_HCN.py
(BASELINE.json configs[2], configs[3]) on synthetic inputs: per gas, the (P,T) cells the atmosphere
needs and the LUT of every isotopologue, built by smm.check_and_build_allluts
(run_0607_lut_HCN.py:104-113).  Only names of the reference API are used from "LOADING PLANET" on;
what differs from the original file: Python 3, no absolute paths, synthetic inputs instead of the
un-shipped HITRAN / climatology files (examples/synthetic_inputs.py).
SR_EXAMPLE_SMALL=1 shrinks the spectral ranges and line counts (used by the tests)."""
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
    base, cart_LUTS, out_dir = syn.work_dirs('run_0607_lut')
    inputs = dict(cart_LUTS=cart_LUTS, out_dir=out_dir, n_threads=8, test=False)

    ### LOADING PLANET
    print('Loading planet...')
    planet = sbm.Titan(1500.)
    grid, Atm, atm = syn.atmosphere(n_bands=7)
    planet.add_atmosphere(Atm)

    ### LOADING MOLECULES
    print('Loading molecules...')
    ch4 = syn.nlte_molec(6, 'CH4', atm, syn.S.CH4_LEVEL_ENERGIES[:4 if small else 12], 7)
    ch4.add_iso(2, LTE=True)
    ch4.link_to_atmos(Atm)
    ch4.add_clim(syn.vmr_profile(grid, atm, 0.015, 7))
    hcn = syn.nlte_molec(23, 'HCN', atm, np.array([0.0, 712.0, 1412.0, 2097.0, 3311.5, 4004.0][:3 if small else 6]), 7)
    hcn.link_to_atmos(Atm)
    hcn.add_clim(syn.vmr_profile(grid, atm, 1.e-6, 7))
    planet.add_gas(ch4)
    planet.add_gas(hcn)

    LUTopt = dict()
    LUTopt['max_pres'] = 0.1   # hPa circa 120 km
    LUTopt['temp_step'] = 5.
    LUTopt['pres_step_log'] = 1.0

    wn_ranges = dict()
    wn_ranges['HCN'] = [3296., 3304.] if small else [3200., 3400.]
    wn_ranges['CH4'] = [2996., 3004.] if small else [2825., 3225.]

    ### LOADING LINES
    print('Loading lines...')
    n_ch4, n_hcn = (150, 60) if small else (30000, 3000)
    db_file = syn.write_hitran_file(os.path.join(base, 'synthetic_hitran.par'), [2800., 3450.] if not small else [2990., 3310.], [
        dict(mol=6, iso=1, n_lines=n_ch4, level_energies=ch4.iso_1.level_energies(), q296=590.52, ratio=ch4.iso_1.ratio),
        dict(mol=6, iso=2, n_lines=n_ch4 // 10, level_energies=None, q296=1180.8, ratio=ch4.iso_2.ratio),
        dict(mol=23, iso=1, n_lines=n_hcn, level_energies=hcn.iso_1.level_energies(), q296=892.2, ratio=hcn.iso_1.ratio)])
    linee = spcl.read_line_database(db_file, freq_range=wn_ranges['CH4'])
    linee += spcl.read_line_database(db_file, freq_range=wn_ranges['HCN'], mol=23)
    linee = smm.check_lines_mols(linee, planet.gases.values())

    allLUTS = dict()
    print(planet.gases.keys())
    for gas in planet.gases:
        print(gas)
        linee_ok = smm.check_lines_mols(linee, [planet.gases[gas]])
        linee_ok = [lin for lin in linee_ok if lin.Freq >= wn_ranges[gas][0] and lin.Freq <= wn_ranges[gas][1]]
        if len(linee_ok) == 0:
            continue
        abs_coeff = smm.prepare_spe_grid(wn_ranges[gas])
        sp_grid = abs_coeff.spectral_grid
        PTcouples = smm.calc_PT_couples_atmosphere(linee_ok, [planet.gases[gas]], planet.atmosphere, **LUTopt)
        print('{} PT couples for {}, {} lines'.format(len(PTcouples), gas, len(linee_ok)))
        t1 = time.time()
        LUTS = smm.check_and_build_allluts(inputs, sp_grid, linee_ok, [planet.gases[gas]], atmosphere=planet.atmosphere, LUTopt=LUTopt)
        print('LUTs of {} built in {:6.2f} s'.format(gas, time.time() - t1))
        allLUTS.update(LUTS)

    print(time.ctime())
    print('Tempo totale: {:6.2f} s'.format(time.time() - time0))
    return planet, linee, allLUTS, wn_ranges, LUTopt


if __name__ == '__main__':
    main()
