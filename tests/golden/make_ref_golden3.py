"""Third set of reference-executed fixtures (tests/golden/ref_golden3.npz): the parameter space
and the update step of the retrieval that smm.inversion_fast_limb runs around the GPU forward
model - LinearProfile_1D_new / RetParam / BayesSet (smm:163-257, 450-656), genvec, chicalc and the
Levenberg-Marquardt step inversion_algebra (smm:3399-3469) - executed from the reference's own
Python by ref_exec.py.  Run from the repo root in a container that has /root/reference:

    python tests/golden/make_ref_golden3.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import ref_exec as R  # noqa: E402


def case(spcl, smm, sbm):
    """Two sets (7 + 5 altitude nodes), 4 'pixels' of 9 channels with synthetic derivative spectra,
    masks that drop some channels, one parameter whose LM step would make it negative."""
    rng = np.random.default_rng(20067)
    z = np.arange(0.0, 1501.0, 10.0)
    alt_gri = sbm.AtmGrid('alt', z)
    bs = smm.BayesSet(tag='fixture')
    n1, n2 = np.arange(450., 1051., 100.), np.arange(550., 1051., 125.)
    bs.add_set(smm.LinearProfile_1D_new('CH4', alt_gri, n1, 0.015 + 1e-5 * n1, 0.5 * (0.015 + 1e-5 * n1)))
    bs.add_set(smm.LinearProfile_1D_new('HCN', alt_gri, n2, 2.e-6 + 0 * n2, 3.e-4 + 0 * n2,
                                        first_guess_prof=1.5e-6 + 0 * n2))
    n_pix, n_ch = 4, 9
    grid = spcl.SpectralGrid(np.linspace(3280., 3320., n_ch), units='nm')
    obs, sims, noise, masks = [], [], [], []
    pars = bs.params()
    for k in range(n_pix):
        sim = 1e-7 * (1 + rng.uniform(0, 1, n_ch))
        sims.append(spcl.SpectralIntensity(sim, grid, units='Wm2'))
        obs.append(spcl.SpectralIntensity(sim * (1 + 0.2 * rng.normal(size=n_ch)), grid, units='Wm2'))
        noise.append(spcl.SpectralObject(np.full(n_ch, 1.5e-8), grid))
        m = np.ones(n_ch)
        m[rng.integers(0, n_ch, 2)] = 0
        masks.append(m)
        for q, par in enumerate(pars):
            scale = 1e-7 / par.apriori * np.exp(-0.5 * ((q % 7) - 1.5 * k) ** 2)
            der = scale * rng.uniform(0.2, 1.0, n_ch)
            if par.nameset == 'HCN' and q == len(pars) - 1:
                der = -40 * np.abs(der)                        # drives this parameter below zero
            par.store_deriv(spcl.SpectralIntensity(der, grid, units='Wm2'), num=k)
    return bs, obs, sims, noise, masks, z


def main():
    spcl, smm, sbm = R.load()
    out = dict()
    bs, obs, sims, noise, masks, z = case(spcl, smm, sbm)
    pars = bs.params()
    out['z'] = z
    out['keys'] = np.array([par.key for par in pars])
    out['namesets'] = np.array([par.nameset for par in pars])
    out['masks_par'] = np.array([par.maskgrid.mask for par in pars if par.nameset == 'CH4'])
    out['apriori'] = bs.apriori_vector()
    out['values0'] = bs.param_vector()
    out['vcm_apriori'] = bs.VCM_apriori()
    out['obs'] = np.array([o.spectrum for o in obs])
    out['sims'] = np.array([s.spectrum for s in sims])
    out['noise'] = np.array([n.spectrum for n in noise])
    out['masks'] = np.array(masks)
    out['derivs'] = np.array([[d.spectrum for d in par.derivatives] for par in pars])   # [par][pix][ch]
    out['jac_nomask'] = np.array(bs.build_jacobian())
    out['jac_mask'] = np.array(bs.build_jacobian(masks=masks))
    ov, sv, nv = smm.genvec(obs, sims, noise, masks=masks)
    out['genvec'] = np.array([ov, sv, nv])
    out['chi'] = np.array([smm.chicalc(obs, sims, noise, masks, 5), smm.chicalc(obs, sims, noise, None, 0)])
    for tag, lam in (('a', 0.1), ('b', 3.0)):
        smm.inversion_algebra(obs, sims, noise, bs, lambda_LM=lam, masks=masks)
        out['values_' + tag] = bs.param_vector()
        out['avk_' + tag] = np.array(bs.av_kernel)
        out['vcm_' + tag] = np.array(bs.VCM)
    out['old_params'] = np.array(bs.old_params)
    out['old_values_last'] = np.array(pars[-1].old_values)
    # ---- involvement rule of a parameter (LinearProfile_1D_new.check_involved, smm:492-501) --------
    ch4 = bs.sets['CH4']
    ranges = [[0., 400.], [500., 700.], [760., 900.], [1060., 1400.]]
    out['involved_ranges'] = np.array(ranges)
    out['involved'] = np.array([[ch4.check_involved(key, {'alt': r}) for key in ch4.alts] for r in ranges])

    # ---- the ladder spline (make_radtran_spline, smm:3377-3396) -----------------------------------
    rng = np.random.default_rng(5)
    alts = np.arange(400., 901., 50.)
    grid = spcl.SpectralGrid(np.linspace(3280., 3320., 9), units='nm')
    ladder = [spcl.SpectralIntensity(1e-7 * np.exp(-a / 300.) * (1 + 0.3 * rng.uniform(size=9)), grid, units='Wm2')
              for a in alts]
    spl = smm.make_radtran_spline(alts, ladder)
    probes = np.array([400., 437.5, 612.0, 650., 899.9])
    out['spline_alts'], out['spline_probes'] = alts, probes
    out['spline_in'] = np.array([l.spectrum for l in ladder])
    out['spline_out'] = np.array([spl(x).spectrum for x in probes])

    # ---- line / level filters (smm:69-160) on the fixture's line file -----------------------------
    import types
    import make_ref_golden as M
    lines = spcl.read_line_database(os.path.join(HERE, 'ref_lines.par'))
    dec = (lambda v: v.decode() if isinstance(v, bytes) else v)
    for l in lines:
        for k in ('Up_lev_str', 'Lo_lev_str'):
            setattr(l, k, dec(getattr(l, k)))
        l.Mol, l.Iso = int(l.Mol), int(l.Iso)

    def planet():
        ch4 = sbm.Molec(6, 'CH4')
        ch4.add_iso(1).add_levels(M.LEVELS + ['1 0 0 0 1A1', '0 0 0 1 1F2'], M.ENERGIES + [2916.5, 1310.8])
        ch4.add_iso(2)
        hcn = sbm.Molec(23, 'HCN')
        hcn.add_iso(1)
        return types.SimpleNamespace(gases={'CH4': ch4, 'HCN': hcn})

    pl = planet()
    with R.quiet():
        ok = smm.check_lines_mols(lines, [pl.gases['CH4'], pl.gases['HCN']])
    out['filter_lines_ok'] = np.array([l.Freq for l in ok])
    out['track_all'] = np.array(sorted('%s/%s/%s' % (g, i, lev) for (g, i), levs in smm.track_all_levels(pl).items()
                                       for lev in levs))
    with R.quiet():
        smm.keep_levels_wlines(pl, lines)
    out['levels_wlines'] = np.array(pl.gases['CH4'].iso_1.levels)
    with R.quiet():
        smm.keep_levels(pl, {('CH4', 'iso_1'): ['lev_00', 'lev_02'], ('CH4', 'iso_2'): [], ('HCN', 'iso_1'): []})
    out['levels_kept'] = np.array(pl.gases['CH4'].iso_1.levels)
    with R.quiet():
        ok = smm.check_lines_mols(lines, [pl.gases['CH4']])
    out['filter_lines_ok2'] = np.array([l.Freq for l in ok])

    np.savez_compressed(os.path.join(HERE, 'ref_golden3.npz'), **out)
    print('wrote', os.path.join(HERE, 'ref_golden3.npz'), 'with', len(out), 'arrays')


if __name__ == '__main__':
    main()
