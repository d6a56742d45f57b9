"""GPU test: a planet whose profiles are interpolated LINEARLY in latitude between band centres
(radtran_3Dvs2D_radtrans_new.py:82-111) is routed to the per-LOS host step builder (which itself
calls the curgods drop-in on the device) and gives the tables of the equivalent 1-D planet."""
import numpy as np
import pytest

from spectrobot_b200 import spect_base_module as sbm
from spectrobot_b200 import spect_main_module as smm
from spectrobot_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def test_latitude_linear_atmosphere_takes_the_host_step_builder():
    """A planet whose profiles are interpolated linearly in latitude
    (radtran_3Dvs2D_radtrans_new.py:82-111) is routed by los_step_tables_device to the per-LOS
    host methods; with identical rows the tables equal those of the 1-D planet, Jacobian
    fractions and tangent-SZA / photon-order options included."""
    energies = S.CH4_LEVEL_ENERGIES[:3]
    p1 = S.titan_planet(energies)                               # 1-D atmosphere, T_vib(z)
    a1 = p1.atmosphere
    lat_c = [-82.5, -45.0, 0.0, 45.0, 82.5]
    g2 = sbm.AtmGrid(['lat', 'alt'], [lat_c, a1.grid.coords['alt']])
    rep = lambda v: np.array([v] * len(lat_c))
    p2 = sbm.Titan(1500.0)
    atm2 = sbm.AtmProfile(g2, rep(a1.temp), 'temp', ['lin', 'lin'])
    atm2.add_profile(rep(a1.pres), 'pres', ['lin', 'exp'])
    p2.add_atmosphere(atm2)
    ch4 = sbm.Molec(6, 'CH4')
    im = ch4.add_iso(1, MM=S.CH4_MM, ratio=S.CH4_RATIO, LTE=False)
    ch4.add_clim(sbm.AtmProfile(g2, rep(p1.gases['CH4'].abundance.vmr), 'vmr', ['lin', 'lin']))
    im1 = p1.gases['CH4'].iso_1
    im.add_levels([getattr(im1, l).lev_string for l in im1.levels], energies,
                  vibtemps=[sbm.AtmProfile(g2, rep(getattr(im1, l).vibtemp.vibtemp), 'vibtemp', ['lin', 'lin'])
                            for l in im1.levels])
    p2.add_gas(ch4)
    assert smm.latitude_linear_profiles(p1) == []
    found = smm.latitude_linear_profiles(p2)
    assert 'atmosphere:temp' in found and 'CH4:vmr' in found and 'CH4/iso_1/lev_02:vibtemp' in found
    z = a1.grid.coords['alt']
    prof = smm.LinearProfile_1D_new('CH4', z, [400.0, 700.0, 1000.0], [0.015] * 3, [0.005] * 3)
    bs = smm.BayesSet('t')
    bs.add_set(prof)
    opts = dict(bayes_set=bs, set_name='CH4', delta_x=10.0, use_tangent_sza=True, LOS_order='photon')
    pix = S.vims_pixels([450.0, 800.0], lat=20.0)
    los2 = [p.LOS() for p in pix]
    gi2, st2, df2 = smm.los_step_tables_device(los2, p2, fszas=[50.0, 60.0], **opts)
    los1 = [p.LOS() for p in pix]
    gi1, st1, df1 = smm.los_step_tables_host(los1, p1, fszas=[50.0, 60.0], **opts)
    assert gi1 == gi2 == [('CH4', 'iso_1')]
    assert np.array_equal(st1.n_steps, st2.n_steps) and st1.n_steps.min() > 3
    for a, b in ((st1.temp, st2.temp), (st1.pres, st2.pres), (st1.column, st2.column),
                 (st1.tvib, st2.tvib), (df1, df2)):
        assert np.allclose(a, b, rtol=1e-12, atol=0.0)
    assert df2.shape == (2, st2.n_steps_max, 3) and np.any(df2 != 0.0)
    assert los2[0].LOS_order == 'photon' and np.all(los2[1].szas == 60.0)
    assert los2[0].involved_retparams[('CH4', prof.set[0].key)] in (True, False)
    # the device builder on the 1-D planet gives the same tables (steps at 1e-10)
    los3 = [p.LOS() for p in pix]
    gi3, st3, df3 = smm.los_step_tables_device(los3, p1, fszas=[50.0, 60.0], **opts)
    assert np.array_equal(st3.n_steps, st2.n_steps)
    n = st2.n_steps_max
    for a, b in ((st3.temp, st2.temp), (st3.pres, st2.pres), (st3.column, st2.column),
                 (st3.tvib, st2.tvib)):
        a, b = np.asarray(a)[..., :n], np.asarray(b)[..., :n]
        live = np.arange(n)[None, :] < st2.n_steps[:, None]
        assert np.allclose(np.where(live, a, 0.0), np.where(live, b, 0.0), rtol=1e-9, atol=0.0)
    assert np.allclose(np.asarray(df3)[:, :n], df2, rtol=1e-8, atol=1e-14)

