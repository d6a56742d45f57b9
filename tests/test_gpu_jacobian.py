"""GPU: analytic LOS Jacobians (k_los_layers_jac behind sr_los_rt_layers_jac_dev /
sr_los_rt_lut_jac_dev / sr_los_rt_lut_jac_lowres_dev) against the
CPU oracle's closed-sum form and against central finite differences of the CUDA forward model."""
import numpy as np
import pytest

from conftest import rel_err
from test_gpu_los import case  # noqa: F401  (module-scoped fixture: small non-LTE LUT + 6 limb LOS)
from test_jacobian_oracle import _layers

pytestmark = pytest.mark.gpu

TOL_JAC = 1e-5      # of the largest |derivative| of the spectrum (north_star's radiance tolerance)


def _jac_err(got, ref):
    scale = np.abs(ref).max(axis=-1, keepdims=True)
    scale = np.where(scale == 0.0, 1.0, scale)
    return float((np.abs(got - ref) / scale).max())


@pytest.mark.parametrize("n_par", [3, 7, 16, 21])
def test_layers_jac_kernel_vs_oracle(oracle, n_par):
    import torch
    from spectrobot_b200 import engine
    rng = np.random.default_rng(11 + n_par)
    n_steps, tau_g, emi_g, tau_o, emi_o, dfrac, i0 = _layers(rng, n_pts=777, n_par=n_par)
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), device="cuda")   # noqa: E731
    nst = torch.tensor(n_steps, dtype=torch.int32, device="cuda")
    for solo in (False, True):
        # the retrieved gas is the only absorber
        ref_r, ref_j = oracle.los_layers_jac(tau_g, emi_g, dfrac, n_steps, i0=i0, solo_absorption=solo)
        rad, jac = engine.los_rt_layers_jac(dev(tau_g), dev(emi_g), dev(dfrac), nst, i0=dev(i0),
                                            solo_absorption=solo)
        assert rel_err(rad.cpu().numpy(), ref_r) < 1e-12
        assert _jac_err(jac.cpu().numpy(), ref_j) < 1e-11
        # with a second absorber
        ref_r, ref_j = oracle.los_layers_jac(tau_g + tau_o, emi_g + emi_o, dfrac, n_steps,
                                             tau_g=tau_g, emi_g=emi_g, i0=i0, solo_absorption=solo)
        rad, jac = engine.los_rt_layers_jac(dev(tau_g + tau_o), dev(emi_g + emi_o), dev(dfrac), nst,
                                            tau_g=dev(tau_g), emi_g=dev(emi_g), i0=dev(i0),
                                            solo_absorption=solo)
        assert rel_err(rad.cpu().numpy(), ref_r) < 1e-12
        assert _jac_err(jac.cpu().numpy(), ref_j) < 1e-11


def _dfrac(st, n_par, seed=3):
    """Synthetic 'triangle' masks: every step depends on two neighbouring parameters."""
    rng = np.random.default_rng(seed)
    n_los, nmax = st["temp"].shape
    d = np.zeros((n_los, nmax, n_par))
    for l in range(n_los):
        for k in range(int(st["n_steps"][l])):
            p = (k * n_par) // max(int(st["n_steps"][l]), 1)
            w = rng.uniform(0.2, 0.8)
            d[l, k, p] = w
            d[l, k, min(p + 1, n_par - 1)] += 1.0 - w
    return d


def test_fused_jac_single_gas(case, oracle, monkeypatch):   # noqa: F811
    eng, st = case["engine"], case["st"]
    n_par = 5
    dfrac = _dfrac(st, n_par)
    rad, jac = eng.los_rt_lut_jac([case["lut"]], case["steps"], dfrac)
    base = eng.los_rt_lut([case["lut"]], case["steps"])
    assert rel_err(rad.cpu().numpy(), base.cpu().numpy()) < 1e-13
    # oracle on the layers the device assembled (abs x column, emi x column)
    tau, emi = eng.los_abs_emi([case["lut"]], case["steps"])
    _, ref_j = oracle.los_layers_jac(tau.cpu().numpy(), emi.cpu().numpy(), dfrac, st["n_steps"])
    jac = jac.cpu().numpy()
    assert _jac_err(jac, ref_j) < 1e-10
    # central finite differences of the fused forward model: columns scaled by (1 +- h f_p)
    h = 1e-5
    scale = np.abs(base.cpu().numpy()).max()
    for p in (0, 3):
        f = dfrac[:, :, p]
        up = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"] * (1 + h * f), st["tvib"])
        dn = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"] * (1 - h * f), st["tvib"])
        fd = (eng.los_rt_lut([case["lut"]], up) - eng.los_rt_lut([case["lut"]], dn)).cpu().numpy() / (2 * h)
        assert np.abs(jac[:, p] - fd).max() < 1e-6 * max(np.abs(fd).max(), scale)
    # LOS blocks / wavenumber chunks of the scratch do not change a bit; a sub-window neither
    monkeypatch.setenv("SR_LOS_BLOCK", "4")
    monkeypatch.setenv("SR_LOS_CHUNK", "768")
    rad_b, jac_b = eng.los_rt_lut_jac([case["lut"]], case["steps"], dfrac)
    assert np.array_equal(rad_b.cpu().numpy(), rad.cpu().numpy())
    assert np.array_equal(jac_b.cpu().numpy(), jac)
    monkeypatch.delenv("SR_LOS_BLOCK")
    monkeypatch.delenv("SR_LOS_CHUNK")
    rad_s, jac_s = eng.los_rt_lut_jac([case["lut"]], case["steps"], dfrac, pt0=1001, n_pts=777)
    assert np.array_equal(jac_s.cpu().numpy(), jac[:, :, 1001:1778])


def test_fused_jac_two_gases_and_lowres(case, oracle, monkeypatch):   # noqa: F811
    """Retrieved gas + a second absorber (its columns do not depend on the parameters), and the
    low-resolution variant: derivative spectra convolved to the channels on the device."""
    eng, S, st, torch = case["engine"], case["S"], case["st"], case["torch"]
    g = case["grid"]
    lines2 = S.line_table(120, 2996.0, 3005.0, n_levels=1, seed=8, q296=107.12, iso_ratio=0.986544)
    ls2 = eng.LineSet(lines2, g, 27.994915, 1)
    g32b = ls2.gcoeff_cells_f32(case["cells"])
    lut2 = eng.Lut(g32b, case["cells"], 5, 1, 0.986544, level_energies=None)
    col = np.concatenate([st["column"], 0.3 * st["column"]], axis=0)
    steps = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], col, None)
    luts = [case["lut"], lut2]
    n_par = 9
    dfrac = _dfrac(st, n_par, seed=5)
    rad, jac = eng.los_rt_lut_jac(luts, steps, dfrac, gas_in_jac=[1, 0])
    jac = jac.cpu().numpy()
    base = eng.los_rt_lut(luts, steps).cpu().numpy()
    assert rel_err(rad.cpu().numpy(), base) < 1e-13
    tau, emi = eng.los_abs_emi(luts, steps)
    only = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], np.concatenate([st["column"], 0 * st["column"]]), None)
    tau_g, emi_g = eng.los_abs_emi(luts, only)
    _, ref_j = oracle.los_layers_jac(tau.cpu().numpy(), emi.cpu().numpy(), dfrac, st["n_steps"],
                                     tau_g=tau_g.cpu().numpy(), emi_g=emi_g.cpu().numpy())
    assert _jac_err(jac, ref_j) < 1e-10
    h = 1e-5
    p = 4
    f = dfrac[:, :, p]
    colp = np.concatenate([st["column"] * (1 + h * f), 0.3 * st["column"]], axis=0)
    colm = np.concatenate([st["column"] * (1 - h * f), 0.3 * st["column"]], axis=0)
    fd = (eng.los_rt_lut(luts, eng.LosSteps(st["n_steps"], st["temp"], st["pres"], colp, None)) -
          eng.los_rt_lut(luts, eng.LosSteps(st["n_steps"], st["temp"], st["pres"], colm, None))).cpu().numpy() / (2 * h)
    assert np.abs(jac[:, p] - fd).max() < 1e-6 * max(np.abs(fd).max(), np.abs(base).max())
    # derivative w.r.t. the OTHER gas' column is a different thing: sanity, the two differ
    _, jac_all = eng.los_rt_lut_jac(luts, steps, dfrac)
    assert not np.allclose(jac_all.cpu().numpy(), jac)

    gdev = torch.as_tensor(g, device="cuda")
    centres = np.linspace(g[0] + 0.3, g[-1] - 0.3, 7)
    widths = np.full(7, 0.08)
    low, jlow = eng.los_rt_lut_jac_lowres(luts, steps, dfrac, gdev, centres, widths, gas_in_jac=[1, 0])
    ref_low = eng.convolve_lowres(gdev, rad, centres, widths).cpu().numpy()
    jd = torch.as_tensor(jac, device="cuda").reshape(-1, len(g))
    ref_jlow = eng.convolve_lowres(gdev, jd, centres, widths).cpu().numpy().reshape(len(base), n_par, 7)
    assert np.array_equal(low.cpu().numpy(), ref_low)
    assert np.array_equal(jlow.cpu().numpy(), ref_jlow)
    monkeypatch.setenv("SR_LOS_BLOCK", "4")
    monkeypatch.setenv("SR_LOS_CHUNK", "1024")
    low_b, jlow_b = eng.los_rt_lut_jac_lowres(luts, steps, dfrac, gdev, centres, widths, gas_in_jac=[1, 0])
    assert np.array_equal(low_b.cpu().numpy(), ref_low)
    assert np.array_equal(jlow_b.cpu().numpy(), ref_jlow)
