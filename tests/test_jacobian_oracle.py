"""CPU: the closed-sum Jacobian oracle (orc_los_layers_jac) against central finite
differences of the oracle's own layer recursion (orc_los_layers)."""
import numpy as np


def _layers(rng, n_los=3, n_steps_max=9, n_pts=40, n_par=5):
    n_steps = np.array([9, 6, 1], dtype=np.int32)[:n_los]
    tau_g = 10.0 ** rng.uniform(-9, 0.7, (n_los, n_steps_max, n_pts))
    emi_g = tau_g * rng.uniform(0.1, 3.0, tau_g.shape)
    tau_o = 10.0 ** rng.uniform(-6, 0.3, tau_g.shape)      # a second absorber
    emi_o = tau_o * rng.uniform(0.1, 3.0, tau_g.shape)
    tau_g[0, 2] = 0.0                                       # an empty layer
    emi_g[0, 2] = 0.0
    dfrac = rng.uniform(0.0, 1.0, (n_los, n_steps_max, n_par))
    dfrac[rng.uniform(size=dfrac.shape) < 0.5] = 0.0        # triangle masks: mostly zero
    i0 = rng.uniform(0.0, 2.0, (n_los, n_pts))
    return n_steps, tau_g, emi_g, tau_o, emi_o, dfrac, i0


def _forward(oracle, tau, emi, n_steps, i0, solo):
    src = np.where(tau > 0.0, emi / np.where(tau > 0.0, tau, 1.0), 0.0)
    return oracle.los_layers(tau, src, n_steps, i0=i0, solo_absorption=solo)


def _fd(oracle, tau_g, emi_g, tau_o, emi_o, dfrac, n_steps, i0, solo, p, h=1e-6):
    f = dfrac[:, :, p][:, :, None]
    up = _forward(oracle, tau_o + tau_g * (1 + h * f), emi_o + emi_g * (1 + h * f), n_steps, i0, solo)
    dn = _forward(oracle, tau_o + tau_g * (1 - h * f), emi_o + emi_g * (1 - h * f), n_steps, i0, solo)
    return (up - dn) / (2 * h)


def test_jacobian_oracle_matches_finite_differences(oracle):
    rng = np.random.default_rng(7)
    n_steps, tau_g, emi_g, tau_o, emi_o, dfrac, i0 = _layers(rng)
    zero = np.zeros_like(tau_g)
    for solo in (False, True):
        for multi in (False, True):
            to, eo = (tau_o, emi_o) if multi else (zero, zero)
            rad, jac = oracle.los_layers_jac(tau_g + to, emi_g + eo, dfrac, n_steps,
                                             tau_g=tau_g if multi else None,
                                             emi_g=emi_g if multi else None, i0=i0,
                                             solo_absorption=solo)
            ref = _forward(oracle, tau_g + to, emi_g + eo, n_steps, i0, solo)
            assert np.allclose(rad, ref, rtol=1e-12, atol=0)
            scale = np.abs(ref).max()
            for p in range(dfrac.shape[2]):
                fd = _fd(oracle, tau_g, emi_g, to, eo, dfrac, n_steps, i0, solo, p)
                assert np.abs(jac[:, p] - fd).max() < 2e-8 * scale, (solo, multi, p)
