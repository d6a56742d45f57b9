"""GPU parity of the LOS path (K3a + K3) against the CPU oracle, through the C ABI.
Tolerance 1e-5 relative on radiances (north_star).  The LOS integral follows DESIGN.md section 6
(the reference's own implementation lives in the missing spect_base_module: parity unpinned)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_RAD = 1e-5


@pytest.fixture(scope="module")
def case():
    """Small non-LTE CH4-like LUT built with the CUDA K2 path + 6 synthetic limb LOS."""
    import torch
    from spectrobot_b200 import engine, synthetic as S
    g = S.spectral_grid(2999.0, 3002.0)
    n_lev = 6
    lines = S.line_table(250, 2996.0, 3005.0, n_levels=n_lev, seed=4)
    atm = S.titan_atmosphere()
    energies = lines["level_energies"]
    st = S.limb_los_steps([400.0, 520.0, 640.0, 760.0, 880.0, 1000.0], [3, 2, 4, 3, 1, 5],
                          [30., 40., 50., 60., 70., 80.], atm, energies)
    cells = S.rect_cells(st["pres"][st["pres"] > 1e-6].min() * 0.9, st["pres"].max() * 1.1,
                         st["temp"].min(), st["temp"].max())
    ls = engine.LineSet(lines, g, S.CH4_MM, n_lev)
    g32 = ls.gcoeff_cells_f32(cells)
    lut = engine.Lut(g32, cells, 6, 1, S.CH4_RATIO, level_energies=energies)
    steps = engine.LosSteps(st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    olut = dict(g32=g32.cpu().numpy(), pt=np.array(cells), level_energy=energies, mol=6, iso=1,
                iso_ratio=S.CH4_RATIO, lte_unidentified=False)
    return dict(engine=engine, S=S, grid=g, lut=lut, steps=steps, st=st, olut=olut, cells=cells,
                torch=torch)


def test_lut_weights_match_reference_rule(case, oracle):
    eng, cells = case["engine"], case["cells"]
    rng = np.random.default_rng(0)
    ps = np.unique([c[0] for c in cells]); ts = np.unique([c[1] for c in cells])
    for _ in range(200):
        P = np.exp(rng.uniform(np.log(ps[0]) - 1.0, np.log(ps[-1])))
        T = rng.uniform(ts[0], ts[-1])
        c_ref, w_ref = oracle.lut_weights(cells, P, T)
        c_got, w_got = eng.lut_weights(cells, P, T)
        assert np.array_equal(c_ref, c_got)
        assert np.allclose(w_ref, w_got, rtol=1e-14, atol=0)
    from spectrobot_b200._lib import SpectrobotError, SR_ERR_LUT
    with pytest.raises(SpectrobotError) as e:
        eng.lut_weights(cells, ps[-1] * 1.5, ts[3])          # smm:1058 "Extrapolating in P"
    assert e.value.code == SR_ERR_LUT


def test_fused_radiances_parity(case, oracle):
    eng, st = case["engine"], case["st"]
    ref = oracle.los_rt([case["olut"]], st["n_steps"], st["temp"], st["pres"], st["column"],
                        st["tvib"])
    got = eng.los_rt_lut([case["lut"]], case["steps"]).cpu().numpy()
    assert ref.max() > 0
    assert rel_err(got, ref) < TOL_RAD
    # host-buffer entry point
    got_h = eng.los_rt_lut_host([case["lut"]], case["steps"])
    assert np.array_equal(got_h, got)
    # a sub-range of grid points gives the same numbers (wavenumber chunking, smm:3190)
    sub = eng.los_rt_lut([case["lut"]], case["steps"], pt0=1001, n_pts=777).cpu().numpy()
    assert np.array_equal(sub, got[:, 1001:1778])


def test_los_blocking_and_chunking_are_invisible(case, monkeypatch):
    """LOS blocks / wavenumber chunks of the scratch (large batches) and the overlapped copies of
    the host entry point must not change a single bit; the thread-per-point kernel (SR_LOS_VER=1)
    agrees with the tensor-path product to rounding."""
    eng = case["engine"]
    base = eng.los_rt_lut([case["lut"]], case["steps"]).cpu().numpy()
    monkeypatch.setenv("SR_LOS_BLOCK", "4")
    monkeypatch.setenv("SR_LOS_CHUNK", "768")
    got = eng.los_rt_lut([case["lut"]], case["steps"]).cpu().numpy()
    assert np.array_equal(got, base)
    got_h = eng.los_rt_lut_host([case["lut"]], case["steps"])
    assert np.array_equal(got_h, base)
    monkeypatch.setenv("SR_LOS_BLOCK", "1")
    got_h = eng.los_rt_lut_host([case["lut"]], case["steps"], pt0=5, n_pts=3001)
    assert np.array_equal(got_h, base[:, 5:3006])
    monkeypatch.delenv("SR_LOS_BLOCK")
    monkeypatch.delenv("SR_LOS_CHUNK")
    monkeypatch.setenv("SR_LOS_VER", "1")
    v1 = eng.los_rt_lut([case["lut"]], case["steps"]).cpu().numpy()
    assert rel_err(v1, base) < 1e-12
    monkeypatch.delenv("SR_LOS_VER")
    # float32 layer scratch (the default of the low-res forward sink only): tau and J rounded to the
    # LUT's storage precision between the two kernels move a hi-res radiance by < 1e-6
    monkeypatch.setenv("SR_LOS_F32", "1")
    f32 = eng.los_rt_lut([case["lut"]], case["steps"]).cpu().numpy()
    assert 0 < rel_err(f32, base) < 1e-6
    monkeypatch.setenv("SR_LOS_BLOCK", "4")
    monkeypatch.setenv("SR_LOS_CHUNK", "768")
    assert np.array_equal(eng.los_rt_lut([case["lut"]], case["steps"]).cpu().numpy(), f32)


def test_lowres_fused_matches_hires_then_convolve(case, oracle, monkeypatch):
    """sr_los_rt_lut_lowres_dev (LOS blocks reduced to the channels on the device) = hi-res
    radiances followed by the instrument convolution, and both match the CPU restatement of
    convolve_to_grid_from_irregular (spect_classes.py:883-918) on the oracle's radiances."""
    eng, st, torch = case["engine"], case["st"], case["torch"]
    g = case["grid"]
    gdev = torch.as_tensor(g, device="cuda")
    centres = np.linspace(g[0] + 0.3, g[-1] - 0.3, 7)
    widths = np.full(7, 0.08)
    hi = eng.los_rt_lut([case["lut"]], case["steps"])
    ref = eng.convolve_lowres(gdev, hi, centres, widths).cpu().numpy()
    got = eng.los_rt_lut_lowres([case["lut"]], case["steps"], gdev, centres, widths).cpu().numpy()
    assert rel_err(got, ref) < 1e-8       # default float32 layer scratch of the low-res sink
    monkeypatch.setenv("SR_LOS_F32", "0")
    got = eng.los_rt_lut_lowres([case["lut"]], case["steps"], gdev, centres, widths).cpu().numpy()
    assert np.array_equal(got, ref)
    monkeypatch.setenv("SR_LOS_BLOCK", "4")          # 6 LOS -> blocks of 4 + 2
    monkeypatch.setenv("SR_LOS_CHUNK", "1024")
    got_b = eng.los_rt_lut_lowres([case["lut"]], case["steps"], gdev, centres, widths).cpu().numpy()
    assert np.array_equal(got_b, ref)
    monkeypatch.delenv("SR_LOS_BLOCK")
    monkeypatch.delenv("SR_LOS_CHUNK")
    rad = oracle.los_rt([case["olut"]], st["n_steps"], st["temp"], st["pres"], st["column"],
                        st["tvib"])
    cpu = np.stack([oracle.convolve_to_grid_from_irregular(g, rad[i], centres, widths) for i in range(len(rad))])
    assert rel_err(got, cpu) < TOL_RAD
    # a sub-window of the grid
    sub = eng.los_rt_lut_lowres([case["lut"]], case["steps"], gdev, centres[2:4], widths[2:4],
                                pt0=1000, n_pts=3000).cpu().numpy()
    ref_sub = eng.convolve_lowres(gdev[1000:4000].contiguous(), hi[:, 1000:4000].contiguous(),
                                  centres[2:4], widths[2:4]).cpu().numpy()
    assert np.array_equal(sub, ref_sub)


def test_materialised_layers_parity(case, oracle):
    eng, st, torch = case["engine"], case["st"], case["torch"]
    ref, tau_ref, src_ref = oracle.los_rt([case["olut"]], st["n_steps"], st["temp"], st["pres"],
                                          st["column"], st["tvib"], materialise=True)
    tau, src = eng.los_tau_src([case["lut"]], case["steps"])
    assert rel_err(tau.cpu().numpy(), tau_ref) < TOL_RAD
    assert rel_err(src.cpu().numpy(), src_ref) < TOL_RAD
    nst = torch.tensor(st["n_steps"], dtype=torch.int32, device="cuda")
    rad = eng.los_rt_layers(tau, src, nst).cpu().numpy()
    assert rel_err(rad, ref) < TOL_RAD
    # K3 alone against the CPU recursion on the very same materialised layers
    ref2 = oracle.los_layers(tau.cpu().numpy(), src.cpu().numpy(), st["n_steps"])
    assert rel_err(rad, ref2) < 1e-12
    # odd point count exercises the scalar (non-vectorised) kernel
    rad_odd = eng.los_rt_layers(tau[:, :, :1501].contiguous(), src[:, :, :1501].contiguous(), nst)
    assert np.array_equal(rad_odd.cpu().numpy(), rad[:, :1501])


def test_solo_absorption_and_initial_intensity(case, oracle):
    eng, st, torch = case["engine"], case["st"], case["torch"]
    n_los, n_grid = st["temp"].shape[0], len(case["grid"])
    i0 = np.full((n_los, n_grid), 3.0e-7)
    ref = oracle.los_rt([case["olut"]], st["n_steps"], st["temp"], st["pres"], st["column"],
                        st["tvib"], i0=i0, solo_absorption=True)
    got = eng.los_rt_lut([case["lut"]], case["steps"], i0=torch.tensor(i0, device="cuda"),
                         solo_absorption=True).cpu().numpy()
    assert rel_err(got, ref) < TOL_RAD
    assert np.all(got <= 3.0e-7 * (1 + 1e-12))       # pure attenuation
    ref = oracle.los_rt([case["olut"]], st["n_steps"], st["temp"], st["pres"], st["column"],
                        st["tvib"], i0=i0)
    got = eng.los_rt_lut([case["lut"]], case["steps"], i0=torch.tensor(i0, device="cuda")).cpu().numpy()
    assert rel_err(got, ref) < TOL_RAD


def test_lte_and_two_gases(case, oracle):
    """LTE (tvib=None -> T_vib = T, smm:2231-2232) and a second, LTE-unidentified gas ('all' set,
    pop = 1/Q, smm:2214-2218) summed into the same layers."""
    eng, S, st, torch = case["engine"], case["S"], case["st"], case["torch"]
    g = case["grid"]
    lines2 = S.line_table(120, 2996.0, 3005.0, n_levels=1, seed=8, q296=107.12, iso_ratio=0.986544)
    ls2 = eng.LineSet(lines2, g, 27.994915, 1)
    g32b = ls2.gcoeff_cells_f32(case["cells"])
    lut2 = eng.Lut(g32b, case["cells"], 5, 1, 0.986544, level_energies=None)
    olut2 = dict(g32=g32b.cpu().numpy(), pt=np.array(case["cells"]), level_energy=None, mol=5,
                 iso=1, iso_ratio=0.986544, lte_unidentified=True)
    col = np.concatenate([st["column"], 0.01 * st["column"]], axis=0)
    steps = eng.LosSteps(st["n_steps"], st["temp"], st["pres"], col, None)
    ref = oracle.los_rt([case["olut"], olut2], st["n_steps"], st["temp"], st["pres"], col, None)
    got = eng.los_rt_lut([case["lut"], lut2], steps).cpu().numpy()
    assert rel_err(got, ref) < TOL_RAD


def test_abscoeff_literal_numpy_restatement(case, oracle):
    """The C oracle's abs/emi assembly against the LITERAL NumPy restatement of
    make_abscoeff_LUTS_fast + LutSet.calculate (pins the oracle's index logic to the reference's
    own Python statements)."""
    st = case["st"]
    l = 2
    n = int(st["n_steps"][l])
    olut = case["olut"]
    a_np, e_np = oracle.make_abscoeff_LUTS_fast(olut, st["temp"][l, :n], st["pres"][l, :n],
                                                tvib=st["tvib"][0][:, l, :n])
    one = dict(n_steps=np.array([n], dtype=np.int32), temp=st["temp"][l:l + 1],
               pres=st["pres"][l:l + 1], column=np.ones_like(st["column"][:, l:l + 1]),
               tvib=st["tvib"][:, :, l:l + 1])
    lut1 = dict(olut, iso_ratio=1.0)
    _, tau, src = oracle.los_rt([lut1], one["n_steps"], one["temp"], one["pres"], one["column"],
                                one["tvib"], materialise=True)
    assert rel_err(tau[0, :n], a_np) < 1e-12
    J = tau[0, :n] * src[0, :n]
    assert rel_err(J, e_np, floor_rel=1e-12) < 1e-9


def test_lut_errors(case):
    eng, S, st = case["engine"], case["S"], case["st"]
    from spectrobot_b200._lib import SpectrobotError, SR_ERR_LUT
    bad = eng.LosSteps(st["n_steps"], st["temp"], st["pres"] * 1e6, st["column"], st["tvib"])
    with pytest.raises(SpectrobotError) as e:
        eng.los_rt_lut([case["lut"]], bad)
    assert e.value.code == SR_ERR_LUT


def test_fused_kernel_matches_two_kernel_path(case, oracle, monkeypatch):
    """k_los_fused2 (SR_LOS_VER=4: product + recursion in one kernel, LOS groups in shared memory)
    against the k_los_mma -> k_los_layers pair (SR_LOS_VER=3) and the oracle: plain, sub-window with
    a ragged tile, initial intensity, solo absorption, emission mask, low-res sink, two gases."""
    eng, st, torch, S = case["engine"], case["st"], case["torch"], case["S"]
    g = case["grid"]
    n_los, n_grid = st["temp"].shape[0], len(g)
    # more LOS than one LOS group holds, in an order that is NOT sorted (the planner sorts them)
    rep = np.array([3, 0, 5, 1, 4, 2] * 13)[:70]
    steps = eng.LosSteps(st["n_steps"][rep], st["temp"][rep], st["pres"][rep], st["column"][:, rep],
                         st["tvib"][:, :, rep])
    i0 = torch.tensor(np.linspace(1e-8, 4e-7, 70)[:, None] * np.ones((1, n_grid)), device="cuda")

    monkeypatch.setenv("SR_LOS_F32", "0")    # structural comparison: FP64 layer scratch in v3

    def run(ver, **kw):
        monkeypatch.setenv("SR_LOS_VER", str(ver))
        out = eng.los_rt_lut([case["lut"]], steps, **kw)
        monkeypatch.delenv("SR_LOS_VER")
        return out.cpu().numpy()
    base = run(3)
    fused = run(4)
    assert rel_err(fused, base) < 1e-12
    ref = oracle.los_rt([case["olut"]], st["n_steps"], st["temp"], st["pres"], st["column"], st["tvib"])
    assert rel_err(fused, ref[rep]) < TOL_RAD
    assert rel_err(run(4, pt0=1000, n_pts=3001), base[:, 1000:4001]) < 1e-12       # ragged last tile
    assert rel_err(run(4, i0=i0), run(3, i0=i0)) < 1e-12
    assert rel_err(run(4, i0=i0, solo_absorption=True), run(3, i0=i0, solo_absorption=True)) < 1e-12
    case["lut"].set_emission_mask(0b101)
    try:
        assert rel_err(run(4), run(3)) < 1e-12
    finally:
        case["lut"].set_emission_mask(None)
    # low-res sink: LOS blocks of one LOS group, rows scattered back to the caller's order
    gdev = torch.as_tensor(g, device="cuda")
    centres = np.linspace(g[0] + 0.3, g[-1] - 0.3, 7)
    widths = np.full(7, 0.08)
    monkeypatch.setenv("SR_LOS_VER", "3")
    low3 = eng.los_rt_lut_lowres([case["lut"]], steps, gdev, centres, widths).cpu().numpy()
    monkeypatch.setenv("SR_LOS_VER", "4")
    low4 = eng.los_rt_lut_lowres([case["lut"]], steps, gdev, centres, widths).cpu().numpy()
    monkeypatch.setenv("SR_LOS_BLOCK", "64")
    low4b = eng.los_rt_lut_lowres([case["lut"]], steps, gdev, centres, widths).cpu().numpy()
    monkeypatch.delenv("SR_LOS_BLOCK")
    monkeypatch.delenv("SR_LOS_VER")
    assert rel_err(low4, low3) < 1e-12 and np.array_equal(low4b, low4)
    # second, LTE-unidentified gas
    lines2 = S.line_table(120, 2996.0, 3005.0, n_levels=1, seed=8, q296=107.12, iso_ratio=0.986544)
    ls2 = eng.LineSet(lines2, g, 27.994915, 1)
    lut2 = eng.Lut(ls2.gcoeff_cells_f32(case["cells"]), case["cells"], 5, 1, 0.986544, level_energies=None)
    col = np.concatenate([st["column"][:, rep], 0.01 * st["column"][:, rep]], axis=0)
    st2 = eng.LosSteps(st["n_steps"][rep], st["temp"][rep], st["pres"][rep], col, None)
    monkeypatch.setenv("SR_LOS_VER", "3")
    two3 = eng.los_rt_lut([case["lut"], lut2], st2).cpu().numpy()
    monkeypatch.setenv("SR_LOS_VER", "4")
    two4 = eng.los_rt_lut([case["lut"], lut2], st2).cpu().numpy()
    monkeypatch.delenv("SR_LOS_VER")
    assert rel_err(two4, two3) < 1e-12
