"""CPU tests of the host-side helpers either side of the hot path against fixtures produced by
EXECUTING the reference's own Python (tests/golden/make_ref_golden2.py): SpectralObject slicing /
arithmetic / re-gridding / degraded grids, the older host convolution chain (convolve_to_grid,
hires_to_lowres_old, tolowres), prepare_fortran_sum, shape and black-body helpers, file-name
helpers, the AbsSetLOS and per-level LutSet streams written by the reference.
"""
import os

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "ref_golden2.npz"))


@pytest.fixture(scope="module")
def mods():
    from spectrobot_b200 import spect_classes as spcl, spect_main_module as smm, spect_base_module as sbm
    return spcl, smm, sbm


@pytest.fixture()
def spe(ref, mods):
    spcl = mods[0]
    return spcl.SpectralIntensity(ref["so_spec"].copy(), spcl.SpectralGrid(ref["so_grid"], units='cm_1'),
                                  units='ergscm2')


class Box(object):
    def put(self, item):
        self.item = item


def test_slicing_and_arithmetic(ref, mods, spe):
    """__getitem__ (:451-460), __add__ / __sub__ with a shorter spectrum (:462-483), add_to_spectrum
    (:929-953), add_to_spectrum_slow (:955-972), element-wise helpers (:693-701, 975-1013)."""
    spcl = mods[0]
    cut = spe[3002.0, 3003.5]
    assert np.array_equal(cut.spectral_grid.grid, ref["so_cut_grid"])
    assert np.array_equal(cut.spectrum, ref["so_cut_spec"])
    assert (spe[3020.0, 3021.0] is None) == bool(ref["so_cut_none"])
    part = spe[3004.0, 3006.0]
    part.spectrum = part.spectrum * 0.5
    assert np.array_equal((spe + part).spectrum, ref["so_add_short"])
    assert np.array_equal((spe - part).spectrum, ref["so_sub_short"])
    assert np.array_equal((spe + 0.25).spectrum, ref["so_add_scalar"])
    assert np.array_equal((spe * spe).spectrum, ref["so_mul"])
    assert np.array_equal(spe.spectrum, ref["so_spec"])          # operators do not touch self
    a = spcl.SpectralObject(spe.spectrum.copy(), spe.spectral_grid)
    a.add_to_spectrum(part, Strength=-2.0)
    assert np.array_equal(a.spectrum, ref["so_add_to_spectrum"])
    a = spcl.SpectralObject(spe.spectrum.copy(), spe.spectral_grid)
    a.add_to_spectrum_slow(part, Strength=3.0)
    assert np.array_equal(a.spectrum, ref["so_add_to_spectrum_slow"])
    assert np.array_equal(spe.exp_elementwise(-0.7).spectrum, ref["so_exp"])
    assert np.array_equal(spe.multiply_elementwise(spe, save=False).spectrum, ref["so_mulel"])
    den = spcl.SpectralObject(spe.spectrum + 1.0, spe.spectral_grid)
    assert np.array_equal(spe.divide_elementwise(den, save=False).spectrum, ref["so_divel"])
    with pytest.raises(ValueError):
        spe.multiply_elementwise(part)
    a = spcl.SpectralObject(spe.spectrum.copy(), spe.spectral_grid)
    assert np.array_equal(a.sum_scalar(1.5), ref["so_sum_scalar"])
    a.exp_elementwise(0.1, save=True)
    assert np.array_equal(a.spectrum, np.exp(ref["so_sum_scalar"] * 0.1))


def test_regridding(ref, mods, spe):
    """interp_to_grid (:501-507), interp_to_regular_grid (:920-927)."""
    spcl = mods[0]
    low = spe.interp_to_grid(spcl.SpectralGrid(ref["so_interp_grid"], units='cm_1'))
    assert np.array_equal(low.spectrum, ref["so_interp"])
    irr = spcl.SpectralObject(ref["so_irr_spec"].copy(), spcl.SpectralGrid(ref["so_irr_grid"], units='cm_1'))
    irr.interp_to_regular_grid()
    assert np.array_equal(irr.spectral_grid.grid, ref["so_reg_grid"])
    assert np.array_equal(irr.spectrum, ref["so_reg_spec"])


def test_degraded_grids(ref, mods, spe):
    """degrade_grid (:509-547), degrade_grid2 (:549-600), best_compressed_grid (smm:1513-1539)."""
    spcl, smm = mods[0], mods[1]
    cases = dict(d1=spe.degrade_grid(),
                 d1b=spe.degrade_grid(thress=[3e-2, 1e-2, 1e-3], factors=[4, 10, 50],
                                      consider_derivatives=False),
                 d2=spe.degrade_grid2(),
                 d2b=spe.degrade_grid2(thres=0.2, num_aside=[3, 2, 2], res_low=[1, 2, 5],
                                       consider_derivatives=False))
    for tag, d in cases.items():
        assert np.array_equal(d.spectral_grid.grid, ref["deg_%s_grid" % tag]), tag
        assert np.array_equal(d.spectrum, ref["deg_%s_spec" % tag]), tag
    weak = spcl.SpectralObject(spe.spectrum * 1e-3, spe.spectral_grid)
    assert np.array_equal(smm.best_compressed_grid([{'a': spe, 'b': weak}], alg=1), ref["best_grid_1"])
    assert np.array_equal(smm.best_compressed_grid([{'a': spe}, {'b': weak}], alg=2), ref["best_grid_2"])
    flat = spcl.SpectralObject(np.zeros(50), spcl.SpectralGrid(np.arange(50.) + 1, units='cm_1'))
    with pytest.raises(IndexError):
        flat.degrade_grid2(consider_derivatives=False)


def test_host_convolution_chain(ref, mods, spe):
    """convolve_to_grid (:832-881) with channel widths (channels inside, at the edges and outside
    the spectrum) and with the new grid's step; hires_to_lowres_old (:1193-1198); smm.tolowres
    (smm:3472-3477) with an nm observation."""
    spcl, smm = mods[0], mods[1]
    ch, wd = ref["cv_centres"], ref["cv_widths"]
    obs = spcl.SpectralIntensity(np.zeros(len(ch)), spcl.SpectralGrid(ch, units='cm_1'), units='nWcm2')
    got = spe.convolve_to_grid(obs.spectral_grid, spectral_widths=list(wd))
    assert rel_err(got.spectrum, ref["cv_result"]) < 1e-13
    assert got.spectrum[-1] == 0.0 and ref["cv_result"][-1] == 0.0       # window outside the spectrum
    reg = spcl.SpectralGrid(ref["cv_reg_grid"], units='cm_1')
    assert rel_err(spe.convolve_to_grid(reg).spectrum, ref["cv_reg_result"]) < 1e-13
    low = spe.hires_to_lowres_old(obs, spectral_widths=list(wd))
    assert [low.units, low.spectral_grid.units] == list(ref["cv_old_units"])
    assert rel_err(low.spectrum, ref["cv_old_result"]) < 1e-13
    hi = spcl.SpectralIntensity(ref["so_spec"].copy(), spcl.SpectralGrid(ref["so_grid"], units='cm_1'),
                                units='ergscm2')
    obs_nm = spcl.SpectralIntensity(np.zeros(len(ref["tl_centres"])),
                                    spcl.SpectralGrid(ref["tl_centres"], units='nm'), units='Wm2')
    obs_nm.add_bands(spcl.SpectralObject(ref["tl_widths"], obs_nm.spectral_grid))
    low = smm.tolowres(hi, obs_nm)
    assert [low.units, low.spectral_grid.units] == list(ref["tl_units"])
    assert rel_err(low.spectrum, ref["tl_result"]) < 1e-12
    assert hi.spectral_grid.units == 'nm'                                # converted in place


def test_prepare_fortran_sum(ref, mods, spe):
    """The staging matrix of add_lines_to_spectrum (:1100-1147): windows clipped at the low end,
    inside, clipped at the high end (zero padding on the left, shifted start) and of full length."""
    spcl = mods[0]
    big = spcl.SpectralObject(np.zeros(len(spe.spectrum)), spe.spectral_grid)
    wins = []
    for k, (g0, n) in enumerate(zip(ref["pfs_win_grid0"], ref["pfs_win_len"])):
        wins.append(spcl.SpectralObject(np.linspace(1.0, 2.0, n) * (1 + k),
                                        spcl.SpectralGrid(g0 + 0.01 * np.arange(n), units='cm_1')))
    # (the fixture's windows were laid out as grid[i0] + 0.01*(j - n//2): same points to rounding)
    box = Box()
    big.prepare_fortran_sum(wins, 0, box, fix_length=64)
    matrix, init, fin = box.item
    assert np.array_equal(init, ref["pfs_init"]) and np.array_equal(fin, ref["pfs_fin"])
    assert np.array_equal(matrix, ref["pfs_matrix"])
    assert matrix.flags['F_CONTIGUOUS']


def test_shape_and_blackbody_helpers(ref, mods):
    """closest_grid_ext (:1945-1964), Lorentz / Doppler / MakeShape_py (:1906-1923, 2011-2026),
    strengths and populations (:1856-1873, 1485-1488), black bodies (:1881-1903, 2080-2125)."""
    spcl = mods[0]
    sg = spcl.SpectralGrid(np.arange(3000.0, 3001.0, 0.01), units='cm_1')
    got = np.array([spcl.closest_grid_ext(sg, w) for w in ref["cge_probes"]], dtype=float)
    assert np.array_equal(got, ref["cge_result"])
    xs = ref["shape_x"]
    assert np.array_equal(spcl.Lorentz_shape(xs, 3000.1, 0.07), ref["shape_lorentz"])
    assert np.array_equal(spcl.Doppler_shape(xs, 3000.1, 0.004), ref["shape_doppler"])
    lg = spcl.SpectralGrid(ref["shape_py_grid"], units='cm_1')
    assert rel_err(spcl.MakeShape_py(lg, 3000.0, 0.004, 0.0045, Strength=2.0).spectrum,
                   ref["shape_py"]) < 1e-14
    assert spcl.Einstein_A_to_LineStrength_hitran(12.3, 3012.5, 180.0, 420.0, 15.0, 312.7,
                                                  iso_ab=0.988) == pytest.approx(float(ref["hit_strength"]), rel=1e-15)
    assert spcl.Boltz_pop_at_T(1533.3, 170.0, 3.0, 240.0) == pytest.approx(float(ref["boltz_pop"]), rel=1e-15)
    assert spcl.alpha_nlte(3019.5, 160.0, 1.3, 25.0) == pytest.approx(float(ref["alpha_nlte"]), rel=1e-15)
    bg = spcl.SpectralGrid(ref["bb_grid"], units='cm_1')
    assert rel_err(spcl.Calc_BB(bg, 180.0).spectrum, ref["bb_erg"]) < 1e-15
    bb = spcl.Calc_BB(bg, 180.0, units='Wm2')
    assert bb.units == 'Wm2' and rel_err(bb.spectrum, ref["bb_wm2"]) < 1e-15
    assert spcl.Calc_BB_single(3019.5, 94.0) == pytest.approx(float(ref["bb_single"]), rel=1e-15)
    got = [spcl.BB(180.0, 3019.5), spcl.BB_erg(180.0, 3019.5), spcl.BB_nm(180.0, 3311.0),
           spcl.BB_nm(5800.0, 500.0)]
    assert np.allclose(got, ref["bb_fun"], rtol=1e-15, atol=0.0)
    assert spcl.convert_cm_1_to_J(3019.5) == pytest.approx(float(ref["cm1_to_J"]), rel=1e-15)
    assert spcl.convert_cm_1_to_eV(8065.544) == pytest.approx(1.0, rel=1e-6)


def test_line_listing_and_nonlte_strength(ref, mods):
    """Print_hitran (:100-109) and CalcStrength_from_Strength (:256-288) of the fixture's lines."""
    spcl, smm, sbm = mods
    lines = spcl.read_line_database(os.path.join(GOLD, "ref_lines.par"))
    with open(os.devnull, 'w') as f:
        assert [lines[0].Print_hitran(ofile=f), lines[7].Print_hitran(ofile=f)] == \
            [str(v) for v in ref["print_hitran"]]
    iso1 = sbm.IsoMolec(6, 1)
    iso1.add_levels(['0 0 0 0 1A1', '0 0 1 0 1F2', '0 1 0 0 1E'], [0.0, 3019.4935, 1533.3326])
    got = [lin.CalcStrength_from_Strength(165.0, T_vib_lower=165.0, T_vib_upper=190.0)
           for lin in lines if lin.Iso == 1 and lin.LinkToMolec(iso1)]
    assert np.allclose(np.array(got), ref["strength_from_strength"], rtol=1e-9, atol=0.0)
    bands = smm.listbands(iso1, lines)
    assert sum(bands.values()) == len(got) and ('lev_01', 'lev_00') in bands
    arr, sums = spcl.sum_strength_lowres(lines, [2995.0, 3010.0], 6, 1, plot=False)
    tot = sum(s.sum() for s in sums.values())
    assert tot == pytest.approx(sum(l.Strength for l in lines if l.Iso == 1), rel=1e-12)


def test_name_helpers(ref, mods, tmp_path):
    """equiv (smm:53-66), find_free_name (smm:34-51), read_mw_list (:1465-1482)."""
    spcl, smm = mods[0], mods[1]
    got = [smm.equiv(0, 0), smm.equiv(0, 1e-20), smm.equiv(1.0, 1.0 + 5e-9), smm.equiv(1.0, 1.0 + 5e-8),
           smm.equiv(-2.0, -2.0), smm.equiv(3.0, 3.1, thres=0.1)]
    assert got == [bool(v) for v in ref["equiv"]]
    work = str(tmp_path) + '/'
    open(work + 'name.pic', 'w').close()
    open(work + 'name_001.pic', 'w').close()
    got = [os.path.basename(smm.find_free_name(work + 'name.pic')),
           os.path.basename(smm.find_free_name(work + 'other.pic')),
           os.path.basename(smm.find_free_name(work + 'name.pic', maxnum=50))]
    assert got == [str(v) for v in ref["free_name"]]
    with open(work + 'mw_list.dat', 'w') as f:
        f.write('2\n1 CH4_a 2900.0 2950.5\n2 HCN_b 3200.0 3290.0\n')
    assert spcl.read_mw_list(work) == (2, ['CH4_a', 'HCN_b'], [[2900.0, 2950.5], [3200.0, 3290.0]])


def test_absset_los_streams(ref, mods, tmp_path):
    """AbsSetLOS (smm:1179-1258): a stream written by the reference is read back one step at a
    time; the product's own stream round-trips; `set` mode keeps the objects."""
    spcl, smm = mods[0], mods[1]
    st = smm.AbsSetLOS(os.path.join(GOLD, "ref_abscoeff_los.pic"))
    st.counter = len(ref["absset_rows"])
    st.prepare_read()
    assert np.array_equal(st.spectral_grid.grid, ref["absset_grid"]) and st.remaining == 3
    for row in ref["absset_rows"]:
        one = st.read_one()
        assert np.array_equal(one.spectrum, row) and one.spectral_grid is st.spectral_grid
    assert st.remaining == 0
    st.finalize_IO()
    sg = spcl.SpectralGrid(ref["absset_grid"], units='cm_1')
    mine = smm.AbsSetLOS(str(tmp_path / 'mine.pic'), spectral_grid=sg)
    mine.prepare_export()
    for row in ref["absset_rows"]:
        mine.add_dump(spcl.SpectralObject(row, sg, link_grid=True))
    mine.finalize_IO()
    assert mine.counter == 3
    back = [mine.read_one().spectrum for _ in range(3)]
    mine.finalize_IO()
    assert np.array_equal(np.array(back), ref["absset_rows"])
    keep = smm.AbsSetLOS(None, spectral_grid=sg)
    keep.add_set(spcl.SpectralObject(ref["absset_rows"][0], sg, link_grid=True))
    assert len(keep) == 1 and keep[0] is keep.set[0] and [k for k in keep] == keep.set
    with pytest.raises(ValueError):
        keep.prepare_export()


def test_lutset_streams(mods, tmp_path):
    """LutSet.load_from_file / load_from_files / prepare_read + load_singlePT_from_file
    (smm:872-979) on the per-level file written by the reference's LookUpTable.make, and
    prepare_export / add_dump / finalize_IO writing a stream the same readers accept."""
    spcl, smm, sbm = mods
    ref1 = np.load(os.path.join(GOLD, "ref_golden.npz"))
    fn = os.path.join(GOLD, "ref_LUT_mol06_iso1_nonLTE_lev_01.pic")
    sp = smm.prepare_spe_grid([2998.0, 3006.0]).spectral_grid
    st = smm.LutSet(6, 1, 16.0313, level=None, filename=fn)
    st.load_from_file(load_just_PT=True)
    assert np.array_equal(np.array(st.PTcouples), ref1["cells_PT"]) and st.sets == []
    st.load_from_file(spectral_grid=sp)
    assert len(st.sets) == 2
    for k, ct in enumerate(smm.CTYPES):
        co = st.sets[0][ct]
        assert co.spectrum.dtype == np.float64 and np.array_equal(co.spectrum, ref1["cells_nonlte"][1, k])
        assert np.array_equal(co.spectral_grid.grid, sp.grid)
    st2 = smm.LutSet(6, 1, 16.0313, level=None, filename=fn)
    st2.load_from_file(load_just_PT=True)
    st2.add_file(fn, [])
    st2.load_from_files(spectral_grid=sp, cartLUTs=GOLD)
    assert len(st2.sets) == 4 and len(st2.PTcouples) == 4
    st3 = smm.LutSet(6, 1, 16.0313, level=None, filename=fn)
    assert np.array_equal(np.array(st3.prepare_read()), ref1["cells_PT"])
    one = st3.load_singlePT_from_file(spectral_grid=sp)
    assert np.array_equal(one['absorption'].spectrum, ref1["cells_nonlte"][1, 2])
    st3.finalize_IO()
    assert st3.temp_file is None
    # write: header, two cells, read back
    out = smm.LutSet(6, 1, 16.0313, level=None, filename=str(tmp_path / 'lev.pic'))
    out.prepare_export([[0.05, 150.0], [2.0, 155.0]], sp)
    for set_ in st.sets:
        d = dict()
        for ct, co in set_.items():
            c2 = spcl.SpectralGcoeff(ct, sp, 6, 1, 16.0313, '', spectrum=co.spectrum, Pres=co.pres, Temp=co.temp)
            c2.erase_grid()
            d[ct] = c2
        out.add_dump(d)
    out.finalize_IO()
    pts, sets = smm.read_lutset_stream(out.filename)
    assert pts == [[0.05, 150.0], [2.0, 155.0]]
    assert np.array_equal(sets[1]['sp_emission'].spectrum, st.sets[1]['sp_emission'].spectrum)
    with pytest.raises(ValueError):
        smm.LutSet(6, 1, 16.0313).prepare_read()


@pytest.mark.skipif(not os.path.exists('/root/reference/spect_classes.py'),
                    reason="reference tree not present (GPU box)")
def test_fixtures2_are_what_the_reference_computes_now(ref):
    """Re-executes a slice of the reference live (this container only): the committed fixture is
    the reference's output, not a hand edit."""
    import sys
    sys.path.insert(0, GOLD)
    import ref_exec as R
    import make_ref_golden2 as M2
    spcl, smm, sbm = R.load()
    spe = M2.test_spectrum(spcl)
    assert np.array_equal(spe.spectrum, ref["so_spec"])
    with R.quiet():
        d2 = spe.degrade_grid2()
    assert np.array_equal(d2.spectral_grid.grid, ref["deg_d2_grid"])
    obs = spcl.SpectralGrid(ref["cv_centres"], units='cm_1')
    assert np.array_equal(spe.convolve_to_grid(obs, spectral_widths=list(ref["cv_widths"])).spectrum,
                          ref["cv_result"])
    assert spcl.BB_nm(180.0, 3311.0) == ref["bb_fun"][2]
