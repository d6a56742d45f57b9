"""ctypes binding of libspectrobot.so (the C ABI declared in include/spectrobot.h).

There is no CPU fallback: if the shared library is missing this module raises at import of the
first symbol, and every compute entry point returns SR_ERR_CUDA when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspectrobot.so")

SR_OK = 0
SR_ERR_ARG, SR_ERR_DW, SR_ERR_CUDA, SR_ERR_GEOMETRY, SR_ERR_TABLE, SR_ERR_LUT, SR_ERR_LIMIT = \
    1, 2, 3, 4, 5, 6, 7
IMXSIG = 13010
IMXLINES = 40000
IMXSIG_LONG = 2000000
IMXSTP = 8000

_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


class SpectrobotError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "libspectrobot status %d: %s" % (code, msg))
        self.code = code


class sr_consts(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("h_cgs", "c_cgs", "k_cgs", "avogadro", "ln2",
                                          "sqrt_ln2", "sqrt_pi_ln2")]


class sr_lines(C.Structure):
    _fields_ = [("n_lines", C.c_int)] + [(n, _dp) for n in (
        "freq", "a_coeff", "air_broad", "t_dep", "e_lower", "g_up", "g_lo", "e_vib_up",
        "e_vib_lo")] + [("up_set", _ip), ("lo_set", _ip)]


class sr_los_steps(C.Structure):
    _fields_ = [("n_los", C.c_int), ("n_steps_max", C.c_int), ("n_gas", C.c_int),
                ("n_sets_max", C.c_int), ("n_steps", _ip), ("temp", _dp), ("pres", _dp),
                ("column", _dp), ("tvib", _dp)]


class sr_atmosphere(C.Structure):
    _fields_ = [("n_band", C.c_int), ("n_z", C.c_int), ("n_gas", C.c_int), ("n_sets_max", C.c_int),
                ("lat_edges", _dp), ("z", _dp), ("temp", _dp), ("pres", _dp), ("vmr", _dp),
                ("tvib", _dp), ("tvib_on", _ip), ("radius_km", C.c_double), ("top_km", C.c_double),
                ("n_sza", C.c_int), ("sza_nodes", _dp)]


class sr_los_rays(C.Structure):
    _fields_ = [("n_los", C.c_int), ("origin", _dp), ("direction", _dp), ("sun", _dp),
                ("sza_fixed", _dp)]


class sr_steps_opt(C.Structure):
    _fields_ = [("delta_x_km", C.c_double), ("max_T_variation", C.c_double),
                ("max_Plog_variation", C.c_double), ("max_opt_depth", C.c_double),
                ("sigma_peak", _dp), ("photon_order", C.c_int)]


class sr_channels(C.Structure):
    _fields_ = [("n_chan", C.c_int), ("centre_dev", _vp), ("width_dev", _vp),
                ("n_sigma", C.c_double), ("units", C.c_int)]


SR_CHAN_SAME_UNITS, SR_CHAN_NM_FROM_CM1 = 0, 1
SR_PROF_LOS_MMA, SR_PROF_LOS_LAYERS, SR_PROF_CONV, SR_PROF_VOIGT_TILE, SR_PROF_VOIGT_CORE, \
    SR_PROF_LOS_FUSED = 0, 1, 2, 3, 4, 5

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header
SIGNATURES = {
    "sr_version": (C.c_int, []),
    "sr_last_error": (C.c_char_p, []),
    "sr_kernel_launch_count": (C.c_longlong, []),
    "sr_device_count": (C.c_int, []),
    "sr_set_device": (C.c_int, [C.c_int]),
    "sr_humliv_bb": (C.c_int, [_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                               C.c_double, _dp]),
    "sr_sum_all_lines": (C.c_int, [_dp, _dp, _ip, _ip, C.c_int, C.c_int, C.c_int, C.c_int, _dp]),
    "sr_bd_tips_2003": (C.c_int, [C.c_int, C.c_int, _dp, _dp, _dp]),
    "sr_partition_sum": (C.c_int, [C.c_int, C.c_int, C.c_double, _dp]),
    "sr_curgod": (C.c_int, [C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp]),
    "sr_curgod_1": (C.c_int, [_dp, _dp, C.c_int, _dp]),
    "sr_curgod_2": (C.c_int, [_dp, _dp, _dp, C.c_int, _dp]),
    "sr_curgod_3": (C.c_int, [_dp, _dp, _dp, _dp, C.c_int, _dp]),
    "sr_curgod_4": (C.c_int, [_dp, _dp, _dp, _dp, C.c_int, _dp]),
    "sr_default_consts": (None, [C.POINTER(sr_consts)]),
    "sr_lineset_create": (C.c_int, [C.POINTER(sr_lines), _dp, C.c_long, _dp, C.c_int, C.c_double,
                                    C.POINTER(sr_consts), C.POINTER(_vp)]),
    "sr_lineset_destroy": (C.c_int, [_vp]),
    "sr_lineset_n_active": (C.c_long, [_vp]),
    "sr_lineset_centres": (C.c_int, [_vp, _ip]),
    "sr_lineset_order": (C.c_int, [_vp, _ip]),
    "sr_lineset_check": (C.c_int, [_vp, _vp]),
    "sr_gcoeff_cells_dev": (C.c_int, [_vp, _dp, C.c_int, _vp, _vp]),
    "sr_gcoeff_cells_host": (C.c_int, [_vp, _dp, C.c_int, _dp]),
    "sr_gcoeff_cells_dev_f32": (C.c_int, [_vp, _dp, C.c_int, _vp, _vp, _vp]),
    "sr_gcoeff_cells_dev_f32_ld": (C.c_int, [_vp, _dp, C.c_int, _vp, C.c_long, _vp]),
    "sr_gcoeff_cells_window_dev": (C.c_int, [_vp, _dp, C.c_int, _vp, C.c_int, C.c_long, C.c_long,
                                             C.c_long, _vp]),
    "sr_lineset_tile_points": (C.c_int, [_vp]),
    "sr_line_shapes_dev": (C.c_int, [_vp, C.c_double, C.c_double, _vp, _vp, _vp]),
    "sr_los_rt_layers_dev": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_long, _vp, C.c_int,
                                       _vp, _vp]),
    "sr_lut_create": (C.c_int, [_vp, _dp, C.c_int, C.c_int, C.c_long, _dp, C.c_int, C.c_int,
                                C.c_double, C.c_int, C.POINTER(sr_consts), C.POINTER(_vp)]),
    "sr_lut_create_ld": (C.c_int, [_vp, C.c_long, _dp, C.c_int, C.c_int, C.c_long, _dp, C.c_int,
                                   C.c_int, C.c_double, C.c_int, C.POINTER(sr_consts),
                                   C.POINTER(_vp)]),
    "sr_lut_destroy": (C.c_int, [_vp]),
    "sr_lut_set_emission_mask": (C.c_int, [_vp, C.c_ulonglong]),
    "sr_los_rt_lut_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_long, C.c_long,
                                    _vp, C.c_int, _vp, _vp]),
    "sr_los_rt_lut_lowres_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_long,
                                           C.c_long, _vp, _vp, _vp, C.c_int, C.c_double, _vp,
                                           C.c_int, _vp, _vp]),
    "sr_los_rt_lut_host": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_long, C.c_long,
                                     _dp, C.c_int, _dp]),
    "sr_los_tau_src_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_long, C.c_long,
                                     _vp, _vp, _vp]),
    "sr_los_abs_emi_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_long, C.c_long,
                                     _vp, _vp, _vp]),
    "sr_los_rt_layers_jac_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int,
                                           C.c_long, _vp, C.c_int, _vp, _vp, _vp]),
    "sr_los_rt_lut_jac_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_int, _ip, _dp,
                                        C.c_long, C.c_long, _vp, C.c_int, _vp, _vp, _vp]),
    "sr_los_rt_lut_jac_lowres_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_int,
                                               _ip, _dp, C.c_long, C.c_long, _vp, _vp, _vp,
                                               C.c_int, C.c_double, _vp, C.c_int, _vp, _vp, _vp]),
    "sr_los_check": (C.c_int, [C.POINTER(_vp), _vp]),
    "sr_lut_weights": (C.c_int, [_dp, C.c_int, C.c_double, C.c_double, _ip, _dp]),
    "sr_convolve_lowres_dev": (C.c_int, [_vp, C.c_long, _vp, C.c_int, _vp, _vp, C.c_int,
                                         C.c_double, _vp, _vp]),
    "sr_convolve_lowres_host": (C.c_int, [_dp, C.c_long, _dp, C.c_int, _dp, _dp, C.c_int,
                                          C.c_double, _dp]),
    "sr_convolve_channels_dev": (C.c_int, [_vp, C.c_long, _vp, C.c_int, C.POINTER(sr_channels), _vp,
                                           _vp]),
    "sr_convolve_channels_host": (C.c_int, [_dp, C.c_long, _dp, C.c_int, _dp, _dp, C.c_int,
                                            C.c_double, C.c_int, _dp]),
    "sr_los_rt_lut_channels_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_long,
                                             C.c_long, _vp, C.POINTER(sr_channels), _vp, C.c_int,
                                             _vp, _vp]),
    "sr_los_rt_lut_jac_channels_dev": (C.c_int, [C.POINTER(_vp), C.POINTER(sr_los_steps), C.c_int,
                                                 _ip, _dp, C.c_long, C.c_long, _vp,
                                                 C.POINTER(sr_channels), _vp, C.c_int, _vp, _vp,
                                                 _vp]),
    "sr_los_steps_build": (C.c_int, [C.POINTER(sr_atmosphere), C.c_int, _dp, _dp, C.c_double,
                                     C.c_double, C.c_double, C.c_int, _dp, C.c_int, C.c_int, _ip,
                                     _dp, _dp, _dp, _dp, _dp, _ip]),
    "sr_los_steps_build_rays": (C.c_int, [C.POINTER(sr_atmosphere), C.POINTER(sr_los_rays),
                                          C.POINTER(sr_steps_opt), C.c_int, _dp, C.c_int, C.c_int,
                                          _ip, _dp, _dp, _dp, _dp, _dp, _ip]),
    "sr_prof_enable": (C.c_int, [C.c_int]),
    "sr_prof_summary": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), _dp, _dp]),
    "sr_fp64_peak": (C.c_int, [C.c_int, _dp]),
}

_LIB = None


def lib():
    """Load libspectrobot.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "spectrobot_b200: %s is missing - build it with `make -C spectrobot_b200/csrc` "
                "(there is no CPU fallback for the hot path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(code):
    if code != SR_OK:
        raise SpectrobotError(code, lib().sr_last_error().decode("utf-8", "replace"))


def dptr(a):
    return a.ctypes.data_as(_dp)


def iptr(a):
    return a.ctypes.data_as(_ip)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def python_consts():
    """sr_consts filled the way spect_classes.py:44-47,1984,1997 forms them (installed scipy)."""
    import math as mt
    import scipy.constants as const
    c = sr_consts()
    c.h_cgs = const.physical_constants['Planck constant'][0] * 1.e7
    c.c_cgs = const.c * 1.e2
    c.k_cgs = const.physical_constants['Boltzmann constant'][0] * 1.e7
    c.avogadro = const.Avogadro
    c.ln2 = mt.log(2.0)
    c.sqrt_ln2 = mt.sqrt(mt.log(2.0))
    c.sqrt_pi_ln2 = mt.sqrt(np.pi / mt.log(2.0))
    return c


def cuda_available():
    return lib().sr_device_count() > 0
