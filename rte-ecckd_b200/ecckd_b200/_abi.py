"""ctypes declarations of include/ecckd_b200.h.

The shared library is built in-tree (rte-ecckd_b200/lib/libecckd_b200.so) by
rte-ecckd_b200/csrc/Makefile.  There is no Python or CPU implementation behind
these calls: if the library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

ERRLEN = 128
NAME_LEN = 32
MAX_GASES = 16
OK, ERR_ARG, ERR_CUDA, ERR_REFERENCE, ERR_IO = 0, 1, 2, 3, 4
KERNEL_AUTO, KERNEL_WALK, KERNEL_GENERIC, KERNEL_QUAD, KERNEL_COLWALK = 0, 1, 2, 3, 4

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libecckd_b200.so")


class VmrView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("col_stride", C.c_int64), ("lay_stride", C.c_int64)]


class GasTable(C.Structure):
    _fields_ = [("name", C.c_char * NAME_LEN), ("composite_only", C.c_int32),
                ("concentration_dependence_code", C.c_int32), ("n_mole_fraction", C.c_int32),
                ("reserved", C.c_int32), ("reference_mole_fraction", C.c_double),
                ("mole_fraction", C.c_void_p), ("coefficient", C.c_void_p)]


class HostTables(C.Structure):
    _fields_ = [("shortwave", C.c_int32), ("ngpt", C.c_int32), ("npress", C.c_int32), ("ntemp", C.c_int32),
                ("nplanck", C.c_int32), ("num_gases", C.c_int32),
                ("log_pressure", C.c_void_p), ("temperature", C.c_void_p), ("temperature_planck", C.c_void_p),
                ("planck_function", C.c_void_p), ("rayleigh_molar_scattering_coeff", C.c_void_p),
                ("solar_irradiance", C.c_void_p), ("gas", C.POINTER(GasTable)),
                ("nband", C.c_int32), ("reserved", C.c_int32), ("band_lims_wvn", C.c_void_p), ("band2gpt", C.c_void_p)]


# every symbol include/ecckd_b200.h declares (tests/test_abi.py checks the export list against the header)
_vp, _i, _cp = C.c_void_p, C.c_int, C.c_char_p
_names = C.POINTER(C.c_char_p)
_views = C.POINTER(VmrView)
_SIGS = {
    "ecckd_version": (C.c_char_p, []),
    "ecckd_debug_walk_plan": (None, [_i, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ecckd_model_create": (_i, [C.POINTER(HostTables), _i, C.POINTER(C.c_int), C.POINTER(_vp), _cp]),
    "ecckd_model_load": (_i, [_cp, _i, C.POINTER(C.c_int), C.POINTER(_vp), _cp]),
    "ecckd_model_destroy": (None, [_vp]),
    "ecckd_tables_load": (_i, [_cp, C.POINTER(C.POINTER(HostTables)), _cp]),
    "ecckd_tables_free": (None, [C.POINTER(HostTables)]),
    "ecckd_get_ngas": (_i, [_vp]),
    "ecckd_get_gas_name": (_i, [_vp, _i, _cp]),
    "ecckd_get_ngpt": (_i, [_vp]),
    "ecckd_get_nband": (_i, [_vp]),
    "ecckd_get_band2gpt": (_i, [_vp, _vp]),
    "ecckd_get_band_lims_wvn": (_i, [_vp, _vp]),
    "ecckd_get_table_warnings": (C.c_char_p, [_vp]),
    "ecckd_host_register": (_i, [_vp, C.c_size_t, _cp]),
    "ecckd_host_unregister": (_i, [_vp]),
    "ecckd_get_ndev": (_i, [_vp]),
    "ecckd_get_device": (_i, [_vp, _i]),
    "ecckd_source_is_internal": (_i, [_vp]),
    "ecckd_source_is_external": (_i, [_vp]),
    "ecckd_get_press_min": (C.c_double, [_vp]),
    "ecckd_get_press_max": (C.c_double, [_vp]),
    "ecckd_get_temp_min": (C.c_double, [_vp]),
    "ecckd_get_temp_max": (C.c_double, [_vp]),
    "ecckd_get_total_solar_irradiance": (C.c_double, [_vp]),
    "ecckd_interpolation_indices_f64": (_i, [_vp, _i, _i, _i, _vp, _vp, _cp, _views, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _cp]),
    "ecckd_rte_lw_f64": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _cp]),
    "ecckd_rte_sw_f64": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp, _vp, _vp, _vp, _cp]),
    "ecckd_debug_div_const": (_i, [C.c_int64, _vp, C.c_double, C.POINTER(C.c_uint64), _cp]),
    "ecckd_planck_indices_f64": (_i, [_vp, _i, C.c_int64, _vp, _vp, _vp, _vp, _cp]),
}
for _p in ("f64", "f32"):
    _SIGS[f"ecckd_gas_optics_lw_{_p}"] = (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _names, _views,
                                              _vp, _vp, _vp, _vp, _vp, _i, _vp, _cp])
    _SIGS[f"ecckd_gas_optics_sw_{_p}"] = (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _names, _views,
                                              _vp, _vp, _vp, _vp, _i, _vp, _cp])
    _SIGS[f"ecckd_gas_optics_lw_host_{_p}"] = (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _names, _views,
                                                   _vp, _vp, _vp, _vp, _vp, _i, _cp])
    _SIGS[f"ecckd_gas_optics_sw_host_{_p}"] = (_i, [_vp, _i, _i, _vp, _vp, _i, _names, _views,
                                                   _vp, _vp, _vp, _vp, _i, _cp])

EXPORTS = tuple(sorted(_SIGS))
_lib = None


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """dlopen libecckd_b200.so and attach the prototypes.  No fallback."""
    global _lib
    if _lib is not None and path == LIB_PATH:
        return _lib
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `make -C rte-ecckd_b200/csrc` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    lib = C.CDLL(path)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path == LIB_PATH:
        _lib = lib
    return lib
