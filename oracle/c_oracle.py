"""ctypes front end of oracle/ecckd_oracle.c.  TEST INFRASTRUCTURE ONLY.

Arrays are numpy float64 in Fortran order with the reference's index order:
plev (ncol, nlay+1), tlay (ncol, nlay), tau (ncol, nlay, ngpt).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

from .ckd_model import CkdModel

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
MAX_GASES = 16

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Model(C.Structure):
    _fields_ = [
        ("shortwave", C.c_int), ("ngpt", C.c_int), ("npress", C.c_int), ("ntemp", C.c_int),
        ("nplanck", C.c_int), ("num_gases", C.c_int),
        ("log_pressure", _dp), ("temperature", _dp), ("planck_function", _dp),
        ("temperature_planck", _dp), ("rayleigh", _dp), ("solar_irradiance", _dp),
        ("gas", (C.c_char * 32) * MAX_GASES),
        ("composite_only", C.c_int * MAX_GASES),
        ("code", C.c_int * MAX_GASES),
        ("nmole", C.c_int * MAX_GASES),
        ("mole_fraction", _dp * MAX_GASES),
        ("reference_mole_fraction", C.c_double * MAX_GASES),
        ("coefficient", _dp * MAX_GASES),
    ]


class _Concs(C.Structure):
    _fields_ = [("ngas", C.c_int), ("names", C.POINTER(C.c_char_p)),
                ("values", C.POINTER(_dp)), ("kinds", _ip)]


def build(fast: bool = False, force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile) and return the .so path."""
    name = "libecckd_oracle_fast.so" if fast else "libecckd_oracle.so"
    path = os.path.join(_BUILD, name)
    src = os.path.join(_HERE, "ecckd_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", os.path.join("_build", name)],
                       check=True, stdout=subprocess.DEVNULL)
    return path


_LIBS: Dict[bool, C.CDLL] = {}


def load(fast: bool = False, rebuild: bool = False) -> C.CDLL:
    if fast in _LIBS and not rebuild:
        return _LIBS[fast]
    lib = C.CDLL(build(fast, force=rebuild))
    lib.oracle_gas_optics_int.restype = C.c_int
    lib.oracle_gas_optics_ext.restype = C.c_int
    lib.oracle_run_blocks.restype = C.c_long
    lib.oracle_max_threads.restype = C.c_int
    _LIBS[fast] = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_dp) if a is not None else _dp()


def _farr(a, shape=None) -> np.ndarray:
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"shape {a.shape} != {shape}")
    return a


GasConcs = Sequence[Tuple[str, Union[float, np.ndarray]]]


class COracle:
    """The C oracle bound to one CkdModel."""

    def __init__(self, model: CkdModel, fast: bool = False):
        self.lib = load(fast)
        self.model = model
        self._keep = []
        m = _Model()
        m.shortwave = int(model.shortwave)
        m.ngpt = model.ngpt
        m.npress = model.log_pressure.shape[0]
        m.ntemp = model.temperature.shape[0]
        m.nplanck = 0 if model.temperature_planck is None else model.temperature_planck.shape[0]
        m.num_gases = model.num_gases

        def keep(a):
            if a is None:
                return _dp()
            a = np.ascontiguousarray(a, dtype=np.float64)   # C order of reversed dims == Fortran order
            self._keep.append(a)
            return a.ctypes.data_as(_dp)

        m.log_pressure = keep(model.log_pressure)
        m.temperature = keep(model.temperature)
        m.planck_function = keep(model.planck_function)
        m.temperature_planck = keep(model.temperature_planck)
        m.rayleigh = keep(model.rayleigh_molar_scattering_coeff)
        m.solar_irradiance = keep(model.solar_irradiance)
        for i, (name, tab) in enumerate(zip(model.gas, model.absorption)):
            m.gas[i].value = name.encode().ljust(31)[:31]
            m.composite_only[i] = int(tab.composite_only)
            m.code[i] = tab.concentration_dependence_code
            m.nmole[i] = 0 if tab.mole_fraction is None else tab.mole_fraction.shape[0]
            m.mole_fraction[i] = keep(tab.mole_fraction)
            m.reference_mole_fraction[i] = tab.reference_mole_fraction
            m.coefficient[i] = keep(tab.coefficient)
        self._m = m

    # -- helpers -------------------------------------------------------------
    def _concs(self, gas_concs: GasConcs, ncol: int, nlay: int):
        n = len(gas_concs)
        names = (C.c_char_p * max(n, 1))()
        values = (_dp * max(n, 1))()
        kinds = (C.c_int * max(n, 1))()
        keep = []
        for j, (name, v) in enumerate(gas_concs):
            names[j] = name.lower().encode()
            a = np.asarray(v, dtype=np.float64)
            if a.ndim == 0:
                a = a.reshape(1).copy(); kinds[j] = 0
            elif a.ndim == 1:
                if a.shape[0] != nlay:
                    raise ValueError(f"gas {name}: profile length {a.shape[0]} != nlay {nlay}")
                a = np.ascontiguousarray(a); kinds[j] = 1
            else:
                a = _farr(a, (ncol, nlay)); kinds[j] = 2
            keep.append(a)
            values[j] = a.ctypes.data_as(_dp)
        c = _Concs(n, names, values, kinds)
        return c, (names, values, kinds, keep)

    # -- entry points --------------------------------------------------------
    def gas_optics_int(self, plev, tlay, tsfc, gas_concs: GasConcs, tlev=None):
        """gas_optics_ecckd.f90:381-426.  Returns (outputs, error_msg)."""
        tlay = _farr(tlay); ncol, nlay = tlay.shape
        plev = _farr(plev, (ncol, nlay + 1)); tsfc = _farr(tsfc, (ncol,))
        tlev_a = _farr(tlev, (ncol, nlay + 1)) if tlev is not None else None
        ng = self.model.ngpt
        out = {k: np.full((ncol, nlay, ng), np.nan, order="F") for k in
               ("tau", "lay_source", "lev_source_inc", "lev_source_dec")}
        out["sfc_source"] = np.full((ncol, ng), np.nan, order="F")
        c, keep = self._concs(gas_concs, ncol, nlay)
        err = C.create_string_buffer(128)
        self.lib.oracle_gas_optics_int(C.byref(self._m), ncol, nlay, _ptr(plev), _ptr(tlay), _ptr(tsfc),
                                       _ptr(tlev_a), C.byref(c), _ptr(out["tau"]), _ptr(out["lay_source"]),
                                       _ptr(out["lev_source_inc"]), _ptr(out["lev_source_dec"]),
                                       _ptr(out["sfc_source"]), err)
        return out, err.value.decode()

    def gas_optics_ext(self, plev, tlay, gas_concs: GasConcs, two_stream: bool = True):
        """gas_optics_ecckd.f90:431-473.  Returns (outputs, error_msg)."""
        tlay = _farr(tlay); ncol, nlay = tlay.shape
        plev = _farr(plev, (ncol, nlay + 1))
        ng = self.model.ngpt
        out = {k: np.full((ncol, nlay, ng), np.nan, order="F") for k in ("tau", "ssa", "g")}
        out["toa_src"] = np.full((ncol, ng), np.nan, order="F")
        c, keep = self._concs(gas_concs, ncol, nlay)
        err = C.create_string_buffer(128)
        self.lib.oracle_gas_optics_ext(C.byref(self._m), ncol, nlay, _ptr(plev), _ptr(tlay), C.byref(c),
                                       int(two_stream), _ptr(out["tau"]), _ptr(out["ssa"]), _ptr(out["g"]),
                                       _ptr(out["toa_src"]), err)
        return out, err.value.decode()

    def indices(self, gas_name: str, plev, tlay, layer_vmr):
        """1-based ip0, it0, iv0 and the upper weights, as the reference computes them."""
        tlay = _farr(tlay); ncol, nlay = tlay.shape
        plev = _farr(plev, (ncol, nlay + 1)); vmr = _farr(layer_vmr, (ncol, nlay))
        gi = self.model.gas.index(gas_name)
        ints = {k: np.zeros((ncol, nlay), dtype=np.int32, order="F") for k in ("ip0", "it0", "iv0")}
        dbl = {k: np.zeros((ncol, nlay), order="F") for k in ("pw1", "tw1", "vw1")}
        self.lib.oracle_indices(C.byref(self._m), gi, ncol, nlay, _ptr(plev), _ptr(tlay), _ptr(vmr),
                                ints["ip0"].ctypes.data_as(_ip), _ptr(dbl["pw1"]),
                                ints["it0"].ctypes.data_as(_ip), _ptr(dbl["tw1"]),
                                ints["iv0"].ctypes.data_as(_ip), _ptr(dbl["vw1"]))
        return {**ints, **dbl}

    def planck_indices(self, temperature):
        t = _farr(temperature)
        if t.ndim == 1:
            t = t.reshape(-1, 1, order="F")
        it0 = np.zeros(t.shape, dtype=np.int32, order="F"); w1 = np.zeros(t.shape, order="F")
        self.lib.oracle_planck_indices(C.byref(self._m), t.shape[0], t.shape[1], _ptr(t),
                                       it0.ctypes.data_as(_ip), _ptr(w1))
        return it0, w1

    def constants(self) -> np.ndarray:
        out = np.zeros(5)
        self.lib.oracle_constants(_ptr(out))
        return out

    def run_blocks(self, lw: bool, block: int, nlay: int, nblocks: int, plev, tlay, tsfc, tlev,
                   gas_concs: GasConcs, nthreads: int) -> float:
        """Timed-baseline driver (oracle_run_blocks).  Inputs are blocked C-contiguous-by-block
        arrays: plev (nblocks, nlay+1, block) in numpy C order == Fortran (block, nlay+1, nblocks)."""
        n = len(gas_concs)
        names = (C.c_char_p * n)(*[g[0].encode() for g in gas_concs])
        kinds = (C.c_int * n)()
        values = (_dp * n)()
        keep = []
        for j, (_, v) in enumerate(gas_concs):
            a = np.ascontiguousarray(np.asarray(v, dtype=np.float64))
            if a.ndim == 0:
                a = a.reshape(1).copy(); kinds[j] = 0
            elif a.ndim == 1:
                kinds[j] = 1
            else:
                kinds[j] = 2
            keep.append(a); values[j] = a.ctypes.data_as(_dp)
        chk = C.c_double(0)
        arrs = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None
                for x in (plev, tlay, tsfc, tlev)]
        self.lib.oracle_run_blocks(C.byref(self._m), int(lw), block, nlay, C.c_long(nblocks),
                                   _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
                                   n, names, values, kinds, nthreads, C.byref(chk))
        return chk.value

    def max_threads(self) -> int:
        return self.lib.oracle_max_threads()
