"""Front end of oracle/_ref/ref_driver: the unmodified reference Fortran on flat binary files.

TEST INFRASTRUCTURE ONLY.  `available()` tells whether the driver exists (or can be built: needs
a Fortran compiler and /root/reference); `run()` dumps a CkdModel and an atmosphere, executes the
driver and reads back what the reference returned, in the oracle's dictionary layout.
"""
from __future__ import annotations

import os
import shutil
import struct
import subprocess
import tempfile
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER = os.path.join(os.path.dirname(HERE), "_ref", "ref_driver")
REFERENCE = os.environ.get("ECCKD_REFERENCE", "/root/reference")


def fortran_compiler() -> Optional[str]:
    for fc in (os.environ.get("FC"), "gfortran", "flang", "flang-new", "ifx", "ifort", "nvfortran"):
        if fc and shutil.which(fc):
            return fc
    return None


def build() -> bool:
    """Build the driver if a compiler and the reference sources are present; True when it exists afterwards."""
    src = os.path.join(REFERENCE, "src", "gas_optics_ecckd.f90")
    fc = fortran_compiler()
    if fc and os.path.exists(src):
        subprocess.run(["make", "-C", HERE, f"FC={fc}", f"REFERENCE={REFERENCE}"], check=False)
    return os.path.exists(DRIVER)


def available() -> bool:
    return os.path.exists(DRIVER) or build()


def _f(a) -> bytes:
    return np.asfortranarray(np.asarray(a, dtype=np.float64)).tobytes(order="F")


def dump_tables(model, path: str) -> None:
    """Layout read by ref_driver.f90 (`read(u) ...` in this order); `model` is an oracle.ckd_model.CkdModel."""
    ngas = len(model.gas)
    ngpt = int(model.ngpt)
    npress, ntemp = model.log_pressure.shape[0], model.temperature.shape[0]
    nplanck = 0 if model.shortwave else model.temperature_planck.shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("=7i", 20260101, int(model.shortwave), ngpt, npress, ntemp, nplanck, ngas))
        for name, tab in zip(model.gas, model.absorption):
            nv = 0 if tab.mole_fraction is None else int(tab.mole_fraction.shape[0])
            f.write(name.encode().ljust(32)[:32])
            f.write(struct.pack("=3i", int(tab.composite_only), int(tab.concentration_dependence_code), nv))
            f.write(struct.pack("=d", float(tab.reference_mole_fraction)))
            if nv:
                f.write(_f(tab.mole_fraction))
            # CkdModel keeps K[v, t, p, g] (C order) == Fortran coefficient(g, p, t, v) in memory
            f.write(np.ascontiguousarray(tab.coefficient, dtype=np.float64).tobytes())
        f.write(_f(model.log_pressure))
        f.write(np.ascontiguousarray(model.temperature, dtype=np.float64).tobytes())      # [t, p] C == (p, t) Fortran
        if model.shortwave:
            f.write(_f(model.solar_irradiance))
            f.write(_f(model.rayleigh_molar_scattering_coeff))
            f.write(struct.pack("=d", float(model.total_solar_irradiance)))
        else:
            f.write(_f(model.temperature_planck))
            f.write(np.ascontiguousarray(model.planck_function, dtype=np.float64).tobytes())  # [T, g] C == (g, T) Fortran


GasList = Sequence[Tuple[str, Union[float, np.ndarray]]]


def dump_inputs(path: str, plev, tlay, tlev, tsfc, gases: GasList) -> None:
    ncol, nlay = tlay.shape
    with open(path, "wb") as f:
        f.write(struct.pack("=5i", 20260102, ncol, nlay, int(tlev is not None), len(gases)))
        f.write(_f(plev)); f.write(_f(tlay))
        f.write(_f(tlev if tlev is not None else np.zeros((ncol, nlay + 1))))
        f.write(_f(tsfc if tsfc is not None else np.zeros(ncol)))
        for name, _ in gases:
            f.write(name.encode().ljust(32)[:32])
        for _, v in gases:
            a = np.asarray(v, dtype=np.float64)
            kind = {0: 0, 1: 1, 2: 2}[a.ndim]
            f.write(struct.pack("=i", kind))
            f.write(struct.pack("=d", float(a)) if kind == 0 else _f(a))


def run(model, plev, tlay, tsfc, gases: GasList, tlev=None) -> Tuple[Dict[str, np.ndarray], str]:
    """Returns (outputs, error_msg) exactly like oracle.c_oracle.COracle.gas_optics_int / gas_optics_ext."""
    if not available():
        raise RuntimeError("oracle/_ref/ref_driver is not built (no Fortran compiler or no reference sources)")
    ncol, nlay = tlay.shape
    ngpt = int(model.ngpt)
    with tempfile.TemporaryDirectory() as d:
        ft, fi, fo = (os.path.join(d, n) for n in ("tables.bin", "inputs.bin", "outputs.bin"))
        dump_tables(model, ft)
        dump_inputs(fi, plev, tlay, tlev, tsfc, gases)
        subprocess.run([DRIVER, ft, fi, fo], check=True)
        raw = open(fo, "rb").read()
    magic, nerr = struct.unpack_from("=2i", raw, 0)
    assert magic == 20260103
    err = raw[8:8 + 128].decode()[:nerr]
    off = 8 + 128

    def take(shape):
        nonlocal off
        n = int(np.prod(shape))
        a = np.frombuffer(raw, dtype=np.float64, count=n, offset=off).reshape(shape, order="F")
        off += 8 * n
        return np.asfortranarray(a)
    names = ("tau", "ssa", "g") if model.shortwave else ("tau", "lay_source", "lev_source_inc", "lev_source_dec")
    out = {k: take((ncol, nlay, ngpt)) for k in names}
    out["toa_src" if model.shortwave else "sfc_source"] = take((ncol, ngpt))
    return out, err
