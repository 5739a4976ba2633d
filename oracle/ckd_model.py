"""CPU restatement of the reference's ck-definition loader.

TEST INFRASTRUCTURE ONLY.  Follows
  /root/reference/example/rfmip-rad-irf/mo_load_coefficients.F90:19-146 (load_and_init)
  /root/reference/example/rfmip-rad-irf/mo_load_coefficients.F90:149-203 (read_gas_input_data)
  /root/reference/example/rfmip-rad-irf/mo_load_coefficients.F90:244-293 (tokenize)
and yields the same components the reference stores in ty_gas_optics_ecckd
(/root/reference/src/gas_optics_ecckd.f90:23-36), as float64 numpy arrays whose
C-order shape is the reverse of the Fortran shape (so `coefficient[v, t, p, g]`
here is `coefficient(g, p, t, v)` there).  netCDF f4 -> wp promotion is exact
(mo_simple_netcdf.F90:60-142 reads into real(wp) buffers).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import netcdf3

# concentration-dependence codes, gas_optics_ecckd.f90:54-57
NONE_, LINEAR, LOOK_UP_TABLE, RELATIVE_LINEAR = 0, 1, 2, 3


@dataclass
class AbsorptionTable:
    """gas_optics_ecckd.f90:13-19"""
    coefficient: np.ndarray            # (nv, nt, np, ng) float64
    composite_only: bool
    concentration_dependence_code: int
    mole_fraction: Optional[np.ndarray] = None   # (nv,) for look_up_table
    reference_mole_fraction: float = 0.0


@dataclass
class CkdModel:
    """gas_optics_ecckd.f90:23-36 (public components) + what %init receives."""
    shortwave: bool
    gas: List[str]
    absorption: List[AbsorptionTable]
    num_gases: int
    num_composite_gases: int
    gpoint_fraction: np.ndarray        # (ng, nwav)  [Fortran (nwav, ng)]
    log_pressure: np.ndarray           # (np,)
    temperature: np.ndarray            # (nt, np)    [Fortran (np, nt)]
    band_lims_wvn: np.ndarray          # (nband, 2)
    band2gpt: np.ndarray               # (nband, 2), 1-based inclusive
    planck_function: Optional[np.ndarray] = None      # (nplanck, ng)
    temperature_planck: Optional[np.ndarray] = None   # (nplanck,)
    rayleigh_molar_scattering_coeff: Optional[np.ndarray] = None  # (ng,)
    solar_irradiance: Optional[np.ndarray] = None     # (ng,)
    total_solar_irradiance: float = 0.0
    source_file: str = field(default="", repr=False)

    @property
    def ngpt(self) -> int:
        return self.gpoint_fraction.shape[0]   # size(gpoint_fraction, 2) in Fortran


def tokenize(buffer: str, max_tokens: int = 16, token_len: int = 32) -> List[str]:
    """mo_load_coefficients.F90:244-293, including its quirk that a trailing
    one-character token is never emitted (found_token is only set at i == n)."""
    n = len(buffer.rstrip(" "))
    tokens: List[str] = []
    found, start = False, 0
    for i in range(1, n + 1):          # 1-based like the Fortran
        ch = buffer[i - 1]
        if found:
            if i == n and ch != " ":
                if len(tokens) + 1 > max_tokens:
                    raise RuntimeError("tokenize: more tokens than can fit in array.")
                if i - start + 1 > token_len:
                    raise RuntimeError("tokenize: token is cut off.")
                tokens.append(buffer[start - 1:i])
                break
            elif ch == " ":
                if len(tokens) + 1 > max_tokens:
                    raise RuntimeError("tokenize: more tokens than can fit in array.")
                if i - start > token_len:
                    raise RuntimeError("tokenize: token is cut off.")
                tokens.append(buffer[start - 1:i - 1])
                found = False
        else:
            if ch != " ":
                found, start = True, i
    return tokens


def _f8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _read_gas_input_data(ds: netcdf3.Dataset, gas_name: str) -> AbsorptionTable:
    """mo_load_coefficients.F90:149-203"""
    mf_name = f"{gas_name}_mole_fraction"
    if mf_name in ds and ds.variables[mf_name].data.ndim == 1:
        coef = _f8(ds[f"{gas_name}_molar_absorption_coeff"])
        if coef.ndim != 4:
            raise RuntimeError(f"LUT gas {gas_name}: coefficient not 4d")
        return AbsorptionTable(coef, False, LOOK_UP_TABLE, mole_fraction=_f8(ds[mf_name]))
    n = int(ds[f"{gas_name}_conc_dependence_code"])
    ref = 0.0
    if n == 0:
        code = NONE_
    elif n == 1:
        code = LINEAR
    elif n == 3:
        code = RELATIVE_LINEAR
        ref = float(np.float64(ds[f"{gas_name}_reference_mole_fraction"]))
    else:
        raise RuntimeError(f"load_and_init_ecckd: bad concentration code for {gas_name}")
    coef = ds[f"{gas_name}_molar_absorption_coeff"]
    if coef.ndim != 3:
        raise RuntimeError(f"load_and_init_ecckd: absorption coefficient not 3d for {gas_name}")
    return AbsorptionTable(_f8(coef)[None, ...], False, code, reference_mole_fraction=ref)


def load_and_init(filename: str) -> CkdModel:
    """mo_load_coefficients.F90:19-146 (the unused `available_gases` argument is dropped)."""
    ds = netcdf3.Dataset(filename)
    # :46-49, log taken in wp.  math.log is libm's log (what gfortran calls);
    # np.log may use a SIMD kernel that differs from libm in the last ulp.
    log_pressure = np.array([math.log(p) for p in _f8(ds["pressure"])], dtype=np.float64)
    temperature = _f8(ds["temperature"])                  # :51-53
    nband = ds["wavenumber1_band"].shape[0]
    band_lims = np.stack([_f8(ds["wavenumber1_band"]), _f8(ds["wavenumber2_band"])], axis=1)
    gpt2band = ds["band_number"].astype(np.int64) + 1    # :62-63
    ng = gpt2band.shape[0]
    band2gpt = np.zeros((nband, 2), dtype=np.int64)
    band2gpt[0, 0] = 1
    band2gpt[nband - 1, 1] = ng
    band = 1
    for i in range(2, ng + 1):                           # :67-73
        if gpt2band[i - 1] > band:
            band2gpt[band - 1, 1] = i - 1
            band += 1
            band2gpt[band - 1, 0] = i
    gpoint_fraction = _f8(ds["gpoint_fraction"])
    shortwave = "solar_irradiance" in ds                 # :84
    kw = {}
    if shortwave:
        sol = _f8(ds["solar_irradiance"])
        kw = dict(solar_irradiance=sol, total_solar_irradiance=float(np.sum(sol)),
                  rayleigh_molar_scattering_coeff=_f8(ds["rayleigh_molar_scattering_coeff"]))
    else:
        kw = dict(temperature_planck=_f8(ds["temperature_planck"]),
                  planck_function=_f8(ds["planck_function"]))

    tokens = tokenize(ds.attrs["constituent_id"])        # :104-107
    composite_tokens: List[str] = []
    uses_composite = False
    for t in tokens:                                     # :109-117
        if t == "composite":
            uses_composite = True
            composite_tokens = tokenize(ds.attrs["composite_constituent_id"])
            break
    gas: List[str] = []
    absorption: List[AbsorptionTable] = []
    for t in tokens:                                     # :119-126
        if t != "composite":
            gas.append(t)
            absorption.append(_read_gas_input_data(ds, t))
    if uses_composite:                                   # :127-143
        for c in composite_tokens:
            if c not in tokens:
                gas.append(c)
                tab = _read_gas_input_data(ds, "composite")
                tab.composite_only = True
                absorption.append(tab)
    if len(gas) > 16:
        raise RuntimeError("more than 16 gases (absorption(16), gas_optics_ecckd.f90:24)")
    return CkdModel(shortwave=shortwave, gas=gas, absorption=absorption, num_gases=len(gas),
                    num_composite_gases=len(composite_tokens), gpoint_fraction=gpoint_fraction,
                    log_pressure=log_pressure, temperature=temperature,
                    band_lims_wvn=band_lims, band2gpt=band2gpt, source_file=filename, **kw)
