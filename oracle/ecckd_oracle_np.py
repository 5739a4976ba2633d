"""numpy twin of the oracle: an independent restatement of the reference's gas optics.

TEST INFRASTRUCTURE ONLY -- never imported by the product (rte-ecckd_b200/).
PARITY UNPINNED: the reference ships no golden vectors and cannot be built in
this image (no Fortran compiler, no netCDF, no rte-rrtmgp); this module and the
C restatement (oracle/ecckd_oracle.c) are written separately from
/root/reference/src/gas_optics_ecckd.f90 and must agree with each other.

Arrays are indexed like the Fortran: plev[col, lev], tau[col, lay, g].
Every arithmetic expression keeps the reference's grouping; numpy float64
ufuncs never contract a*b+c into an FMA, so this is plain IEEE arithmetic.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .ckd_model import CkdModel, LOOK_UP_TABLE, RELATIVE_LINEAR, LINEAR, NONE_

f4, f8 = np.float32, np.float64

# gas_optics_ecckd.f90:51-53 -- default-real literals assigned to real(wp)
GRAVITY = f8(f4(9.80665))
DRY_AIR_MOLAR_MASS = f8(f4(28.970))
PI = f8(f4(3.14159265359))
MILLI = f8(f4(0.001))                      # the literal 0.001 at :107 and :314

_libm_log = np.frompyfunc(math.log, 1, 1)


def log(x: np.ndarray) -> np.ndarray:
    """libm log, elementwise (np.log may differ from libm in the last ulp)."""
    return np.asarray(_libm_log(np.asarray(x, dtype=f8)), dtype=f8)


def global_weight() -> np.float64:
    """gas_optics_ecckd.f90:107 -- 1./(gravity*0.001*dry_air_molar_mass), left to right."""
    return f8(1.0) / ((GRAVITY * MILLI) * DRY_AIR_MOLAR_MASS)


GasConcs = Sequence[Tuple[str, Union[float, np.ndarray]]]


def get_vmr(value, ncol: int, nlay: int, name: str) -> np.ndarray:
    """ty_gas_concs%get_vmr broadcasting (rte-rrtmgp, not in tree; call site :351):
    scalar -> everywhere, (nlay,) -> every column, (ncol, nlay) -> as is."""
    v = np.asarray(value, dtype=f8)
    if v.ndim == 0:
        return np.full((ncol, nlay), v, dtype=f8)
    if v.ndim == 1:
        if v.shape[0] != nlay:
            raise ValueError(f"ty_gas_concs%get_vmr; gas {name} array is wrong size (nlay)")
        return np.broadcast_to(v[None, :], (ncol, nlay)).astype(f8)
    if v.shape != (ncol, nlay):
        raise ValueError(f"ty_gas_concs%get_vmr; gas {name} array is wrong size")
    return v.astype(f8)


def interpolation_indices(model: CkdModel, plev: np.ndarray, tlay: np.ndarray):
    """gas_optics_ecckd.f90:104-106,120-140.  Returns 1-based ip0, it0 and weights."""
    logp = model.log_pressure
    npress = logp.shape[0]
    ntemp = model.temperature.shape[0]
    log_p_0 = logp[0]
    d_log_p = logp[1] - logp[0]
    dt = model.temperature[1, 0] - model.temperature[0, 0]   # temperature(1,2)-temperature(1,1)
    nlay = tlay.shape[1]
    log_pressure = log(f8(0.5) * (plev[:, 1:nlay + 1] + plev[:, 0:nlay]))
    x = (log_pressure - log_p_0) / d_log_p
    x = f8(1.0) + np.maximum(f8(0.0), np.minimum(x, f8(npress) - f8(1.0001)))
    ip0 = x.astype(np.int64)                                  # int(): truncation
    pw1 = x - ip0
    pw0 = f8(1.0) - pw1
    tcol = model.temperature[0, :]                            # temperature(:,1)
    t0 = pw0 * tcol[ip0 - 1] + pw1 * tcol[ip0]
    y = (tlay - t0) / dt
    y = f8(1.0) + np.maximum(f8(0.0), np.minimum(y, f8(ntemp) - f8(1.0001)))
    it0 = y.astype(np.int64)
    tw1 = y - it0
    tw0 = f8(1.0) - tw1
    return ip0, pw0, pw1, it0, tw0, tw1


def vmr_indices(mole_fraction: np.ndarray, layer_vmr: np.ndarray):
    """gas_optics_ecckd.f90:153-163.  Returns 1-based iv0 and weights."""
    nv = mole_fraction.shape[0]
    log_vmr = log(np.maximum(layer_vmr, mole_fraction[0]))
    d_log_vmr = f8(math.log(mole_fraction[1] / mole_fraction[0]))
    z = (log_vmr - f8(math.log(mole_fraction[0]))) / d_log_vmr
    z = f8(1.0) + np.maximum(f8(0.0), np.minimum(z, f8(nv) - f8(1.001)))
    iv0 = z.astype(np.int64)
    vw1 = z - iv0
    vw0 = f8(1.0) - vw1
    return iv0, vw0, vw1


def calculate_optical_depth(model: CkdModel, gas: int, plev, tlay, layer_vmr) -> np.ndarray:
    """gas_optics_ecckd.f90:64-241 with logarithmic_interpolation = .false. (:368-369)."""
    tab = model.absorption[gas]
    K = tab.coefficient                                       # [v, t, p, g]
    nlay = tlay.shape[1]
    ip0, pw0, pw1, it0, tw0, tw1 = interpolation_indices(model, plev, tlay)
    simple_weight = global_weight() * (plev[:, 1:nlay + 1] - plev[:, 0:nlay])
    code = tab.concentration_dependence_code
    if code == RELATIVE_LINEAR:
        weight = simple_weight * (layer_vmr - f8(tab.reference_mole_fraction))
    else:
        weight = simple_weight * layer_vmr
    ip, it = ip0 - 1, it0 - 1                                 # 0-based
    e = lambda a: a[..., None]

    def bilinear(v):
        return (e(tw0) * (e(pw0) * K[v, it, ip, :] + e(pw1) * K[v, it, ip + 1, :]) +
                e(tw1) * (e(pw0) * K[v, it + 1, ip, :] + e(pw1) * K[v, it + 1, ip + 1, :]))

    if code == LOOK_UP_TABLE:
        iv0, vw0, vw1 = vmr_indices(tab.mole_fraction, layer_vmr)
        iv = iv0 - 1
        od = e(weight) * (e(vw0) * bilinear(iv) + e(vw1) * bilinear(iv + 1))
    elif code in (LINEAR, RELATIVE_LINEAR):
        od = e(weight) * bilinear(0)
    else:
        od = e(simple_weight) * bilinear(0)
    od = np.where(od < 0.0, f8(0.0), od)                      # :233-238
    return od


def calculate_planck_function(model: CkdModel, temperature: np.ndarray) -> np.ndarray:
    """gas_optics_ecckd.f90:245-289.  temperature[col, lev] -> planck[col, lev, g]."""
    tp = model.temperature_planck
    P = model.planck_function                                 # [itp, g]
    dt = tp[1] - tp[0]
    t0 = tp[0]
    x = (temperature - t0) / dt
    ge = x >= 0
    xi = f8(1.0) + x
    with np.errstate(invalid="ignore"):
        it0 = np.minimum(np.where(ge, xi, 1.0).astype(np.int64), tp.shape[0] - 1)
    w1 = xi - it0
    w0 = f8(1.0) - w1
    e = lambda a: a[..., None]
    hi = e(w0) * P[it0 - 1, :] + e(w1) * P[it0, :]
    lo = e(temperature / t0) * P[0, :]
    return np.where(e(ge), hi, lo) / PI


def calculate_rayleigh_optical_depth(model: CkdModel, plev: np.ndarray) -> np.ndarray:
    """gas_optics_ecckd.f90:293-319"""
    nlay = plev.shape[1] - 1
    moles = (plev[:, 1:nlay + 1] - plev[:, 0:nlay]) * (f8(1.0) / ((GRAVITY * MILLI) * DRY_AIR_MOLAR_MASS))
    return moles[..., None] * model.rayleigh_molar_scattering_coeff[None, None, :]


def gas_optical_depth(model: CkdModel, plev, tlay, gas_concs: GasConcs) -> np.ndarray:
    """gas_optics_ecckd.f90:323-376: caller's gas order, unknown gases skipped,
    composite table evaluated once."""
    ncol, nlay = tlay.shape
    tau = np.zeros((ncol, nlay, model.ngpt), dtype=f8)
    first_calc = True
    for name, value in gas_concs:
        name = name.strip().lower()
        try:
            i = model.gas[:model.num_gases].index(name)
        except ValueError:
            continue
        layer_vmr = get_vmr(value, ncol, nlay, name)
        if model.absorption[i].composite_only and not first_calc:
            continue
        tau = tau + calculate_optical_depth(model, i, plev, tlay, layer_vmr)
        if model.absorption[i].composite_only:
            first_calc = False
    return tau


def gas_optics_int(model: CkdModel, plev, tlay, tsfc, gas_concs: GasConcs, tlev=None) -> Dict[str, np.ndarray]:
    """gas_optics_ecckd.f90:381-426 (LW).  Raises like the reference's error string."""
    plev, tlay, tsfc = (np.asarray(a, dtype=f8) for a in (plev, tlay, tsfc))
    nlay = tlay.shape[1]
    out = {"tau": gas_optical_depth(model, plev, tlay, gas_concs)}
    out["lay_source"] = calculate_planck_function(model, tlay)
    out["sfc_source"] = calculate_planck_function(model, tsfc[:, None])[:, 0, :]
    if tlev is None:
        raise ValueError("tlev is required for ecckd")
    buf = calculate_planck_function(model, np.asarray(tlev, dtype=f8))
    out["lev_source_inc"] = buf[:, 1:, :]
    out["lev_source_dec"] = buf[:, :nlay, :]
    return out


def gas_optics_ext(model: CkdModel, plev, tlay, gas_concs: GasConcs) -> Dict[str, np.ndarray]:
    """gas_optics_ecckd.f90:431-473 (SW, two-stream)."""
    plev, tlay = (np.asarray(a, dtype=f8) for a in (plev, tlay))
    tau = gas_optical_depth(model, plev, tlay, gas_concs)
    od_ray = calculate_rayleigh_optical_depth(model, plev)
    tau = tau + od_ray
    with np.errstate(invalid="ignore", divide="ignore"):
        ssa = od_ray / tau
    ncol = tlay.shape[0]
    toa = np.broadcast_to(model.solar_irradiance[None, :], (ncol, model.ngpt)).copy()
    return {"tau": tau, "ssa": ssa, "g": np.zeros_like(tau), "toa_src": toa}
