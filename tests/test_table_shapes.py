"""Table-shape sweep (BASELINE configs[2]): models built through ecckd_model_create from host
tables -- the Fortran shim's route -- with shapes the three shipped files do not have, GPU vs oracle.

* g-point counts 16 / 64 / 96 (LW) and 54 (SW) by slicing / tiling the shipped tables (rrtmgp-tol0.0161 is absent
  from the reference tree itself, .MISSING_LARGE_BLOBS);
* models that do NOT qualify for the table-walking kernels and must take generic_kernel with the reference's
  per-g clamp (:233-238): a table with negative entries, two look-up-table gases, tables that are not exact in fp32,
  more than 64 pressure levels;
* a call of >= 2^31 output elements (the quad kernel addresses outputs with 64-bit offsets / tensor maps).
"""
import copy

import numpy as np
import pytest
import torch

import ecckd_b200
from ecckd_b200 import synthetic
from conftest import atm_numpy, rel_err
from test_gpu_parity import run_gpu, run_oracle, check, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def oracle_for(m):
    from oracle.c_oracle import COracle
    return COracle(m)


def regrid_g(m, ngpt):
    """The same model on `ngpt` g-points: slices / tiles every per-g array (scaled so that tiles differ)."""
    m = copy.deepcopy(m)
    old = m.gpoint_fraction.shape[0]
    idx = np.arange(ngpt) % old
    scale = 1.0 + 0.25 * (np.arange(ngpt) // old)                 # exact in fp32 (multiples of 0.25)
    f4 = lambda a: a.astype(np.float32).astype(np.float64)
    for tab in m.absorption:
        tab.coefficient = f4(tab.coefficient[..., idx] * scale)
    m.gpoint_fraction = np.ascontiguousarray(m.gpoint_fraction[idx])
    if m.shortwave:
        m.solar_irradiance = f4(m.solar_irradiance[idx] * scale)
        m.rayleigh_molar_scattering_coeff = f4(m.rayleigh_molar_scattering_coeff[idx] * scale)
    else:
        m.planck_function = f4(m.planck_function[:, idx] * scale)
    return m


def share_composite(m):
    """composite aliases (o2, n2) must keep sharing ONE table object after deepcopy/regrid edits"""
    first = None
    for tab in m.absorption:
        if tab.composite_only:
            if first is None:
                first = tab
            else:
                tab.coefficient = first.coefficient
    return m


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("key,ngpt", [("lw_fsck", 16), ("lw_fsck", 64), ("lw_fsck", 96), ("lw_rrtmgp", 40), ("sw_wide", 54), ("sw_wide", 5)])
def test_gpoint_count_sweep(key, ngpt, dtype, dev, ck_models):
    m = share_composite(regrid_g(ck_models[key], ngpt))
    go = ecckd_b200.GasOpticsEcckd.from_components(m)
    assert go.get_ngpt() == ngpt and go.get_nband() == 0
    inp = atm_numpy(synthetic.make_atmosphere(68, 45))      # a multiple of 4: quad-eligible in fp32 too
    ref = run_oracle(oracle_for(m), m, *inp, dtype)
    for kern in (ecckd_b200.KERNEL_QUAD, ecckd_b200.KERNEL_COLWALK, ecckd_b200.KERNEL_GENERIC):
        check(run_gpu(go, *inp, dtype, dev, kern), ref, TOL[dtype], f"{key} ngpt={ngpt} kernel={kern}")


def regrid_planck(m, nplanck, t0, dt):
    """The same longwave model with a Planck table of `nplanck` rows starting at t0 with spacing dt (resampled linearly
    from the file's table, rounded to fp32 like the file's entries)."""
    m = copy.deepcopy(m)
    told, pold = m.temperature_planck, m.planck_function            # (n,), (n, ng)
    tnew = t0 + dt * np.arange(nplanck)
    pnew = np.stack([np.interp(tnew, told, pold[:, g]) for g in range(pold.shape[1])], axis=1)
    m.temperature_planck = tnew.astype(np.float64)
    m.planck_function = pnew.astype(np.float32).astype(np.float64)
    return m


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("nplanck,t0,dt", [(2, 200.0, 60.0), (33, 180.0, 4.0), (256, 150.0, 0.75), (231, 260.0, 1.0)],
                         ids=["2rows", "33rows", "256rows-max", "table-above-most-of-the-atmosphere"])
def test_planck_table_size_sweep(nplanck, t0, dt, dtype, dev, ck_models):
    """The block's shared-memory copy of the Planck table (pitch nplanck + 1, fp64 rows or fp32 pairs), extrapolation
    above the last row and proportional scaling below the first (:278-284) -- against the oracle, on all three kernels."""
    m = share_composite(regrid_planck(ck_models["lw_fsck"], nplanck, t0, dt))
    go = ecckd_b200.GasOpticsEcckd.from_components(m)
    inp = atm_numpy(synthetic.make_atmosphere(100, 61))
    ref = run_oracle(oracle_for(m), m, *inp, dtype)
    for kern in (ecckd_b200.KERNEL_QUAD, ecckd_b200.KERNEL_COLWALK, ecckd_b200.KERNEL_GENERIC):
        check(run_gpu(go, *inp, dtype, dev, kern), ref, TOL[dtype], f"nplanck={nplanck} kernel={kern}")


def _negative_entry(m):
    m = copy.deepcopy(m)
    tab = next(t for t, n in zip(m.absorption, m.gas) if n == "co2")
    tab.coefficient[0, 2, 20:30, :] *= -0.5                      # od < 0 for these cells: clamped per g-point (:233-238)
    return share_composite(m)


def _two_lut_gases(m):
    m = copy.deepcopy(m)
    tab = next(t for t, n in zip(m.absorption, m.gas) if n == "o3")
    k = tab.coefficient                                           # (1, nt, np, ng) -> a 3-entry look-up table in o3
    tab.coefficient = np.concatenate([k * 0.5, k, k * 1.5], axis=0).astype(np.float32).astype(np.float64)
    tab.mole_fraction = np.array([1e-8, 1e-7, 1e-6])
    tab.concentration_dependence_code = 2
    return share_composite(m)


def _not_fp32_exact(m):
    m = copy.deepcopy(m)
    for tab in m.absorption:
        tab.coefficient = tab.coefficient * (1.0 + 2.0 ** -40)
    return share_composite(m)


def _many_pressures(m):
    """70 pressure levels on the same uniform log grid (the walking kernels pad to 64)."""
    m = copy.deepcopy(m)
    np_old, np_new = m.log_pressure.shape[0], 70
    d = m.log_pressure[1] - m.log_pressure[0]
    m.log_pressure = m.log_pressure[0] + d * np.arange(np_new)
    idx = np.minimum(np.arange(np_new), np_old - 1)
    m.temperature = np.ascontiguousarray(m.temperature[:, idx])
    for tab in m.absorption:
        tab.coefficient = np.ascontiguousarray(tab.coefficient[:, :, idx, :])
    return share_composite(m)


@pytest.mark.parametrize("make", [_negative_entry, _two_lut_gases, _not_fp32_exact, _many_pressures],
                         ids=["negative-entry", "two-lut-gases", "not-fp32-exact", "70-pressures"])
@pytest.mark.parametrize("key", ["lw_fsck", "sw_wide"])
def test_non_qualifying_models_take_the_generic_kernel(key, make, dev, ck_models):
    m = make(ck_models[key])
    go = ecckd_b200.GasOpticsEcckd.from_components(m)
    atm = synthetic.make_atmosphere(64, 50)
    inp = list(atm_numpy(atm))
    if make is _two_lut_gases:                                    # put o3 inside its new table's range
        inp[4] = [(n, np.asfortranarray(np.clip(v, 2e-8, 9e-7)) if n == "o3" else v) for n, v in inp[4]]
    if make is _negative_entry:                                   # enough co2 for the negative cells to matter
        inp[4] = [(n, 2e-3 if n == "co2" else v) for n, v in inp[4]]
    ref = run_oracle(oracle_for(m), m, *inp)
    check(run_gpu(go, *inp, torch.float64, dev, ecckd_b200.KERNEL_AUTO), ref, 1e-12, f"{key} {make.__name__}")
    with pytest.raises(ecckd_b200.EcckdError):                    # the table-walking kernels must refuse, not mis-compute
        run_gpu(go, *inp, torch.float64, dev, ecckd_b200.KERNEL_WALK)
    if make is _not_fp32_exact:
        with pytest.raises(ecckd_b200.EcckdError):                # fp32 entry points need fp32-exact tables
            run_gpu(go, *inp, torch.float32, dev, ecckd_b200.KERNEL_AUTO)
    if make is _negative_entry:                                   # the clamp really fired somewhere
        unclamped = copy.deepcopy(m)
        for tab in unclamped.absorption:
            tab.coefficient = np.abs(tab.coefficient)
        assert rel_err(run_oracle(oracle_for(share_composite(unclamped)), unclamped, *inp)["tau"], ref["tau"]) > 1e-6


@pytest.mark.parametrize("key,dtype", [("sw_wide", torch.float64), ("lw_fsck", torch.float32)], ids=["sw-f64", "lw-f32"])
def test_call_with_2_to_the_31_output_elements(key, dtype, dev, ck_models):
    """ncol * nlay * ngpt >= 2^31: one device call, quad kernel (AUTO), every plane written, sampled columns bitwise
    equal to a small call on the same columns."""
    go = ecckd_b200.GasOpticsEcckd.load(__import__("conftest").FILES[key])
    ngpt = go.get_ngpt()
    nlay = 64
    ncol = -(-(2 ** 31) // (nlay * ngpt)) + 31
    ncol += (-ncol) % 4
    assert ncol * nlay * ngpt >= 2 ** 31
    free, _ = torch.cuda.mem_get_info()
    need = ncol * nlay * ngpt * (8 if dtype == torch.float64 else 4) * (3 if key == "sw_wide" else 4) * 1.1
    if free < need:
        pytest.skip(f"needs {need / 2**30:.0f} GiB of device memory")
    atm = synthetic.make_atmosphere(ncol, nlay, device=dev)
    if dtype == torch.float32:
        atm = atm.astype(dtype)
    lw = go.source_is_internal()
    npdt = np.float64 if dtype == torch.float64 else np.float32
    out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, ecckd_b200.PreparedGases(atm.gas_list(), ncol, nlay, npdt, dev),
                        tlev=atm.tlev if lw else None, kernel=ecckd_b200.KERNEL_QUAD)
    torch.cuda.synchronize()
    pick = torch.tensor([0, 1, 31, 32, ncol // 2, ncol - 33, ncol - 2, ncol - 1], device=dev)
    sub = lambda t: ecckd_b200.to_fortran(t[pick], dtype, dev)
    gl = [(n, sub(v) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
    small = go.gas_optics(None, sub(atm.plev), sub(atm.tlay), atm.tsfc[pick].contiguous() if lw else None, gl,
                          tlev=sub(atm.tlev) if lw else None)
    for k, v in small.items():
        assert torch.equal(out[k][pick], v), f"{key} {k}: big call differs from the small call on the same columns"
    assert torch.isfinite(out["tau"][:, :, ngpt - 1]).all() and (out["tau"][:, nlay - 1, :] > 0).all()
    del out
    torch.cuda.empty_cache()
