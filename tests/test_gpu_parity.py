"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on identical inputs.

Bars (BASELINE.json north_star): interpolation indices bit-exact; fp64 tau/ssa/g/sources within 1e-12
relative; the fp32 variant within 1e-5 relative (vs the fp64 oracle on the fp32-rounded inputs).
"""
import os

import numpy as np
import pytest
import torch

import ecckd_b200
from ecckd_b200 import synthetic
from conftest import FILES, atm_numpy, rel_err

pytestmark = pytest.mark.gpu

TOL = {torch.float64: 1e-12, torch.float32: 1e-5}
# "walk" = the product path (quad kernel when the call admits TMA tensor maps, else the one-column-per-warp kernel);
# "colwalk" forces the latter (the fallback for odd column counts) also where the quad kernel would run
KERNELS = {"walk": ecckd_b200.KERNEL_WALK, "colwalk": ecckd_b200.KERNEL_COLWALK, "generic": ecckd_b200.KERNEL_GENERIC}
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def gpu_models(dev):
    return {k: ecckd_b200.GasOpticsEcckd.load(v) for k, v in FILES.items()}


def to_dev(a, dtype, dev):
    return ecckd_b200.to_fortran(torch.as_tensor(np.ascontiguousarray(a)), dtype, dev)


def run_gpu(go, plev, tlay, tlev, tsfc, gases, dtype, dev, kernel=ecckd_b200.KERNEL_AUTO, **kw):
    """numpy inputs -> device call -> numpy outputs"""
    lw = go.source_is_internal()
    gl = [(n, to_dev(v, dtype, dev) if isinstance(v, np.ndarray) else v) for n, v in gases]
    out = go.gas_optics(None, to_dev(plev, dtype, dev), to_dev(tlay, dtype, dev),
                        torch.as_tensor(tsfc).to(dev, dtype) if lw else None, gl,
                        tlev=to_dev(tlev, dtype, dev) if (lw and tlev is not None) else None, kernel=kernel, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def run_oracle(co, m, plev, tlay, tlev, tsfc, gases, dtype=torch.float64):
    if dtype == torch.float32:   # the fp32 variant sees fp32-rounded inputs
        r = lambda a: np.asfortranarray(np.asarray(a, dtype=np.float32).astype(np.float64))
        plev, tlay, tlev, tsfc = r(plev), r(tlay), r(tlev), r(tsfc)
        gases = [(n, r(v) if isinstance(v, np.ndarray) else float(np.float32(v))) for n, v in gases]
    out, err = co.gas_optics_ext(plev, tlay, gases) if m.shortwave else co.gas_optics_int(plev, tlay, tsfc, gases, tlev)
    assert err == ""
    return out


def check(out, ref, tol, what=""):
    for k, v in ref.items():
        e = rel_err(out[k], v)
        assert e <= tol, f"{what} {k}: rel err {e:.3e} > {tol}"


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("key", list(FILES))
def test_gcm_block_parity(key, kernel, dtype, dev, gpu_models, ck_models, c_oracles):
    """137-layer GCM-like columns (config 4 at oracle-sized ncol; 70 columns = 4 full tiles + a ragged one)."""
    inp = atm_numpy(synthetic.make_atmosphere(70, 137))
    out = run_gpu(gpu_models[key], *inp, dtype, dev, KERNELS[kernel])
    check(out, run_oracle(c_oracles[key], ck_models[key], *inp, dtype), TOL[dtype], f"{key}/{kernel}")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("key", list(FILES))
def test_rfmip_shaped_parity(key, dtype, dev, gpu_models, ck_models, c_oracles):
    """Configs 1-3: 18 experiments x 100 columns x 60 layers, scalar gases straddling the ch4/n2o references."""
    for e, atm in enumerate(synthetic.make_rfmip_like()):
        inp = atm_numpy(atm)
        out = run_gpu(gpu_models[key], *inp, dtype, dev)
        check(out, run_oracle(c_oracles[key], ck_models[key], *inp, dtype), TOL[dtype], f"{key}/exp{e}")


@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("key", list(FILES))
def test_committed_golden_fixtures(key, kernel, dev, gpu_models):
    """GPU vs tests/golden (hand-built columns hitting every clamp / branch)."""
    z = np.load(os.path.join(GOLD, "inputs.npz"))
    f = np.asfortranarray
    gases = [("co2", 420e-6), ("ch4", 1.0e-6), ("n2o", 4.0e-7), ("o2", 0.209), ("cfc11", 2.3e-10),
             ("cfc12", 5.2e-10), ("h2o", f(z["h2o"])), ("o3", f(z["o3"])), ("no2", 0.0)]
    out = run_gpu(gpu_models[key], f(z["plev"]), f(z["tlay"]), f(z["tlev"]), z["tsfc"], gases, torch.float64, dev, KERNELS[kernel])
    gold = np.load(os.path.join(GOLD, f"{key}.npz"))
    check(out, {k: gold[k] for k in out}, 1e-12, f"golden {key}/{kernel}")


@pytest.mark.parametrize("key", list(FILES))
def test_indices_are_bit_exact(key, dev, gpu_models, c_oracles):
    """ip0 / it0 / iv0 as the kernels compute them == the oracle's, exactly; weights to the last bits
    (the only non-IEEE step is log(): CUDA's vs libm's, both < 1 ulp)."""
    atms = [synthetic.make_atmosphere(257, 137)] + synthetic.make_rfmip_like(nexp=3)
    z = np.load(os.path.join(GOLD, "inputs.npz"))
    cases = [atm_numpy(a) for a in atms]
    cases.append((np.asfortranarray(z["plev"]), np.asfortranarray(z["tlay"]), None, None, [("h2o", np.asfortranarray(z["h2o"]))]))
    for plev, tlay, _, _, gl in cases:
        h2o = dict(gl)["h2o"]
        ref = c_oracles[key].indices("h2o", plev, tlay, h2o)
        got = gpu_models[key].interpolation_indices(to_dev(plev, torch.float64, dev), to_dev(tlay, torch.float64, dev),
                                                    "h2o", to_dev(h2o, torch.float64, dev))
        got = {k: v.cpu().numpy() for k, v in got.items()}
        for k in ("ip0", "it0", "iv0"):
            bad = np.argwhere(got[k] != ref[k])
            msg = ""
            if bad.size:   # would only happen when the index variable sits within an ulp of an integer
                i, j = bad[0]
                msg = f"{k} differs at {tuple(bad[0])}: gpu {got[k][i, j]} oracle {ref[k][i, j]} weights {got[k.replace('i', '', 1)[0] + 'w1'][i, j] if False else ''}"
            assert not bad.size, msg
        for k in ("pw1", "tw1", "vw1"):
            assert np.max(np.abs(got[k] - ref[k])) <= 2e-13, k


def test_constant_divisor_division_is_the_ieee_quotient(dev, ck_models):
    """The index arithmetic divides by per-model constants as q0 = x*RN(1/c), q = fma(fma(-c, q0, x), RN(1/c), q0).
    That must be the correctly rounded quotient, bit for bit, for every divisor the three ck files produce (and a few
    awkward ones): 2^22 dividends each -- random over the ranges the algorithm sees, plus zeros, infinities, NaN and
    neighbours of integer quotients (where an index would flip)."""
    import ctypes as C
    from ecckd_b200 import _abi
    lib = _abi.load_library()
    g = torch.Generator(device=dev).manual_seed(7)
    divisors = {3.0, 1.0 / 3.0, 7.0, 1e-3, 123456.789, float(np.float64(np.float32(3.14159265359)))}
    for m in ck_models.values():
        divisors.add(float(m.log_pressure[1] - m.log_pressure[0]))
        divisors.add(float(m.temperature[1, 0] - m.temperature[0, 0]))
        if not m.shortwave:
            divisors.add(float(m.temperature_planck[1] - m.temperature_planck[0])); divisors.add(float(m.temperature_planck[0]))
        for tab in m.absorption:
            if tab.mole_fraction is not None:
                divisors.add(float(np.log(tab.mole_fraction[1] / tab.mole_fraction[0])))
    n = 1 << 22
    for c in sorted(divisors):
        x = torch.cat([
            (torch.rand(n // 4, generator=g, device=dev, dtype=torch.float64) - 0.5) * 60.0,          # log-pressure differences
            (torch.rand(n // 4, generator=g, device=dev, dtype=torch.float64) - 0.3) * 400.0,         # temperature differences
            torch.exp((torch.rand(n // 4, generator=g, device=dev, dtype=torch.float64) - 0.5) * 200.0),
            # quotients next to integers: x = c * k * (1 + j ulp)
            c * torch.randint(1, 64, (n // 4,), generator=g, device=dev).double()
              * (1.0 + (torch.randint(-3, 4, (n // 4,), generator=g, device=dev).double() * 2.220446049250313e-16)),
            torch.tensor([0.0, -0.0, float("inf"), float("-inf"), float("nan"), 1e-310, 1e308, -1e308], device=dev, dtype=torch.float64)])
        bad = C.c_uint64(123)
        err = C.create_string_buffer(_abi.ERRLEN)
        rc = lib.ecckd_debug_div_const(x.numel(), x.data_ptr(), c, C.byref(bad), err)
        assert rc == _abi.OK, err.value
        assert bad.value == 0, f"divisor {c!r}: {bad.value} of {x.numel()} quotients differ from the IEEE division"


def test_planck_indices_bit_exact(dev, gpu_models, c_oracles):
    T = np.concatenate([np.linspace(90, 420, 4001), [120.0, 119.99999999, 350.0, 350.00000001, 121.0]])
    it_ref, w_ref = c_oracles["lw_fsck"].planck_indices(T)
    it, w = gpu_models["lw_fsck"].planck_indices(torch.as_tensor(T).to(dev))
    assert np.array_equal(it.cpu().numpy(), it_ref[:, 0])
    assert np.array_equal(w.cpu().numpy(), w_ref[:, 0])        # pure IEEE arithmetic: identical


# ---- shapes: empty, single, ragged ----------------------------------------------------------------
@pytest.mark.parametrize("kernel", list(KERNELS))
@pytest.mark.parametrize("ncol,nlay", [(1, 1), (1, 137), (17, 33), (16, 32), (15, 31), (33, 64), (3, 65), (2, 3), (34, 20), (64, 137), (30, 5)])
def test_ragged_shapes(ncol, nlay, kernel, dev, gpu_models, ck_models, c_oracles):
    for key in ("lw_rrtmgp", "sw_wide"):      # 36 g-points = two g-groups, 27 = a partial one
        inp = atm_numpy(synthetic.make_atmosphere(ncol, nlay, col0=1000))
        out = run_gpu(gpu_models[key], *inp, torch.float64, dev, KERNELS[kernel])
        check(out, run_oracle(c_oracles[key], ck_models[key], *inp), 1e-12, f"{key} {ncol}x{nlay}")


def test_empty_call_is_a_no_op(dev, gpu_models):
    go = gpu_models["lw_fsck"]
    z2 = lambda n: ecckd_b200.fortran_empty((0, n), torch.float64, dev)
    out = go.gas_optics(None, z2(11), z2(10), torch.empty(0, dtype=torch.float64, device=dev), [("co2", 4e-4)], tlev=z2(11))
    assert out["tau"].shape == (0, 10, 32)


# ---- gas handling (gas_optical_depth :323-376) ----------------------------------------------------
def _small():
    return atm_numpy(synthetic.make_atmosphere(40, 60, col0=77))


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_gas_list_semantics(kernel, dev, gpu_models, ck_models, c_oracles):
    plev, tlay, tlev, tsfc, gases = _small()
    g = dict(gases)
    variants = {
        "unknown gases skipped": gases + [("sf6", 1e-9), ("co", 1e-7)],
        "reordered": list(reversed(gases)),
        "no o3": [x for x in gases if x[0] != "o3"],
        "composite via n2 only": [x for x in gases if x[0] != "o2"] + [("n2", 0.781)],
        "composite twice (o2 then n2)": gases + [("n2", 0.781)],
        "no composite": [x for x in gases if x[0] != "o2"],
        "nothing": [],
        "only h2o": [("h2o", g["h2o"])],
        "ch4/n2o below reference": [(n, (1e-7 if n in ("ch4", "n2o") else v)) for n, v in gases],
        "profiles": [(n, (np.linspace(1, 2, tlay.shape[1]) * v if n in ("co2", "cfc11") else v)) for n, v in gases],
        "all fields": [(n, v if isinstance(v, np.ndarray) else np.asfortranarray(v * (1 + 0.2 * np.cos(np.arange(tlay.size)).reshape(tlay.shape)))) for n, v in gases],
        "zero concentrations": [(n, (0.0 if not isinstance(v, np.ndarray) else v * 0)) for n, v in gases],
    }
    for key in ("lw_fsck", "sw_wide"):
        for name, gl in variants.items():
            out = run_gpu(gpu_models[key], plev, tlay, tlev, tsfc, gl, torch.float64, dev, KERNELS[kernel])
            check(out, run_oracle(c_oracles[key], ck_models[key], plev, tlay, tlev, tsfc, gl), 1e-12, f"{key}/{kernel}/{name}")


@pytest.mark.parametrize("kernel", list(KERNELS))
def test_surface_first_ordering(kernel, dev, gpu_models, ck_models, c_oracles):
    """Q6: upside-down columns.  Gas weights go negative -> clamped to zero, except relative_linear gases
    below their reference (negative x negative), which the reference keeps; Rayleigh goes negative unclamped."""
    plev, tlay, tlev, tsfc, gases = _small()
    f = np.asfortranarray
    flip = lambda a: f(a[:, ::-1])
    gl = [(n, flip(v) if isinstance(v, np.ndarray) else v) for n, v in gases]
    gl = [(n, (1e-7 if n == "ch4" else v)) for n, v in gl]      # ch4 below reference: contributes when dp < 0
    for key in ("lw_fsck", "sw_wide"):
        out = run_gpu(gpu_models[key], flip(plev), flip(tlay), flip(tlev), tsfc, gl, torch.float64, dev, KERNELS[kernel])
        ref = run_oracle(c_oracles[key], ck_models[key], flip(plev), flip(tlay), flip(tlev), tsfc, gl)
        # SW: negative Rayleigh + positive ch4 term cancel, which amplifies the (1e-16-level) differences in
        # summation order; the bar for this unphysical input is therefore 1e-10
        check(out, ref, 1e-12 if key == "lw_fsck" else 1e-10, f"{key}/{kernel}/upside-down")
        if key == "lw_fsck":
            assert (ref["tau"] > 0).any()       # the ch4 term survived the clamp


# ---- reference error behaviour --------------------------------------------------------------------
def test_missing_tlev_error_after_partial_write(dev, gpu_models, ck_models, c_oracles):
    plev, tlay, tlev, tsfc, gases = _small()
    with pytest.raises(ecckd_b200.EcckdError) as ei:
        run_gpu(gpu_models["lw_fsck"], plev, tlay, None, tsfc, gases, torch.float64, dev)
    assert ei.value.code == ecckd_b200.ERR_REFERENCE and str(ei.value) == "tlev is required for ecckd"   # :414-417
    ref = run_oracle(c_oracles["lw_fsck"], ck_models["lw_fsck"], plev, tlay, tlev, tsfc, gases)
    part = {k: v.cpu().numpy() for k, v in ei.value.partial.items()}
    check(part, {k: ref[k] for k in ("tau", "lay_source", "sfc_source")}, 1e-12, "partial")


def test_sw_needs_two_stream(dev, gpu_models, ck_models, c_oracles):
    plev, tlay, tlev, tsfc, gases = _small()
    with pytest.raises(ecckd_b200.EcckdError) as ei:
        run_gpu(gpu_models["sw_wide"], plev, tlay, None, None, gases, torch.float64, dev, two_stream=False)
    assert ei.value.code == ecckd_b200.ERR_REFERENCE and str(ei.value) == "shortwave must use ty_optical_props_2str"   # :461-463
    ref = run_oracle(c_oracles["sw_wide"], ck_models["sw_wide"], plev, tlay, tlev, tsfc, gases)
    check({"tau": ei.value.partial["tau"].cpu().numpy()}, {"tau": ref["tau"]}, 1e-12, "1scl tau")


def test_wrong_band_and_bad_arguments(dev, gpu_models):
    plev, tlay, tlev, tsfc, gases = _small()
    with pytest.raises(ecckd_b200.EcckdError) as ei:       # LW call on SW tables
        run_gpu(gpu_models["sw_wide"], plev, tlay, tlev, tsfc, gases, torch.float64, dev) if False else \
            gpu_models["sw_wide"].gas_optics(None, to_dev(plev, torch.float64, dev), to_dev(tlay, torch.float64, dev),
                                             torch.as_tensor(tsfc).to(dev), [], tlev=to_dev(tlev, torch.float64, dev))
    assert ei.value.code == ecckd_b200.ERR_ARG
    with pytest.raises(ValueError):                         # row-major array refused
        gpu_models["lw_fsck"].gas_optics(None, torch.zeros(40, 61, dtype=torch.float64, device=dev), to_dev(tlay, torch.float64, dev),
                                         torch.as_tensor(tsfc).to(dev), [], tlev=to_dev(tlev, torch.float64, dev))


def test_getters(gpu_models, ck_models):
    for key, go in gpu_models.items():
        m = ck_models[key]
        assert go.get_ngas() == m.num_gases and go.get_gases() == m.gas and go.get_ngpt() == m.ngpt
        assert go.source_is_internal() == (not m.shortwave) and go.source_is_external() == m.shortwave
        assert go.get_press_min() == np.exp(m.log_pressure[0]) and go.get_press_max() == np.exp(m.log_pressure[-1])   # :517-533
        assert go.get_temp_min() == m.temperature.min() and go.get_temp_max() == m.temperature.max()                 # :537-553
        assert np.array_equal(go.get_band2gpt().T, m.band2gpt) and go.get_nband() == m.band2gpt.shape[0]
    assert abs(gpu_models["sw_wide"].get_total_solar_irradiance() - ck_models["sw_wide"].total_solar_irradiance) < 1e-9
    assert abs(gpu_models["lw_fsck"].get_press_min() - 0.6940532) < 1e-6


# ---- host-array path (what the plain Fortran shim calls) --------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("block_cols", [0, 7, 64])
def test_host_path_equals_device_path(block_cols, dtype, dev, gpu_models):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    plev, tlay, tlev, tsfc, gases = atm_numpy(synthetic.make_atmosphere(50, 60, col0=5))
    c = lambda a: np.asfortranarray(np.asarray(a, dtype=dtype))
    gl = [(n, c(v) if isinstance(v, np.ndarray) else v) for n, v in gases] + [("cfc11", 1e-10 * np.ones(60))]
    for key in ("lw_fsck", "sw_wide"):
        go = gpu_models[key]
        lw = go.source_is_internal()
        host = go.gas_optics(None, c(plev), c(tlay), c(tsfc) if lw else None, gl, tlev=c(tlev) if lw else None, block_cols=block_cols)
        devo = run_gpu(go, c(plev), c(tlay), c(tlev), c(tsfc), gl, tdt, dev)
        for k in devo:
            assert np.array_equal(host[k], devo[k]), f"{key} {k}"      # same kernel, same per-column arithmetic: bitwise


def test_host_path_error_paths(dev, gpu_models):
    plev, tlay, tlev, tsfc, gases = _small()
    with pytest.raises(ecckd_b200.EcckdError) as ei:
        gpu_models["lw_fsck"].gas_optics(None, plev, tlay, tsfc, gases, tlev=None)
    assert str(ei.value) == "tlev is required for ecckd" and np.isfinite(ei.value.partial["tau"]).all()


# ---- sharding and size-independent properties -------------------------------------------------------
def test_column_shards_reproduce_the_unsharded_result_bitwise(dev, gpu_models):
    """SURVEY 8e determinism check: any column split gives bit-identical outputs (no cross-column coupling)."""
    from ecckd_b200 import sharding
    plev, tlay, tlev, tsfc, gases = atm_numpy(synthetic.make_atmosphere(203, 137))
    for key in ("lw_fsck", "sw_wide"):
        full = run_gpu(gpu_models[key], plev, tlay, tlev, tsfc, gases, torch.float64, dev)
        for world in (2, 8):
            parts = []
            for r in range(world):
                c0, c1 = sharding.shard_columns(203, world, r)
                f = np.asfortranarray
                gl = [(n, f(v[c0:c1]) if isinstance(v, np.ndarray) else v) for n, v in gases]
                parts.append(run_gpu(gpu_models[key], f(plev[c0:c1]), f(tlay[c0:c1]), f(tlev[c0:c1]), tsfc[c0:c1], gl, torch.float64, dev))
            for k in full:
                assert np.array_equal(np.concatenate([p[k] for p in parts], 0), full[k]), f"{key} {k} world={world}"


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_large_block_properties(dtype, dev, gpu_models, ck_models, c_oracles):
    """A 32768-column x 137-layer block generated on the device (the bench's generator and block shape class):
    size-independent properties plus oracle parity on sampled columns and walk-vs-generic agreement."""
    ncol, nlay = 32768, 137
    atm = synthetic.make_atmosphere(ncol, nlay, device=dev)
    if dtype == torch.float32:
        atm = atm.astype(dtype)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    pick = torch.tensor([0, 1, 15, 16, 17, 4095, 20000, ncol - 1], device=dev)
    for key in ("lw_fsck", "sw_wide"):
        go, m = gpu_models[key], ck_models[key]
        lw = not m.shortwave
        gases = ecckd_b200.PreparedGases(atm.gas_list(), ncol, nlay, npdt, dev)
        call = lambda kern: go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, gases,
                                          tlev=atm.tlev if lw else None, kernel=kern)
        out = call(ecckd_b200.KERNEL_WALK)
        gen = call(ecckd_b200.KERNEL_GENERIC)
        torch.cuda.synchronize()
        for k in out:                                                     # two independent CUDA formulations agree
            a, b = out[k].double(), gen[k].double()
            e = ((a - b).abs() / b.abs().clamp_min(1e-300)).max().item() if a.numel() else 0.0
            assert e <= (1e-12 if dtype == torch.float64 else 2e-5), f"{key} {k}: walk vs generic {e:.3e}"
        assert torch.isfinite(out["tau"]).all() and (out["tau"] > 0).all()
        if lw:
            # the level source is computed once and written twice (:423-424): inc(:, j) == dec(:, j+1), bitwise
            assert torch.equal(out["lev_source_inc"][:, :-1, :], out["lev_source_dec"][:, 1:, :])
            assert (out["lay_source"] > 0).all() and (out["sfc_source"] > 0).all()
        else:
            assert (out["g"] == 0).all() and ((out["ssa"] > 0) & (out["ssa"] < 1)).all()
            sol = torch.as_tensor(m.solar_irradiance, device=dev).to(dtype)
            assert torch.equal(out["toa_src"], sol[None, :].expand(ncol, -1))
        # oracle parity on sampled columns of exactly these device-generated inputs
        sel = lambda t: np.asfortranarray(t[pick].double().cpu().numpy())
        gl = [(n, sel(v) if torch.is_tensor(v) else (float(np.float32(v)) if dtype == torch.float32 else v)) for n, v in atm.gas_list()]
        ref = run_oracle(c_oracles[key], m, sel(atm.plev), sel(atm.tlay), sel(atm.tlev), atm.tsfc[pick].double().cpu().numpy(), gl)
        check({k: v[pick].cpu().numpy() for k, v in out.items()}, ref, TOL[dtype], f"{key} sampled")
        # permuting the columns permutes the outputs, bitwise (columns are independent)
        perm = torch.randperm(4096, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
        sub = lambda t: ecckd_b200.to_fortran(t[:4096][perm], dtype, dev)
        gl2 = [(n, sub(v) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
        o2 = go.gas_optics(None, sub(atm.plev), sub(atm.tlay), atm.tsfc[:4096][perm].contiguous() if lw else None, gl2,
                           tlev=sub(atm.tlev) if lw else None)
        assert torch.equal(o2["tau"], out["tau"][:4096][perm])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("ncol,nlay", [(4, 1), (32, 7), (36, 137), (5000, 9), (4736 + 32, 3)])
def test_quad_kernel_equals_column_walk_bitwise(ncol, nlay, dtype, dev, gpu_models):
    """The quad kernel (4 columns per warp, TMA stores) and the one-column-per-warp kernel run the same explicitly
    rounded arithmetic per g-point: every output agrees bit for bit, for all three ck files, also with several
    tiles per persistent block (5000 columns = 157 tiles > 148 SMs)."""
    atm = synthetic.make_atmosphere(ncol, nlay, device=dev)
    if dtype == torch.float32:
        atm = atm.astype(dtype)
    npdt = np.float64 if dtype == torch.float64 else np.float32
    for key, go in gpu_models.items():
        lw = go.source_is_internal()
        gases = ecckd_b200.PreparedGases(atm.gas_list(), ncol, nlay, npdt, dev)
        call = lambda kern: go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, gases,
                                          tlev=atm.tlev if lw else None, kernel=kern)
        q, w = call(ecckd_b200.KERNEL_QUAD), call(ecckd_b200.KERNEL_COLWALK)
        torch.cuda.synchronize()
        for k in q:
            assert torch.equal(q[k], w[k]), f"{key} {k}: quad and column-walk kernels differ"


def test_quad_kernel_refuses_what_it_cannot_map(dev, gpu_models):
    """Odd fp64 column counts have no 16-byte-aligned layer pitch: no tensor map, so QUAD is an argument error and
    AUTO / WALK fall back to the column-walk kernel."""
    atm = synthetic.make_atmosphere(33, 5, device=dev)
    go = gpu_models["lw_fsck"]
    with pytest.raises(ecckd_b200.EcckdError):
        go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc, atm.gas_list(), tlev=atm.tlev, kernel=ecckd_b200.KERNEL_QUAD)
    go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc, atm.gas_list(), tlev=atm.tlev, kernel=ecckd_b200.KERNEL_WALK)


@pytest.mark.parametrize("ncol", [18922, 18921], ids=["even-tma", "odd-writer"])
def test_persistent_blocks_with_ragged_last_chunk(ncol, dev, gpu_models, ck_models, c_oracles):
    """Enough columns for the persistent two-tile-chunk geometry (1183 tiles: the last chunk holds one tile, the last
    tile 10 / 9 columns), few layers so that the oracle covers every column; even ncol takes the TMA path, odd ncol
    the writer-warp path."""
    inp = atm_numpy(synthetic.make_atmosphere(ncol, 3))
    for key in ("lw_fsck", "lw_rrtmgp", "sw_wide"):
        out = run_gpu(gpu_models[key], *inp, torch.float64, dev, ecckd_b200.KERNEL_WALK)
        check(out, run_oracle(c_oracles[key], ck_models[key], *inp), 1e-12, f"{key}/ncol={ncol}")


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("varying", [("co2",), ("co2", "ch4"), ("co2", "ch4", "n2o"), ("co2", "ch4", "n2o", "cfc11", "cfc12", "o2")],
                         ids=["nvar2", "nvar3", "nvar4", "nvar7"])
def test_persistent_blocks_with_several_varying_gases(varying, dtype, dev, gpu_models, ck_models, c_oracles):
    """Bilinear gases given as (ncol, nlay) fields instead of scalars (co2 / ch4 / n2o profiles, or everything as a field
    like a host model that never passes scalars): with o3 that is nvar = 2, 3 (the quad kernel's three-gas
    instantiation), 4 and 7 (the 8-slot column-walk instantiation).  9472 columns: several tiles per persistent block.
    ch4 / n2o wobble across their reference mole fractions, so the per-gas clamp (:233-238) fires in part of the cells."""
    ncol, nlay = 9472, 3
    plev, tlay, tlev, tsfc, gases = atm_numpy(synthetic.make_atmosphere(ncol, nlay))
    wobble = 1 + 0.2 * np.cos(np.arange(tlay.size)).reshape(tlay.shape, order="F")
    gl = [(n, v if isinstance(v, np.ndarray) else (np.asfortranarray(v * wobble) if n in varying else v)) for n, v in gases]
    for key in ("lw_fsck", "sw_wide"):
        ref = run_oracle(c_oracles[key], ck_models[key], plev, tlay, tlev, tsfc, gl, dtype)
        for kern in (ecckd_b200.KERNEL_WALK, ecckd_b200.KERNEL_COLWALK):
            out = run_gpu(gpu_models[key], plev, tlay, tlev, tsfc, gl, dtype, dev, kern)
            check(out, ref, TOL[dtype], f"{key}/varying={varying}/kernel={kern}")


def test_linearity_in_a_linear_gas(dev, gpu_models):
    """tau is affine in the concentration of a code-1 gas: tau(2x) - tau(x) == tau(x) - tau(0) to rounding."""
    plev, tlay, tlev, tsfc, gases = _small()
    base = dict(gases)
    def tau(scale):
        gl = [(n, (v * scale if n == "o3" else v)) for n, v in gases]
        return run_gpu(gpu_models["lw_fsck"], plev, tlay, tlev, tsfc, gl, torch.float64, dev)["tau"]
    t0, t1, t2 = tau(0.0), tau(1.0), tau(2.0)
    d1, d2 = t1 - t0, t2 - t1
    assert np.max(np.abs(d2 - d1) / np.maximum(t2, 1e-300)) <= 1e-13


# ---- downstream fluxes (north_star: within 0.01 W m-2) ------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_downstream_fluxes_within_0p01_wm2(dtype, dev, gpu_models, ck_models, c_oracles):
    """The same flux solver (tests/rte_solvers.py: the rte-rrtmgp formulation, which is not available offline) applied
    to the oracle's and to the GPU's optical properties and sources: broadband fluxes agree to < 0.01 W m-2,
    for the fp64 path and for the fp32 variant."""
    import rte_solvers
    inp = atm_numpy(synthetic.make_atmosphere(64, 137, col0=4242))
    out = run_gpu(gpu_models["lw_fsck"], *inp, dtype, dev)
    ref = run_oracle(c_oracles["lw_fsck"], ck_models["lw_fsck"], *inp)     # fp64 oracle on the fp64 inputs
    f = lambda d: rte_solvers.lw_fluxes(*(np.asarray(d[k], dtype=np.float64) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec", "sfc_source")))
    (up_g, dn_g), (up_r, dn_r) = f(out), f(ref)
    assert 150 < up_r[:, 0].mean() < 400 and 150 < dn_r[:, -1].mean() < 500        # sane OLR / surface downwelling
    assert max(np.abs(up_g - up_r).max(), np.abs(dn_g - dn_r).max()) < 0.01
    out = run_gpu(gpu_models["sw_wide"], *inp, dtype, dev)
    ref = run_oracle(c_oracles["sw_wide"], ck_models["sw_wide"], *inp)
    f = lambda d: rte_solvers.sw_fluxes(*(np.asarray(d[k], dtype=np.float64) for k in ("tau", "ssa", "g", "toa_src")))
    sg, sr = f(out), f(ref)
    assert 500 < sr[1][:, -1].mean() < 820                                           # surface insolation at mu0 = 0.6
    assert max(np.abs(a - b).max() for a, b in zip(sg, sr)) < 0.01


@pytest.mark.parametrize("key", ["lw_fsck", "lw_rrtmgp", "sw_wide"])
def test_device_flux_solvers_match_the_numpy_solvers(key, dev, gpu_models, ck_models, c_oracles):
    """Row N1 on the device: ecckd_rte_{lw,sw}_f64 on the GPU's optical properties vs tests/rte_solvers.py on the SAME
    arrays (<= 1e-6 W m-2: same formulation; exp() differs in the last bit between CUDA and numpy, and the linear-in-tau
    source factor (1-t)/x - t amplifies that for thin layers), and vs the numpy solver on the ORACLE's optical
    properties (< 0.01 W m-2, the BASELINE criterion).  lw_rrtmgp has 36 g-points: two g-passes per block."""
    import rte_solvers
    ncol, nlay = 70, 137
    atm = synthetic.make_atmosphere(ncol, nlay, col0=977, device=dev)
    go, m = gpu_models[key], ck_models[key]
    lw = not m.shortwave
    out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, atm.gas_list(), tlev=atm.tlev if lw else None)
    host = {k: v.cpu().numpy() for k, v in out.items()}
    inp = atm_numpy(synthetic.make_atmosphere(ncol, nlay, col0=977))
    ref = run_oracle(c_oracles[key], m, *inp)
    if lw:
        got = [t.cpu().numpy() for t in go.rte_lw(out)]
        f = lambda d: rte_solvers.lw_fluxes(*(np.asarray(d[k], dtype=np.float64) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec", "sfc_source")))
    else:
        got = [t.cpu().numpy() for t in go.rte_sw(out)]
        f = lambda d: rte_solvers.sw_fluxes(*(np.asarray(d[k], dtype=np.float64) for k in ("tau", "ssa", "g", "toa_src")))
    same, orc = f(host), f(ref)
    for a, b, c in zip(got, same, orc):
        assert np.isfinite(a).all()
        assert np.abs(a - b).max() <= 1e-6, f"{key}: device solver vs numpy solver on the same optical properties: {np.abs(a - b).max():.3e} W m-2"
        assert np.abs(a - c).max() < 0.01, f"{key}: device pipeline vs oracle + numpy solver: {np.abs(a - c).max():.3e} W m-2"
    assert got[0].max() > 50.0                                       # real fluxes, not zeros


def test_in_process_multi_device_host_sharding(dev):
    """A model replicated on every visible GPU: the host entry points shard the columns over the devices (no
    collective) and reproduce the single-device result bitwise.  Needs >= 2 GPUs (verified on a 2-GPU box with
    tools/multidev_host_check.py); skipped on the single-GPU test box."""
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("single GPU")
    plev, tlay, tlev, tsfc, gases = atm_numpy(synthetic.make_atmosphere(333, 60))
    for key in ("lw_fsck", "sw_wide"):
        one = ecckd_b200.GasOpticsEcckd.load(FILES[key], devices=[0])
        many = ecckd_b200.GasOpticsEcckd.load(FILES[key], devices=list(range(ndev)))
        lw = one.source_is_internal()
        a = one.gas_optics(None, plev, tlay, tsfc if lw else None, gases, tlev=tlev if lw else None)
        b = many.gas_optics(None, plev, tlay, tsfc if lw else None, gases, tlev=tlev if lw else None, block_cols=50)
        for k in a:
            assert np.array_equal(a[k], b[k]), (key, k)
