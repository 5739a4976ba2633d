"""Pins the oracle.  The reference ships no golden vectors and cannot be built here (parity unpinned by
the reference's own tests), so the oracle is pinned by (a) bit-for-bit agreement of two independently
written restatements (C, numpy), (b) hand-derivable constants, (c) the reference's documented branch
behaviour on hand-built inputs, (d) committed regression fixtures."""
import os

import numpy as np
import pytest

from conftest import FILES, atm_numpy, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden_inputs():
    z = np.load(os.path.join(GOLD, "inputs.npz"))
    f = np.asfortranarray
    gases = [("co2", 420e-6), ("ch4", 1.0e-6), ("n2o", 4.0e-7), ("o2", 0.209), ("cfc11", 2.3e-10),
             ("cfc12", 5.2e-10), ("h2o", f(z["h2o"])), ("o3", f(z["o3"])), ("no2", 0.0)]
    return f(z["plev"]), f(z["tlay"]), f(z["tlev"]), z["tsfc"], gases


def _run(co, m, plev, tlay, tlev, tsfc, gases):
    if m.shortwave:
        return co.gas_optics_ext(plev, tlay, gases)
    return co.gas_optics_int(plev, tlay, tsfc, gases, tlev)


# ---- (a) twin agreement -------------------------------------------------------------------------
@pytest.mark.parametrize("key", list(FILES))
def test_c_oracle_equals_numpy_twin_bitwise(key, ck_models, c_oracles):
    from ecckd_b200 import synthetic
    from oracle import ecckd_oracle_np as onp
    m, co = ck_models[key], c_oracles[key]
    for atm in [synthetic.make_atmosphere(9, 137)] + synthetic.make_rfmip_like(nexp=4, ncol_per_exp=5, nlay=60):
        plev, tlay, tlev, tsfc, gl = atm_numpy(atm)
        out_c, err = _run(co, m, plev, tlay, tlev, tsfc, gl)
        assert err == ""
        out_n = onp.gas_optics_ext(m, plev, tlay, gl) if m.shortwave else onp.gas_optics_int(m, plev, tlay, tsfc, gl, tlev)
        for k in out_c:
            assert np.array_equal(out_c[k], out_n[k]), k


@pytest.mark.parametrize("key", list(FILES))
def test_indices_twin_and_ranges(key, ck_models, c_oracles):
    from oracle import ecckd_oracle_np as onp
    m, co = ck_models[key], c_oracles[key]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    h2o = dict(gases)["h2o"]
    idx = co.indices("h2o", plev, tlay, h2o)
    ip0, pw0, pw1, it0, tw0, tw1 = onp.interpolation_indices(m, plev, tlay)
    iv0, vw0, vw1 = onp.vmr_indices(m.absorption[0].mole_fraction, h2o)
    assert np.array_equal(idx["ip0"], ip0) and np.array_equal(idx["it0"], it0) and np.array_equal(idx["iv0"], iv0)
    assert np.array_equal(idx["pw1"], pw1) and np.array_equal(idx["tw1"], tw1) and np.array_equal(idx["vw1"], vw1)
    # index ranges; the hand-built input reaches both clamps of every axis
    assert ip0.min() == 1 and ip0.max() == 52 and it0.min() == 1 and it0.max() == 5
    assert iv0.min() == 1 and iv0.max() == 11
    assert idx["vw1"].max() <= 0.999 + 1e-12 and idx["pw1"].max() <= 0.9999 + 1e-12


# ---- (b) constants ------------------------------------------------------------------------------
def test_default_real_literals(c_oracles):
    c = c_oracles["lw_fsck"].constants()
    assert c[0] == 9.8066501617431641 and c[1] == 28.969999313354492           # gravity, dry_air_molar_mass
    assert c[2] == 3.1415927410125732 and c[3] == 0.0010000000474974513        # pi, 0.001
    assert c[4] == float.fromhex("0x1.c28c365d373bcp+1")                       # global_weight
    from oracle import ecckd_oracle_np as onp
    assert float(onp.global_weight()) == c[4]


# ---- (c) branch behaviour -----------------------------------------------------------------------
def test_unknown_gas_is_skipped_and_caller_order_rules(ck_models, c_oracles):
    m, co = ck_models["lw_fsck"], c_oracles["lw_fsck"]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    base, _ = co.gas_optics_int(plev, tlay, tsfc, gases, tlev)
    extra, _ = co.gas_optics_int(plev, tlay, tsfc, gases + [("sf6", 1e-9), ("co", 1e-7)], tlev)
    assert np.array_equal(base["tau"], extra["tau"])                            # :358-364
    # a model gas the caller does not supply contributes nothing
    no_o3, _ = co.gas_optics_int(plev, tlay, tsfc, [g for g in gases if g[0] != "o3"], tlev)
    assert (no_o3["tau"] <= base["tau"]).all() and (no_o3["tau"] < base["tau"]).any()


def test_composite_is_counted_once(ck_models, c_oracles):
    co = c_oracles["lw_fsck"]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    only = lambda names: [g for g in gases if g[0] in names]
    t_o2, _ = co.gas_optics_int(plev, tlay, tsfc, only({"o2"}), tlev)
    t_n2, _ = co.gas_optics_int(plev, tlay, tsfc, [("n2", 0.781)], tlev)
    t_both, _ = co.gas_optics_int(plev, tlay, tsfc, only({"o2"}) + [("n2", 0.781)], tlev)
    t_none, _ = co.gas_optics_int(plev, tlay, tsfc, [], tlev)
    assert np.array_equal(t_o2["tau"], t_n2["tau"]) and np.array_equal(t_o2["tau"], t_both["tau"])   # :365-373, vmr ignored :216
    assert (t_none["tau"] == 0).all() and (t_o2["tau"] > 0).all()


def test_relative_linear_clamp(ck_models, c_oracles):
    """ch4 / n2o below their reference mole fraction contribute exactly zero (:146, :233-238)."""
    co = c_oracles["lw_fsck"]
    plev, tlay, tlev, tsfc, _ = _golden_inputs()
    ref = ck_models["lw_fsck"].absorption[3].reference_mole_fraction
    below, _ = co.gas_optics_int(plev, tlay, tsfc, [("ch4", 0.5 * ref)], tlev)
    at, _ = co.gas_optics_int(plev, tlay, tsfc, [("ch4", ref)], tlev)
    above, _ = co.gas_optics_int(plev, tlay, tsfc, [("ch4", 2 * ref)], tlev)
    assert (below["tau"] == 0).all() and (at["tau"] == 0).all() and (above["tau"] > 0).any()


def test_surface_first_ordering_clamps_everything(c_oracles):
    """Q6: negative layer thickness makes every gas optical depth non-positive -> clamped to 0."""
    co = c_oracles["lw_fsck"]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    f = np.asfortranarray
    gl = [(n, f(v[:, ::-1]) if isinstance(v, np.ndarray) else v) for n, v in gases if n not in ("ch4", "n2o")]
    out, err = co.gas_optics_int(f(plev[:, ::-1]), f(tlay[:, ::-1]), tsfc, gl, f(tlev[:, ::-1]))
    assert err == "" and (out["tau"] == 0).all()


def test_error_paths(c_oracles):
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    out, err = c_oracles["lw_fsck"].gas_optics_int(plev, tlay, tsfc, gases, None)
    assert err == "tlev is required for ecckd"                                    # :414-417 ...
    assert np.isfinite(out["tau"]).all() and np.isfinite(out["lay_source"]).all() and np.isfinite(out["sfc_source"]).all()
    assert np.isnan(out["lev_source_inc"]).all()                                  # ... raised after tau/lay/sfc were written
    out, err = c_oracles["sw_wide"].gas_optics_ext(plev, tlay, gases, two_stream=False)
    assert err == "shortwave must use ty_optical_props_2str" and np.isfinite(out["tau"]).all() and np.isnan(out["ssa"]).all()


def test_planck_branches(ck_models, c_oracles):
    m, co = ck_models["lw_fsck"], c_oracles["lw_fsck"]
    T = np.array([100.0, 119.999, 120.0, 120.5, 349.9, 350.0, 350.7, 400.0])
    it0, w1 = co.planck_indices(T)
    assert list(it0[:, 0]) == [0, 0, 1, 1, 230, 230, 230, 230]                    # :276-284, min(int, n-1) :278
    assert w1[6, 0] > 1.0 and w1[7, 0] > 50.0                                     # linear extrapolation above the LUT (Q9)
    from oracle import ecckd_oracle_np as onp
    B = onp.calculate_planck_function(m, T[:, None])[:, 0, :]
    np.testing.assert_array_equal(B[0], ((100.0 / 120.0) * m.planck_function[0]) / onp.PI)   # :284, :288
    np.testing.assert_array_equal(B[2], (1.0 * m.planck_function[0] + 0.0 * m.planck_function[1]) / onp.PI)


def test_sw_outputs(ck_models, c_oracles):
    m, co = ck_models["sw_wide"], c_oracles["sw_wide"]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    out, err = co.gas_optics_ext(plev, tlay, gases)
    assert err == "" and (out["g"] == 0).all()                                    # :460
    assert np.array_equal(out["toa_src"], np.broadcast_to(m.solar_irradiance, out["toa_src"].shape))   # :468-472, no TSI scaling
    assert ((out["ssa"] > 0) & (out["ssa"] < 1)).all()


# ---- (d) regression fixtures --------------------------------------------------------------------
@pytest.mark.parametrize("key", list(FILES))
def test_oracle_reproduces_committed_fixtures(key, ck_models, c_oracles):
    m, co = ck_models[key], c_oracles[key]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    out, err = _run(co, m, plev, tlay, tlev, tsfc, gases)
    assert err == ""
    z = np.load(os.path.join(GOLD, f"{key}.npz"))
    for k in out:
        # libm's log may differ in the last bit across glibc builds; everything else is plain IEEE arithmetic
        assert rel_err(out[k], z[k]) <= 1e-14, k
    idx = co.indices("h2o", plev, tlay, dict(gases)["h2o"])
    for k in ("ip0", "it0", "iv0"):
        assert np.array_equal(idx[k], z[f"idx_{k}"]), k


def test_fast_build_matches_checker_build(ck_models):
    """The -O3 -march=native build that is timed as the CPU baseline computes the same thing (FMA contraction only)."""
    from oracle.c_oracle import COracle
    m = ck_models["lw_fsck"]
    plev, tlay, tlev, tsfc, gases = _golden_inputs()
    a, _ = COracle(m).gas_optics_int(plev, tlay, tsfc, gases, tlev)
    b, _ = COracle(m, fast=True).gas_optics_int(plev, tlay, tsfc, gases, tlev)
    for k in a:
        assert rel_err(b[k], a[k]) <= 1e-13, k
