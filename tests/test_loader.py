"""The product's C++ netCDF-3 reader + load_and_init against the oracle's independent numpy restatement,
both following example/rfmip-rad-irf/mo_load_coefficients.F90:19-203."""
import ctypes as C

import numpy as np
import pytest

from conftest import FILES


def _load_cxx(path):
    from ecckd_b200 import _abi
    lib = _abi.load_library()
    tp = C.POINTER(_abi.HostTables)()
    err = C.create_string_buffer(_abi.ERRLEN)
    rc = lib.ecckd_tables_load(path.encode(), C.byref(tp), err)
    return lib, rc, tp, err.value.decode()


def _arr(ptr, n, dtype=np.float64):
    if not ptr:
        return None
    ct = C.c_double if dtype == np.float64 else C.c_int32
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


@pytest.mark.parametrize("key", list(FILES))
def test_cxx_loader_equals_numpy_loader(key, ck_models):
    m = ck_models[key]
    lib, rc, tp, err = _load_cxx(FILES[key])
    assert rc == 0, err
    t = tp.contents
    try:
        assert bool(t.shortwave) == m.shortwave
        assert (t.ngpt, t.npress, t.ntemp) == (m.ngpt, m.log_pressure.shape[0], m.temperature.shape[0])
        assert t.num_gases == m.num_gases
        np.testing.assert_array_equal(_arr(t.log_pressure, t.npress), m.log_pressure)     # same libm log
        np.testing.assert_array_equal(_arr(t.temperature, t.npress * t.ntemp), m.temperature.ravel())
        if m.shortwave:
            np.testing.assert_array_equal(_arr(t.solar_irradiance, t.ngpt), m.solar_irradiance)
            np.testing.assert_array_equal(_arr(t.rayleigh_molar_scattering_coeff, t.ngpt), m.rayleigh_molar_scattering_coeff)
            assert t.nplanck == 0
        else:
            assert t.nplanck == m.temperature_planck.shape[0]
            np.testing.assert_array_equal(_arr(t.temperature_planck, t.nplanck), m.temperature_planck)
            np.testing.assert_array_equal(_arr(t.planck_function, t.nplanck * t.ngpt), m.planck_function.ravel())
        assert t.nband == m.band2gpt.shape[0]
        np.testing.assert_array_equal(_arr(t.band2gpt, 2 * t.nband, np.int32).reshape(-1, 2), m.band2gpt)
        np.testing.assert_array_equal(_arr(t.band_lims_wvn, 2 * t.nband).reshape(-1, 2), m.band_lims_wvn)
        for i in range(t.num_gases):
            g, tab = t.gas[i], m.absorption[i]
            assert g.name.decode().strip() == m.gas[i]
            assert bool(g.composite_only) == tab.composite_only
            assert g.concentration_dependence_code == tab.concentration_dependence_code
            assert g.reference_mole_fraction == tab.reference_mole_fraction
            nmf = 0 if tab.mole_fraction is None else tab.mole_fraction.shape[0]
            assert g.n_mole_fraction == nmf
            if nmf:
                np.testing.assert_array_equal(_arr(g.mole_fraction, nmf), tab.mole_fraction)
            np.testing.assert_array_equal(_arr(g.coefficient, tab.coefficient.size), tab.coefficient.ravel())
    finally:
        lib.ecckd_tables_free(tp)


def test_gas_lists_follow_the_loader_rules(ck_models):
    # file order minus "composite", then the composite constituents not already listed (:119-143)
    assert ck_models["lw_fsck"].gas == ["h2o", "o3", "co2", "ch4", "n2o", "cfc11", "cfc12", "o2", "n2"]
    assert ck_models["lw_rrtmgp"].gas == ["h2o", "o3", "co2", "ch4", "n2o", "cfc11", "cfc12", "o2", "n2"]
    assert ck_models["sw_wide"].gas == ["h2o", "o3", "co2", "ch4", "n2o", "o2", "n2"]
    for m in ck_models.values():
        codes = dict(zip(m.gas, (t.concentration_dependence_code for t in m.absorption)))
        assert codes["h2o"] == 2 and codes["o3"] == 1 and codes["co2"] == 1
        assert codes["ch4"] == 3 and codes["n2o"] == 3 and codes["o2"] == 0 and codes["n2"] == 0
        assert [t.composite_only for t in m.absorption][-2:] == [True, True]
        assert m.num_composite_gases == 4


def test_table_facts(ck_models):
    """the facts the kernels rely on."""
    for m in ck_models.values():
        assert m.log_pressure.shape == (53,) and m.temperature.shape == (6, 53)
        assert abs((m.log_pressure[1] - m.log_pressure[0]) - 0.2302585450313313) < 1e-6
        assert m.temperature[1, 0] - m.temperature[0, 0] == 20.0
        for t in m.absorption:
            assert (t.coefficient >= 0).all()                                           # clamp == sign test of the weight
            assert np.array_equal(t.coefficient, t.coefficient.astype(np.float32))      # exact in fp32
    assert ck_models["lw_fsck"].ngpt == 32 and ck_models["lw_rrtmgp"].ngpt == 36 and ck_models["sw_wide"].ngpt == 27
    assert ck_models["lw_rrtmgp"].band2gpt.shape == (16, 2) and ck_models["sw_wide"].band2gpt.shape == (5, 2)
    assert abs(ck_models["sw_wide"].total_solar_irradiance - 1361.0000124) < 1e-6
    ch4 = ck_models["lw_fsck"].absorption[3]
    assert ch4.reference_mole_fraction == float(np.float32(1.921e-06))


def test_tokenize_quirk():
    from oracle.ckd_model import tokenize
    assert tokenize("composite h2o o3 co2 ch4 n2o cfc11 cfc12") == ["composite", "h2o", "o3", "co2", "ch4", "n2o", "cfc11", "cfc12"]
    assert tokenize("  o2   n2  ") == ["o2", "n2"]
    # a trailing one-character token is dropped by the reference's tokenizer (:260-291)
    assert tokenize("o2 n2 x") == ["o2", "n2"]


def test_loader_errors(tmp_path):
    lib, rc, tp, err = _load_cxx(str(tmp_path / "missing.nc"))
    assert rc != 0 and "can't open file" in err
    bad = tmp_path / "bad.nc"
    bad.write_bytes(b"HDF5 not classic")
    lib, rc, tp, err = _load_cxx(str(bad))
    assert rc != 0 and "netCDF classic" in err
    from oracle import netcdf3
    with pytest.raises(ValueError):
        netcdf3.Dataset(str(bad))
