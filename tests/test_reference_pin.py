"""Pins the oracle to the reference itself: the UNMODIFIED /root/reference/src/gas_optics_ecckd.f90, compiled
against stub rte-rrtmgp modules (oracle/ref), run on the committed golden inputs and on an RFMIP-shaped case,
must agree with the C oracle to <= 1e-14 relative (different libm `log`/`exp` builds may differ in the last bit).

Skipped -- and parity stays UNPINNED -- when no Fortran compiler or no reference tree is available
(this image has neither gfortran nor flang; see oracle/ref/README.md for the recipe)."""
import os

import numpy as np
import pytest

from conftest import FILES, atm_numpy, rel_err
from oracle.ref import run_ref

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

needs_ref = pytest.mark.skipif(not run_ref.available(),
                               reason="no Fortran compiler / reference sources: oracle/_ref/ref_driver cannot be built "
                                      "(parity unpinned, recipe in oracle/ref)")


def test_recipe_is_complete():
    """The recipe is committed whether or not it can run here: stubs for every module the reference uses
    (src/gas_optics_ecckd.f90:2-7), the driver, the Makefile and the front end."""
    here = os.path.dirname(run_ref.__file__)
    for f in ("Makefile", "ref_driver.f90", "run_ref.py", "stubs/mo_rte_kind.F90", "stubs/mo_optical_props.F90",
              "stubs/mo_source_functions.F90", "stubs/mo_gas_concentrations.F90", "stubs/mo_gas_optics.F90"):
        assert os.path.exists(os.path.join(here, f)), f
    # every `use` of the reference module is served by a stub
    src = os.path.join(run_ref.REFERENCE, "src", "gas_optics_ecckd.f90")
    if os.path.exists(src):
        used = {l.split()[1].rstrip(",") for l in open(src) if l.startswith("use ") and "intrinsic" not in l}
        assert used == {"mo_gas_concentrations", "mo_gas_optics", "mo_optical_props", "mo_rte_kind", "mo_source_functions"}


def test_flat_files_round_trip(tmp_path, ck_models):
    """The dump the driver reads: sizes add up (a malformed file would make every Fortran read misalign)."""
    for key, m in ck_models.items():
        p = tmp_path / f"{key}.bin"
        run_ref.dump_tables(m, str(p))
        n = 7 * 4
        for tab in m.absorption:
            nv = 0 if tab.mole_fraction is None else tab.mole_fraction.shape[0]
            n += 32 + 12 + 8 + 8 * nv + 8 * tab.coefficient.size
        n += 8 * (m.log_pressure.size + m.temperature.size)
        n += 8 * (2 * m.ngpt + 1) if m.shortwave else 8 * (m.temperature_planck.size + m.planck_function.size)
        assert os.path.getsize(p) == n


@needs_ref
@pytest.mark.parametrize("key", list(FILES))
def test_oracle_equals_reference_fortran_on_golden_inputs(key, ck_models, c_oracles):
    z = np.load(os.path.join(GOLD, "inputs.npz"))
    f = np.asfortranarray
    gases = [("co2", 420e-6), ("ch4", 1.0e-6), ("n2o", 4.0e-7), ("o2", 0.209), ("cfc11", 2.3e-10),
             ("cfc12", 5.2e-10), ("h2o", f(z["h2o"])), ("o3", f(z["o3"])), ("no2", 0.0)]
    m, co = ck_models[key], c_oracles[key]
    ref, err = run_ref.run(m, f(z["plev"]), f(z["tlay"]), z["tsfc"], gases, tlev=f(z["tlev"]))
    assert err == ""
    out, err2 = (co.gas_optics_ext(f(z["plev"]), f(z["tlay"]), gases) if m.shortwave
                 else co.gas_optics_int(f(z["plev"]), f(z["tlay"]), z["tsfc"], gases, f(z["tlev"])))
    assert err2 == ""
    for k, v in ref.items():
        assert rel_err(out[k], v) <= 1e-14, f"{key} {k}: oracle vs reference Fortran"
    gold = np.load(os.path.join(GOLD, f"{key}.npz"))
    for k, v in ref.items():
        assert rel_err(gold[k], v) <= 1e-14, f"{key} {k}: committed fixture vs reference Fortran"


@needs_ref
@pytest.mark.parametrize("key", list(FILES))
def test_oracle_equals_reference_fortran_rfmip_shaped(key, ck_models, c_oracles):
    from ecckd_b200 import synthetic
    m, co = ck_models[key], c_oracles[key]
    for atm in synthetic.make_rfmip_like()[:3]:
        plev, tlay, tlev, tsfc, gl = atm_numpy(atm)
        ref, err = run_ref.run(m, plev, tlay, tsfc, gl, tlev=tlev)
        assert err == ""
        out, _ = co.gas_optics_ext(plev, tlay, gl) if m.shortwave else co.gas_optics_int(plev, tlay, tsfc, gl, tlev)
        for k, v in ref.items():
            assert rel_err(out[k], v) <= 1e-14, f"{key} {k}"


@needs_ref
def test_reference_error_strings(ck_models):
    """tlev missing (:414-417) -- after tau, lay_source and sfc_source have been written (Q13)."""
    m = ck_models["lw_fsck"]
    from ecckd_b200 import synthetic
    plev, tlay, tlev, tsfc, gl = atm_numpy(synthetic.make_atmosphere(4, 10))
    ref, err = run_ref.run(m, plev, tlay, tsfc, gl, tlev=None)
    assert err == "tlev is required for ecckd"
    assert (ref["tau"] > 0).all() and (ref["lay_source"] > 0).all() and (ref["lev_source_inc"] == 0).all()
