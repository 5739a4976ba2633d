"""Shared fixtures.  `-m "not gpu"` covers the oracle, the host logic and the ABI surface;
`-m gpu` is the parity suite proper (CUDA path through the C ABI vs the CPU oracle)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rte-ecckd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

DATA = os.path.join(ROOT, "data")
FILES = {
    "lw_fsck": os.path.join(DATA, "ecckd-1.2_lw_ckd-definition_climate_fsck-tol0.0161.nc"),
    "lw_rrtmgp": os.path.join(DATA, "ecckd-1.2_lw_ckd-definition_climate_rrtmgp-tol0.061.nc"),
    "sw_wide": os.path.join(DATA, "ecckd-1.2_sw_ckd-definition_climate_wide-tol0.05.nc"),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the C oracle exist (no-op when already built)."""
    import __graft_entry__ as ge
    from ecckd_b200 import _abi
    if not os.path.exists(_abi.LIB_PATH):
        ge.build()
    from oracle import c_oracle
    c_oracle.build()


@pytest.fixture(scope="session")
def ck_models():
    from oracle import ckd_model
    return {k: ckd_model.load_and_init(v) for k, v in FILES.items()}


@pytest.fixture(scope="session")
def c_oracles(ck_models):
    from oracle.c_oracle import COracle
    return {k: COracle(m) for k, m in ck_models.items()}


def rel_err(a, b):
    """max |a-b| / |b| with exact zeros (and bit-equal values) counting as 0.
    A NaN on one side only is an infinite error (a kernel that writes NaN where the
    oracle is finite must fail); NaN on both sides at the same element counts as equal."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return float("inf")
    if a.size == 0:
        return 0.0
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    if np.any(nan_a != nan_b):
        return float("inf")
    den = np.maximum(np.abs(b), np.finfo(np.float64).tiny)
    with np.errstate(invalid="ignore"):
        e = np.where((a == b) | (nan_a & nan_b), 0.0, np.abs(a - b) / den)
    return float(np.max(e))


def atm_numpy(atm):
    """(plev, tlay, tlev, tsfc, gas list) of a synthetic.Atmosphere as Fortran-order numpy."""
    import torch
    F = lambda t: np.asfortranarray(t.cpu().numpy())
    gl = [(n, F(v) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
    return F(atm.plev), F(atm.tlay), F(atm.tlev), atm.tsfc.cpu().numpy().copy(), gl
