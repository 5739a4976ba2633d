"""Generates tests/golden/*.npz.

The reference ships no golden vectors and cannot be built offline (no Fortran
compiler / netCDF / rte-rrtmgp), so these are REGRESSION fixtures produced by the
C oracle (oracle/ecckd_oracle.c) on small fixed inputs -- they pin the oracle's
behaviour over time, they are not reference output.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
from oracle import ckd_model                     # noqa: E402
from oracle.c_oracle import COracle              # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "lw_fsck": "ecckd-1.2_lw_ckd-definition_climate_fsck-tol0.0161.nc",
    "lw_rrtmgp": "ecckd-1.2_lw_ckd-definition_climate_rrtmgp-tol0.061.nc",
    "sw_wide": "ecckd-1.2_sw_ckd-definition_climate_wide-tol0.05.nc",
}


def inputs(ncol=5, nlay=12):
    """Hand-built columns that hit every branch: pressures below/above the table, temperatures
    off both ends, h2o below mole_fraction(1) and above the top, ch4/n2o on both sides of their
    reference, a Planck temperature below 120 K and one above 350 K."""
    rng = np.random.RandomState(20260101)
    psfc = np.array([101325.0, 85000.0, 120000.0, 60000.0, 100000.0])[:ncol]
    s = (np.arange(nlay + 1) / nlay) ** 2.2
    plev = 0.3 + (psfc[:, None] - 0.3) * s[None, :]             # top level 0.3 Pa < press_min, col 2 bottom > press_max
    plev[:, 1] = 0.6                                            # thin top layer wholly above the table (ip0 clamps to 1)
    play = 0.5 * (plev[:, 1:] + plev[:, :-1])
    tlay = 215.0 + 70.0 * (play / psfc[:, None]) ** 0.3 + rng.uniform(-25, 25, size=play.shape)
    tlay[0, 0] = 100.0                                          # below every table and below T_planck(1)
    tlay[1, -1] = 365.0                                         # above T_planck(end)
    tlev = np.empty_like(plev)
    tlev[:, 1:-1] = 0.5 * (tlay[:, 1:] + tlay[:, :-1]); tlev[:, 0] = tlay[:, 0]; tlev[:, -1] = tlay[:, -1]
    tlev[3, 4] = 119.5
    tsfc = tlev[:, -1] + np.array([1.5, -2.0, 0.0, 3.0, 80.0])[:ncol]
    h2o = np.clip(2e-2 * (play / psfc[:, None]) ** 3.5, 0, None) * rng.uniform(0.5, 2.0, size=play.shape)
    h2o[0, :3] = 1e-9                                           # below mole_fraction(1)
    h2o[4, -2:] = 0.2                                           # above the LUT top
    o3 = 8e-6 * np.exp(-0.5 * np.log(play / 1000.0) ** 2) + 2e-8
    f = lambda a: np.asfortranarray(a)
    gases = [("co2", 420e-6), ("ch4", 1.0e-6), ("n2o", 4.0e-7), ("o2", 0.209), ("cfc11", 2.3e-10),
             ("cfc12", 5.2e-10), ("h2o", f(h2o)), ("o3", f(o3)), ("no2", 0.0)]
    return f(plev), f(tlay), f(tlev), tsfc, gases


def main():
    plev, tlay, tlev, tsfc, gases = inputs()
    save = dict(plev=plev, tlay=tlay, tlev=tlev, tsfc=tsfc, h2o=dict(gases)["h2o"], o3=dict(gases)["o3"])
    np.savez_compressed(os.path.join(HERE, "inputs.npz"), **save)
    for key, fn in FILES.items():
        m = ckd_model.load_and_init(os.path.join(ROOT, "data", fn))
        co = COracle(m)
        if m.shortwave:
            out, err = co.gas_optics_ext(plev, tlay, gases)
        else:
            out, err = co.gas_optics_int(plev, tlay, tsfc, gases, tlev)
        assert err == ""
        idx = co.indices("h2o", plev, tlay, dict(gases)["h2o"])
        np.savez_compressed(os.path.join(HERE, f"{key}.npz"), **out, **{f"idx_{k}": v for k, v in idx.items()})
        print(key, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
