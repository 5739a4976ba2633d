#!/usr/bin/env python
"""RFMIP-RAD-IRF-shaped driver: the call sequence of the reference's example programs
(ecckd_rfmip_lw.F90:107-136, ecckd_rfmip_sw.F90:112-162) on the B200 path.

The RFMIP input file is not available offline, so the 18 experiments x 100 columns x 60 layers come from
ecckd_b200.synthetic.make_rfmip_like(); everything else follows the drivers: load the ck-definition, loop over
blocks of columns, gas_optics -> flux solver, TSI renormalisation and night-column masking for the shortwave.
The flux solvers are the test-side numpy ones (tests/rte_solvers.py; rte-rrtmgp is not available either).

    python examples/rfmip_like_driver.py [--block 100]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200"), os.path.join(ROOT, "tests")]
import ecckd_b200                      # noqa: E402
from ecckd_b200 import synthetic       # noqa: E402
import rte_solvers                     # noqa: E402

DATA = os.path.join(ROOT, "data")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--block", type=int, default=100, help="columns per gas_optics call (the reference uses 1)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lw = ecckd_b200.GasOpticsEcckd.load(os.path.join(DATA, "ecckd-1.2_lw_ckd-definition_climate_fsck-tol0.0161.nc"))
    sw = ecckd_b200.GasOpticsEcckd.load(os.path.join(DATA, "ecckd-1.2_sw_ckd-definition_climate_wide-tol0.05.nc"))
    D = lambda t: ecckd_b200.to_fortran(t, torch.float64, dev)
    tsi_target, mu0 = 1361.0, 0.6
    print("expt   OLR   LW dn sfc   SW dn sfc   SW up toa   [W m-2, mean over 100 columns]")
    for e, atm in enumerate(synthetic.make_rfmip_like()):
        olr, lwdn, swdn, swup = [], [], [], []
        for c0 in range(0, atm.ncol, args.block):
            sl = slice(c0, min(c0 + args.block, atm.ncol))
            gases = [(n, D(v[sl]) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
            plev, tlay, tlev = D(atm.plev[sl]), D(atm.tlay[sl]), D(atm.tlev[sl])
            o = lw.gas_optics(None, plev, tlay, atm.tsfc[sl].to(dev), gases, tlev=tlev)               # ecckd_rfmip_lw.F90:120
            n = {k: v.cpu().numpy() for k, v in o.items()}
            up, dn = rte_solvers.lw_fluxes(n["tau"], n["lay_source"], n["lev_source_inc"], n["lev_source_dec"], n["sfc_source"])
            olr.append(up[:, 0]); lwdn.append(dn[:, -1])
            o = sw.gas_optics(None, plev, tlay, None, gases)                                           # ecckd_rfmip_sw.F90:118
            n = {k: v.cpu().numpy() for k, v in o.items()}
            toa = n["toa_src"] * (tsi_target / n["toa_src"].sum(axis=1, keepdims=True))               # TSI renormalisation :126-133
            u, d, _ = rte_solvers.sw_fluxes(n["tau"], n["ssa"], n["g"], toa, mu0=mu0)
            swdn.append(d[:, -1]); swup.append(u[:, 0])
        m = lambda x: float(np.concatenate(x).mean())
        print(f"{e + 1:4d} {m(olr):7.2f} {m(lwdn):9.2f} {m(swdn):11.2f} {m(swup):11.2f}")


if __name__ == "__main__":
    main()
