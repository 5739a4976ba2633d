"""Latency of RFMIP-shaped calls (BASELINE configs 1-3: 1800 columns x 60 layers; and the reference driver's block_size=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
import numpy as np, torch
import ecckd_b200
from ecckd_b200 import synthetic
import __graft_entry__ as ge
dev = torch.device("cuda:0")
for path in (ge.LW_FILE, ge.LW2_FILE, ge.SW_FILE):
    go = ecckd_b200.GasOpticsEcckd.load(path)
    lw = go.source_is_internal()
    for ncol in (1, 100, 1800):
        atm = synthetic.make_atmosphere(ncol, 60, device=dev)
        gases = ecckd_b200.PreparedGases(atm.gas_list(), ncol, 60, np.float64, dev)
        out = None
        for _ in range(20):
            out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, gases, tlev=atm.tlev if lw else None, out=out)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 200
        for _ in range(n):
            out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, gases, tlev=atm.tlev if lw else None, out=out)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"{os.path.basename(path)[10:48]:40s} ncol={ncol:5d} nlay=60: {dt*1e6:8.1f} us per call (back-to-back, incl. launch overhead)  {ncol*60*go.get_ngpt()/dt:.3e} gptl/s")
