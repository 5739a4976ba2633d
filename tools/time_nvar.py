"""Time the varying-gas instantiations of the quad kernel on one launch-sized block (developer probe).
usage: time_nvar.py [ncol]   -- prints LW / SW ms for nvar = 1 (default call), 2, 3, 4 and 7"""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
import numpy as np, torch
import ecckd_b200
from ecckd_b200 import synthetic
import __graft_entry__ as ge

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
atm = synthetic.make_atmosphere(ncol, 137, device=dev)
lw = ecckd_b200.GasOpticsEcckd.load(ge.LW_FILE); sw = ecckd_b200.GasOpticsEcckd.load(ge.SW_FILE)
wob = 1.0 + 0.2 * torch.cos(torch.arange(ncol * 137, device=dev, dtype=torch.float64)).reshape(137, ncol).t()
def timed(go, gases, lw_):
    out, ts = None, []
    for r in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw_ else None, gases, tlev=atm.tlev if lw_ else None, out=out)
        e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    return statistics.mean(ts)
cases = [("nvar1", ()), ("nvar2", ("co2",)), ("nvar3", ("co2", "ch4")), ("nvar4", ("co2", "ch4", "n2o")), ("nvar7", ("co2", "ch4", "n2o", "o2", "cfc11", "cfc12"))]
res = []
for label, names in cases:
    gl = [(n, (v if torch.is_tensor(v) else (ecckd_b200.to_fortran(v * wob, torch.float64, dev) if n in names else v))) for n, v in atm.gas_list()]
    gv = ecckd_b200.PreparedGases(gl, ncol, 137, np.float64, dev)
    res.append("%s %.3f/%.3f" % (label, timed(lw, gv, True), timed(sw, gv, False)))
print("LW/SW ms:", "  ".join(res))
