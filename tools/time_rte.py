"""Time the device flux solvers (row N1) on one launch-sized block.  usage: time_rte.py [ncol]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
import numpy as np, torch
import ecckd_b200
from ecckd_b200 import synthetic
import __graft_entry__ as ge
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda:0")
atm = synthetic.make_atmosphere(ncol, 137, device=dev)
gases = ecckd_b200.PreparedGases(atm.gas_list(), ncol, 137, np.float64, dev)
def t(f, n=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
lw = ecckd_b200.GasOpticsEcckd.load(ge.LW_FILE); sw = ecckd_b200.GasOpticsEcckd.load(ge.SW_FILE)
ol = lw.gas_optics(None, atm.plev, atm.tlay, atm.tsfc, gases, tlev=atm.tlev)
osw = sw.gas_optics(None, atm.plev, atm.tlay, None, gases)
print(f"rte_lw {ncol} columns: {t(lambda: lw.rte_lw(ol)):.3f} ms   (reads {6 * ncol * 137 * 32 * 8 / 1e9:.2f} GB)")
print(f"rte_sw {ncol} columns: {t(lambda: sw.rte_sw(osw)):.3f} ms   (moves {14 * ncol * 138 * 27 * 8 / 1e9:.2f} GB)")
