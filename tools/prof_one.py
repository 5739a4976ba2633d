"""Run one kernel variant a few times (for ncu).  usage: prof_one.py lw|sw f64|f32 walk|generic [ncol] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
import numpy as np, torch
import ecckd_b200
from ecckd_b200 import synthetic
import __graft_entry__ as ge

band, prec, kern = sys.argv[1:4]
ncol = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
dev = torch.device("cuda:0")
dt = torch.float64 if prec == "f64" else torch.float32
atm = synthetic.make_atmosphere(ncol, 137, device=dev)
if dt != torch.float64:
    atm = atm.astype(dt)
go = ecckd_b200.GasOpticsEcckd.load(ge.LW_FILE if band == "lw" else ge.SW_FILE)
lw = band == "lw"
gases = ecckd_b200.PreparedGases(atm.gas_list(), ncol, 137, np.float64 if prec == "f64" else np.float32, dev)
out = None
k = {"walk": ecckd_b200.KERNEL_WALK, "generic": ecckd_b200.KERNEL_GENERIC}[kern]
ts = []
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, gases, tlev=atm.tlev if lw else None, kernel=k, out=out)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = sorted(ts[1:]) if len(ts) > 1 else ts
ng = go.get_ngpt(); w = 8 if prec == "f64" else 4
byts = w * ncol * ((2 * 138 + 3 * 137 + 1) + (4 * 137 * ng + ng)) if lw else w * ncol * ((138 + 3 * 137) + (3 * 137 * ng + ng))
med = ts[len(ts) // 2]
print(f"{band} {prec} {kern} ncol={ncol}: min {ts[0]:.3f} ms  median {med:.3f} ms  ({byts / med / 1e6 / 6550.1 * 100:.1f}% of 6550 GB/s at median)")
