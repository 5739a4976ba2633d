"""Developer probe (not a test, not the bench): parity of every kernel variant against the
oracle on a small case, then CUDA-event timings of the walk/generic kernels on a larger block."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
import numpy as np, torch
import ecckd_b200
from ecckd_b200 import synthetic
from oracle import ckd_model
from oracle.c_oracle import COracle
import __graft_entry__ as ge

dev = torch.device("cuda:0")
def rel(a, b):
    return ge._rel_err(a, b)

def parity(path, ncol, nlay, all2d=False):
    atm = synthetic.make_atmosphere(ncol, nlay)
    glist = atm.gas_list()
    if all2d:   # every gas as a (ncol, nlay) field with some variation: exercises the NB=8 walk kernel
        glist = [(n, v if torch.is_tensor(v) else ecckd_b200.to_fortran(v * (1.0 + 0.3 * torch.sin(torch.arange(ncol * nlay, dtype=torch.float64)).reshape(ncol, nlay)), torch.float64)) for n, v in glist]
    gl_cpu = [(n, v.numpy() if torch.is_tensor(v) else v) for n, v in glist]
    F = lambda t: np.asfortranarray(t.numpy())
    m = ckd_model.load_and_init(path)
    co = COracle(m)
    go = ecckd_b200.GasOpticsEcckd.load(path)
    lw = not m.shortwave
    if lw:
        ref, _ = co.gas_optics_int(F(atm.plev), F(atm.tlay), atm.tsfc.numpy(), gl_cpu, F(atm.tlev))
    else:
        ref, _ = co.gas_optics_ext(F(atm.plev), F(atm.tlay), gl_cpu)
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 1e-5)):
        D = lambda t: ecckd_b200.to_fortran(t, dt, dev)
        gl = [(n, D(v) if torch.is_tensor(v) else v) for n, v in glist]
        for kern, kname in ((ecckd_b200.KERNEL_WALK, "walk"), (ecckd_b200.KERNEL_GENERIC, "generic")):
            out = go.gas_optics(None, D(atm.plev), D(atm.tlay), atm.tsfc.to(dev, dt) if lw else None, gl,
                                tlev=D(atm.tlev) if lw else None, kernel=kern)
            torch.cuda.synchronize()
            errs = {k: rel(out[k].cpu().numpy(), v) for k, v in ref.items()}
            print(os.path.basename(path)[10:40], "all2d" if all2d else "", str(dt)[6:], kname, "OK" if max(errs.values()) <= tol else "FAIL",
                  {k: f"{e:.2e}" for k, e in errs.items()}, flush=True)

def timing(path, ncol, nlay, dt, kern, reps=5):
    atm = synthetic.make_atmosphere(ncol, nlay, device=dev)
    if dt != torch.float64:
        atm = atm.astype(dt)
    go = ecckd_b200.GasOpticsEcckd.load(path)
    lw = go.source_is_internal()
    npdt = np.float64 if dt == torch.float64 else np.float32
    gases = ecckd_b200.PreparedGases(atm.gas_list(), ncol, nlay, npdt, dev)
    out = None
    ts = []
    for r in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw else None, gases, tlev=atm.tlev if lw else None,
                            kernel=kern, out=out)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[2:]) * 1e-3
    ng = go.get_ngpt()
    w = 8 if dt == torch.float64 else 4
    if lw:
        bytes_ = w * ncol * (2 * (nlay + 1) + 3 * nlay + 1) + w * ncol * (4 * nlay * ng + ng)
    else:
        bytes_ = w * ncol * ((nlay + 1) + 3 * nlay) + w * ncol * (3 * nlay * ng + ng)
    print(f"{'LW' if lw else 'SW'} {str(dt)[6:]} kern={kern} ncol={ncol}: {t*1e3:.3f} ms  {ncol*nlay*ng/t:.3e} gptl/s  "
          f"{bytes_/t/1e9:.0f} GB/s ({bytes_/t/1e9/6550.1*100:.1f}% of measured HBM)", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for p in (ge.LW_FILE, ge.LW2_FILE, ge.SW_FILE):
        parity(p, 67, 61)
        parity(p, 35, 70, all2d=True)
    ncol = int(os.environ.get("NCOL", 65536))
    for p in (ge.LW_FILE, ge.SW_FILE):
        for dt in (torch.float64, torch.float32):
            for kern in (ecckd_b200.KERNEL_WALK, ecckd_b200.KERNEL_GENERIC):
                timing(p, ncol, 137, dt, kern)
