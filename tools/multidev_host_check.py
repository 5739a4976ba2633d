"""One process, a model replicated on every visible GPU: the host-array entry points shard the columns over the
devices (no collective) and must reproduce the single-device result bitwise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "rte-ecckd_b200")]
import numpy as np, torch
import ecckd_b200
from ecckd_b200 import synthetic
import __graft_entry__ as ge
ndev = torch.cuda.device_count()
print("devices:", ndev)
atm = synthetic.make_atmosphere(1003, 137)
F = lambda t: np.asfortranarray(t.numpy())
gl = [(n, F(v) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
for path, lw in ((ge.LW_FILE, True), (ge.SW_FILE, False)):
    one = ecckd_b200.GasOpticsEcckd.load(path, devices=[0])
    many = ecckd_b200.GasOpticsEcckd.load(path, devices=list(range(ndev)))
    assert many.get_ndev() == ndev
    args = (None, F(atm.plev), F(atm.tlay), atm.tsfc.numpy() if lw else None, gl)
    kw = dict(tlev=F(atm.tlev)) if lw else {}
    a = one.gas_optics(*args, **kw)
    for bc in (0, 100, 333):
        b = many.gas_optics(*args, block_cols=bc, **kw)
        for k in a:
            assert np.array_equal(a[k], b[k]), (path, k, bc)
    print(os.path.basename(path)[10:20], "sharded over", ndev, "devices == single device (bitwise), block_cols 0/100/333")
print("OK")
