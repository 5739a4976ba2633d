"""Split an `ncu --page source --csv` dump of quad_kernel into warp roles (regions delimited by USETMAXREG) and
summarise executed instructions and stall samples per role.  usage: ncu_roles.py src.csv ncells"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ncell = float(sys.argv[2])
ci = {n: i for i, n in enumerate(hdr)}
cuts = [i for i, r in enumerate(data) if "USETMAXREG" in r[ci["Source"]]]
names = ["setup", "helpers(store+planck)", "index", "compute"]
bounds = [0] + cuts + [len(data)]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(float(r[ci["# Samples"]]) for r in data)
for k in range(len(bounds) - 1):
    seg = data[bounds[k]:bounds[k + 1]]
    ins = sum(float(r[ci["Instructions Executed"]]) for r in seg)
    smp = sum(float(r[ci["# Samples"]]) for r in seg)
    agg = {n[6:]: sum(float(r[ci[n]]) for r in seg) for n in stalls}
    top = {k2: f"{100*v/max(smp,1):.0f}%" for k2, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]}
    print(f"{names[k] if k < len(names) else k:24s} instr/cell {ins/ncell:6.1f}  samples {100*smp/tot:5.1f}%  {top}")
