"""Summarise an `ncu --page source --csv` dump: per-SASS-instruction executed counts and stall samples.
usage: ncu_src.py src.csv ncells [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ncell = float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ci = {n: i for i, n in enumerate(hdr)}
tot_inst = sum(float(r[ci["Instructions Executed"]]) for r in data)
tot_samp = sum(float(r[ci["# Samples"]]) for r in data)
print(f"SASS instrs {len(data)}  executed warp-instrs {tot_inst:.3e} ({tot_inst/ncell:.1f}/cell)  samples {tot_samp:.0f}")
# opcode histogram by executed count
op = collections.Counter(); ops = collections.Counter()
for r in data:
    s = r[ci["Source"]].strip()
    if s.startswith("@"): s = s.split(None, 1)[1]
    o = s.split()[0].split(".")[0]
    op[o] += float(r[ci["Instructions Executed"]]); ops[o] += float(r[ci["# Samples"]])
print("opcode: exec/cell  samples%")
for o, c in op.most_common(28):
    print(f"  {o:10s} {c/ncell:8.1f}  {100*ops[o]/tot_samp:5.1f}%")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
agg = {n: sum(float(r[ci[n]]) for r in data) for n in stalls}
print("stall totals:", {k[6:]: f"{100*v/tot_samp:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot_samp})
print(f"top {top} instructions by samples:")
for idx, r in sorted(enumerate(data), key=lambda ir: -float(ir[1][ci["# Samples"]]))[:top]:
    st = {n[6:]: int(float(r[ci[n]])) for n in stalls if float(r[ci[n]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"  #{idx:4d} {float(r[ci['# Samples']])/tot_samp*100:5.2f}% exec/cell {float(r[ci['Instructions Executed']])/ncell:6.2f}  {r[ci['Source']].strip()[:70]:70s} {st}")
