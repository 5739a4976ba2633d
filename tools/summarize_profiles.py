"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep>   (the .ncu-rep holds one LW and one SW quad_kernel launch)"""
import csv, json, os, subprocess, sys, collections
tag, launches, rep = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel totals and share of the GPU time -----------------------------------
rows = [r for r in csv.reader(open(launches)) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]; ci = {n: i for i, n in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr_i + 1:]:
    if len(r) <= ci["Metric Value"]: continue
    name = r[ci["Kernel Name"]]
    val = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    agg[name][0] += 1; agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches_summary.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras`\n")
    f.write("(cold-cache, serialised per-launch times: compare SHARES, not absolutes; includes input generation and warm-up)\n\n")
    f.write(f"{'launches':>8} {'total ms':>10} {'avg us':>10} {'share':>7}  kernel\n")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        f.write(f"{n:8d} {ns/1e6:10.3f} {ns/n/1e3:10.1f} {100*ns/tot:6.2f}%  {name[:150]}\n")
print(open(os.path.join(out_dir, f"{tag}_launches_summary.txt")).read())

# ---- full capture: headline counters of each profiled walk-kernel launch -----------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, units = rr[0], rr[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
traffic = {}
with open(os.path.join(out_dir, f"{tag}_quad_kernel_ncu.txt"), "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:quad_kernel --launch-skip 20 -c 2 (bench.py blocks of 65536 columns x 137 layers, fp64)\n\n")
    for r in rr[2:]:
        d = dict(zip(h, r)); u = dict(zip(h, units))
        nm = d.get("Kernel Name", "")
        f.write(f"== {nm[:160]}\n")
        for k in keys[1:]:
            if k in d: f.write(f"   {k:70s} {d[k]:>20s} {u.get(k,'')}\n")
        def to_bytes(k):
            v = float(d[k].replace(",", "")); return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[k]]
        tot_b = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        ncol = 65536
        import re
        mt = re.search(r"quad_kernel<\w+, (?:\(bool\))?(\d)", nm)
        band = "sw" if (mt and mt.group(1) == "1") else "lw"
        traffic[f"{band}_dram_bytes_per_column"] = tot_b / ncol
        ncell = ncol * 137
        f.write(f"   dram read+write per launch: {tot_b/1e9:.3f} GB  ({tot_b/ncol:.0f} B/column)\n")
        try:
            f.write(f"   warp-instructions per (column, layer): {float(d['smsp__inst_executed.sum'].replace(',', ''))/ncell:.1f};  "
                    f"shared wavefronts per cell: {float(d['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'].replace(',', ''))/ncell:.1f} "
                    f"(of which bank conflicts {float(d['l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum'].replace(',', ''))/ncell:.1f});  "
                    f"global-load sectors per cell: {float(d['l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum'].replace(',', ''))/ncell:.1f}\n\n")
        except Exception:
            f.write("\n")
traffic["source"] = f"profiles/{tag}_quad_kernel_ncu.txt (ncu --set full, one launch of 65536 columns each)"
json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)
print(open(os.path.join(out_dir, f"{tag}_quad_kernel_ncu.txt")).read())
