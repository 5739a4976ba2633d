#!/usr/bin/env python
"""bench.py -- LW+SW ECCKD gas-optics throughput (g-point-layers/s) on B200, with the HBM roofline
fraction of the dominant kernel and the CPU restatement of the reference timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path over the workload: longwave (ecckd-1.2 fsck-tol0.0161, 32 g-points:
tau + Planck lay/lev/sfc sources) plus shortwave (wide-tol0.05, 27 g-points: tau/ssa/g + Rayleigh + TOA
source) gas optics of a synthetic GCM block of --cols columns x 137 layers per GPU (BASELINE.json
configs[3]: 1M columns on one B200; N GPUs = N x 1M columns, column-sharded, no collective).  Inputs are
resident in HBM when the timed region starts; outputs (140 + 89 GB per 1M columns in fp64, more than HBM)
go through a ring of two column-block output buffers, as a host model consumes them block by block.

`value`  : whole-job g-point-layers/s, device-resident arrays, CUDA events, max over ranks.
`e2e`    : the same rate through the host-array C-ABI entry points (ecckd_gas_optics_{lw,sw}_host_f64):
           pinned host inputs -> H2D -> kernels -> D2H of every output, all inside the timed region.
`roofline`: LW quad kernel (roofline_sw: SW; roofline_f32 / roofline_sw_f32: the fp32 variant of configs[3]),
           algorithmic bytes / CUDA-event duration vs the measured HBM copy bandwidth.
`cpu_baseline` / --impl reference: oracle/ecckd_oracle.c (-O3 -march=native, all host threads) -- a port:
           the Fortran reference cannot be built offline (no Fortran compiler here; `cpu_baseline.fortran` records
           whether the box running the bench has one -- oracle/ref is the recipe that would then pin and time it).
`extras`  : RFMIP-shaped call latencies (configs[0-2]: 1 / 100 / 1800 columns x 60 layers, device and host path, with the
           CPU port's time for the same call), calls with co2/ch4/n2o as profiles, the generic and column-walk kernels,
           the host path on pageable memory.
--total-cols N: BASELINE configs[4] as named -- N columns (6.6M) column-sharded over the GPUs, strong scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rte-ecckd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

DATA = os.path.join(ROOT, "data")
LW_FILE = os.path.join(DATA, "ecckd-1.2_lw_ckd-definition_climate_fsck-tol0.0161.nc")
SW_FILE = os.path.join(DATA, "ecckd-1.2_sw_ckd-definition_climate_wide-tol0.05.nc")
NLAY = 137
NG_LW, NG_SW = 32, 27
METRIC = "LW+SW gas-optics g-point-layers/s"
UNIT = "g-point-layers/s"


def algorithmic_bytes(ncol: int, nlay: int, ngpt: int, lw: bool, w: int = 8) -> int:
    """inputs read + optical properties and sources written (2 two-dimensional gases)."""
    if lw:
        return w * ncol * (2 * (nlay + 1) + 3 * nlay + 1) + w * ncol * (4 * nlay * ngpt + ngpt)
    return w * ncol * ((nlay + 1) + 3 * nlay) + w * ncol * (3 * nlay * ngpt + ngpt)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (device copy, measured)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per LW launch from the committed ncu capture, scaled per column."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU arm
_REBUILT = False


def cpu_port_rate(ncol: int, threads: int, block: int = 16, reps: int = 1):
    """Time the oracle's -O3 -march=native build on `ncol` synthetic columns (LW + SW), blocks of `block`
    columns handed to `threads` POSIX threads (the reference driver's loop over independent column blocks).
    Returns (g-point-layers/s, seconds, threads)."""
    import torch
    from ecckd_b200 import synthetic
    from oracle import ckd_model
    from oracle.c_oracle import COracle, load

    global _REBUILT
    if not _REBUILT:
        load(fast=True, rebuild=True)      # -march=native must match THIS host, not the box the repo was built on
        _REBUILT = True
    atm = synthetic.make_atmosphere(ncol, NLAY)
    nblocks = ncol // block
    ncol = nblocks * block

    def blocked(t, n2):   # (ncol, n2) column-major -> (nblocks, n2, block) C order == Fortran (block, n2, nblocks)
        a = t.numpy()[:ncol]
        return np.ascontiguousarray(a.reshape(nblocks, block, n2).transpose(0, 2, 1))

    plev, tlay, tlev = blocked(atm.plev, NLAY + 1), blocked(atm.tlay, NLAY), blocked(atm.tlev, NLAY + 1)
    tsfc = np.ascontiguousarray(atm.tsfc.numpy()[:ncol])
    gases = [(n, blocked(v, NLAY) if torch.is_tensor(v) else v) for n, v in atm.gas_list()]
    lw = COracle(ckd_model.load_and_init(LW_FILE), fast=True)
    sw = COracle(ckd_model.load_and_init(SW_FILE), fast=True)
    if threads <= 0:
        threads = lw.max_threads()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        lw.run_blocks(True, block, NLAY, nblocks, plev, tlay, tsfc, tlev, gases, threads)
        sw.run_blocks(False, block, NLAY, nblocks, plev, tlay, None, None, gases, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ncol * NLAY * (NG_LW + NG_SW) / best, best, threads


def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: the reference's CPU path (port) on the box's host cores, same metric / config."""
    if rank != 0:
        return
    ncol = args.ref_cols
    for _ in range(args.warmup):
        cpu_port_rate(min(ncol, 2048), 0)
    times, threads = [], 0
    t_all = time.perf_counter()
    for _ in range(args.steps):
        rate, dt, threads = cpu_port_rate(ncol, 0)
        times.append(dt)
    total = time.perf_counter() - t_all
    ms = 1e3 * sum(times) / len(times)
    value = ncol * NLAY * (NG_LW + NG_SW) / (ms * 1e-3)
    sample = f"{ncol} synthetic columns x {NLAY} layers per step (LW 32 + SW 27 g-points), blocks of 16 columns over {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic GCM block, columns x 137 layers, LW+SW (bounded sample of BASELINE.json configs[3])",
                   "columns_per_step": ncol, "nlay": NLAY, "ngpt_lw": NG_LW, "ngpt_sw": NG_SW,
                   "note": "CPU restatement (oracle/ecckd_oracle.c, -O3 -march=native) of src/gas_optics_ecckd.f90; the Fortran "
                           "reference itself cannot be built offline (no Fortran compiler / netCDF-Fortran / rte-rrtmgp)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "fortran": fortran_probe()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": total,
    }
    print(json.dumps(line))


def bind_near_gpu(local_rank: int, world: int) -> dict:
    """Pin this rank's host threads to the NUMA node its GPU hangs off (sysfs: the GPU's PCI device -> numa_node -> cpulist),
    or, when the topology is not exposed, to an equal slice of the cores.  Pinned host buffers are then first-touched on
    that node: with every rank allocating on node 0 the e2e (host-array) path of 8 ranks shared one memory controller."""
    info = {"numa_node": None, "cpus": None}
    if world <= 1:          # one rank: leave the process unbound (the CPU baseline then has every host core)
        return info
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        node = None
        if bus is None:
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True).stdout.strip()
            bus = out.lower().replace("00000000:", "0000:") if out else None
        if bus:
            path = f"/sys/bus/pci/devices/{str(bus).lower()}/numa_node"
            if os.path.exists(path):
                node = int(open(path).read().strip())
        cpus = None
        if node is not None and node >= 0:
            cl = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
            cpus = set()
            for part in cl.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
        if not cpus and world > 1:
            allc = sorted(os.sched_getaffinity(0))
            per = max(1, len(allc) // world)
            cpus = set(allc[local_rank * per:(local_rank + 1) * per]) or set(allc)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"numa_node": node, "cpus": len(cpus)}
    except Exception as exc:
        info["error"] = str(exc)
    return info


def fortran_probe() -> dict:
    """Is there a Fortran compiler on THIS box (probe at run time)?  With one, `make -C oracle/ref` builds the
    unmodified reference against stub rte-rrtmgp modules -- but only next to /root/reference, which the GPU box lacks."""
    from oracle.ref import run_ref
    fc = run_ref.fortran_compiler()
    return {"compiler": fc, "reference_tree": os.path.isdir(run_ref.REFERENCE), "ref_driver_built": os.path.exists(run_ref.DRIVER)}


def rfmip_latencies(dev) -> dict:
    """BASELINE configs[0-2] (RFMIP-shaped synthetic: the RFMIP file is not available offline): microseconds per
    gas_optics call of 1 / 100 / 1800 columns x 60 layers for the three ck files -- device arrays (CUDA events over
    back-to-back calls), host arrays through the host entry points (wall clock), and the CPU port on one thread (the
    reference drivers run block_size = 1 on one thread, ecckd_rfmip_lw.F90:39,107) for the same call."""
    import torch
    import ecckd_b200
    from ecckd_b200 import synthetic
    from oracle import ckd_model
    from oracle.c_oracle import COracle
    files = {"lw_fsck-tol0.0161": LW_FILE, "lw_rrtmgp-tol0.061": os.path.join(DATA, "ecckd-1.2_lw_ckd-definition_climate_rrtmgp-tol0.061.nc"),
             "sw_wide-tol0.05": SW_FILE}
    exps = synthetic.make_rfmip_like()                       # 18 experiments x 100 columns x 60 layers
    out = {}
    for label, path in files.items():
        go = ecckd_b200.GasOpticsEcckd.load(path, devices=[dev.index or 0])
        lw_ = go.source_is_internal()
        co = COracle(ckd_model.load_and_init(path), fast=True)
        res = {}
        for ncol in (1, 100, 1800):
            if ncol <= 100:
                a = exps[5]
                sel = lambda t: t[:ncol].contiguous() if t.dim() == 1 else ecckd_b200.to_fortran(t[:ncol], torch.float64)
                plev, tlay, tlev, tsfc = sel(a.plev), sel(a.tlay), sel(a.tlev), sel(a.tsfc)
                gl = [(n, sel(v) if torch.is_tensor(v) else v) for n, v in a.gas_list()]
            else:                                            # all experiments in one call: scalars differ per experiment -> use experiment 5's
                cat = lambda k: ecckd_b200.to_fortran(torch.cat([getattr(e, k) for e in exps], 0), torch.float64)
                plev, tlay, tlev = cat("plev"), cat("tlay"), cat("tlev")
                tsfc = torch.cat([e.tsfc for e in exps], 0)
                gl = [(n, (cat(n) if torch.is_tensor(v) else v)) for n, v in exps[5].gas_list()]
            F = lambda t: np.asfortranarray(t.numpy())
            gl_np = [(n, F(v) if torch.is_tensor(v) else v) for n, v in gl]
            D = lambda t: ecckd_b200.to_fortran(t, torch.float64, dev)
            gd = ecckd_b200.PreparedGases([(n, D(v) if torch.is_tensor(v) else v) for n, v in gl], ncol, 60, np.float64, dev)
            dp, dt_, dl, ds = D(plev), D(tlay), D(tlev), tsfc.to(dev)
            o = None
            call = lambda: go.gas_optics(None, dp, dt_, ds if lw_ else None, gd, tlev=dl if lw_ else None, out=o)
            for _ in range(20):
                o = call()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_rep = 200
            e0.record()
            for _ in range(n_rep):
                o = call()
            e1.record(); torch.cuda.synchronize()
            dev_us = e0.elapsed_time(e1) * 1e3 / n_rep
            gh = ecckd_b200.PreparedGases(gl_np, ncol, 60, np.float64, None)
            hp, ht, hl, hs = F(plev), F(tlay), F(tlev), tsfc.numpy()
            oh = None
            hcall = lambda: go.gas_optics(None, hp, ht, hs if lw_ else None, gh, tlev=hl if lw_ else None, out=oh)
            for _ in range(5):
                oh = hcall()
            w0 = time.perf_counter()
            for _ in range(50):
                oh = hcall()
            host_us = (time.perf_counter() - w0) * 1e6 / 50
            w0 = time.perf_counter()
            n_cpu = 3 if ncol > 100 else 20
            for _ in range(n_cpu):
                (co.gas_optics_int(hp, ht, hs, gl_np, hl) if lw_ else co.gas_optics_ext(hp, ht, gl_np))
            cpu_us = (time.perf_counter() - w0) * 1e6 / n_cpu
            res[f"ncol{ncol}"] = {"device_us": dev_us, "host_path_us": host_us, "cpu_port_1thread_us": cpu_us}
        out[label] = res
        go.close()
    out["note"] = ("per gas_optics call through the Python binding; device_us = CUDA events over 200 back-to-back calls on device arrays; "
                   "host_path_us = wall clock of the synchronous host entry point on pageable numpy arrays (H2D + kernels + D2H); "
                   "cpu_port_1thread_us = oracle/ecckd_oracle.c -O3 on one thread, as the reference drivers run")
    return out


# ------------------------------------------------------------------------------------------ GPU arm
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cols", type=int, default=1 << 20, help="columns per GPU per step (BASELINE configs[3]: 1M)")
    ap.add_argument("--block", type=int, default=1 << 16, help="columns per kernel launch / output ring slot")
    ap.add_argument("--e2e-cols", type=int, default=1 << 17, help="columns per GPU per e2e step (host-array path)")
    ap.add_argument("--e2e-block", type=int, default=1 << 14)
    ap.add_argument("--cpu-cols", type=int, default=1 << 19, help="bounded sample for cpu_baseline")
    ap.add_argument("--ref-cols", type=int, default=1 << 15, help="columns per step of --impl reference")
    ap.add_argument("--total-cols", type=int, default=0,
                    help="strong scaling: this many columns in total, column-sharded over the GPUs (configs[4]: 6600000)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    rank, local_rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import ecckd_b200
    from ecckd_b200 import sharding, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices: the product path has no CPU fallback")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_near_gpu(local_rank, world)     # before any pinned allocation: first-touch puts the e2e buffers on that node

    if args.total_cols > 0:                     # strong scaling: contiguous shards of one global grid (configs[4])
        col_base, c1 = sharding.shard_columns(args.total_cols, world, rank)
        ncol = c1 - col_base
        ncol_all = args.total_cols
    else:                                       # weak scaling: every GPU owns its own `cols` columns of the global grid
        ncol, col_base, ncol_all = args.cols, rank * args.cols, args.cols * world
    B = min(args.block, ncol)
    nblk = (ncol + B - 1) // B
    lw = ecckd_b200.GasOpticsEcckd.load(LW_FILE, devices=[local_rank])
    sw = ecckd_b200.GasOpticsEcckd.load(SW_FILE, devices=[local_rank])

    # ---- inputs resident in HBM, already blocked like the reference driver blocks them -------------
    blocks = []
    for b in range(nblk):
        nb = min(B, ncol - b * B)
        atm = synthetic.make_atmosphere(nb, NLAY, col0=col_base + b * B, device=dev)
        blocks.append((atm, ecckd_b200.PreparedGases(atm.gas_list(), nb, NLAY, np.float64, dev)))
    FE = ecckd_b200.fortran_empty
    ring = [{"lw": {k: FE((B, NLAY, NG_LW), torch.float64, dev) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec")},
             "sw": {k: FE((B, NLAY, NG_SW), torch.float64, dev) for k in ("tau", "ssa", "g")}} for _ in range(2)]
    for r in ring:
        r["lw"]["sfc_source"] = FE((B, NG_LW), torch.float64, dev)
        r["sw"]["toa_src"] = FE((B, NG_SW), torch.float64, dev)

    def outs(slot, band, nb):
        o = ring[slot][band]
        return o if nb == B else {k: FE(tuple([nb] + list(v.shape[1:])), torch.float64, dev) for k, v in o.items()}

    ev_pairs = {"lw": [], "sw": []}

    def step(timed: bool):
        for b, (atm, gases) in enumerate(blocks):
            slot = b & 1
            if timed:
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
            lw.gas_optics(None, atm.plev, atm.tlay, atm.tsfc, gases, tlev=atm.tlev, out=outs(slot, "lw", atm.ncol))
            if timed:
                e[1].record()
            sw.gas_optics(None, atm.plev, atm.tlay, None, gases, out=outs(slot, "sw", atm.ncol))
            if timed:
                e[2].record()
                ev_pairs["lw"].append((e[0], e[1], atm.ncol)); ev_pairs["sw"].append((e[1], e[2], atm.ncol))

    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    sharding.barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    torch.cuda.synchronize()
    sharding.barrier()
    clocks = sampler.stop() if sampler else None
    ms_step = sharding.max_over_ranks(t0.elapsed_time(t1) / args.steps, dev)
    gptl_step = ncol_all * NLAY * (NG_LW + NG_SW)
    value = gptl_step / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel (LW walk): live CUDA-event durations of every timed launch -
    peak, peak_src = measured_peak()
    full = [(a.elapsed_time(bb), n) for a, bb, n in ev_pairs["lw"] if n == B]
    lw_ms = statistics.mean(t for t, _ in full) if full else float("nan")
    bytes_lw = algorithmic_bytes(B, NLAY, NG_LW, True)
    achieved = bytes_lw / (lw_ms * 1e-3) / 1e9
    traffic = profiled_traffic()
    roofline = {"bound": "hbm", "kernel": "ecckd::quad_kernel<double, LW> (+ per-call table-mix kernel, <1% of the time)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic["lw_dram_bytes_per_column"] * B if traffic else None),
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed ncu --set full capture "
                                  "(profiles/traffic.json), scaled per column -- not re-measured in this run",
                "algorithmic_bytes_per_launch": bytes_lw, "columns_per_launch": B, "ms_per_launch": lw_ms, "peak_source": peak_src}
    fsw = [(a.elapsed_time(bb), n) for a, bb, n in ev_pairs["sw"] if n == B]
    sw_ms = statistics.mean(t for t, _ in fsw) if fsw else float("nan")
    bytes_sw = algorithmic_bytes(B, NLAY, NG_SW, False)
    roofline_sw = {"bound": "hbm", "kernel": "ecckd::quad_kernel<double, SW>", "achieved": bytes_sw / (sw_ms * 1e-3) / 1e9, "peak": peak,
                   "unit": "GB/s", "frac": bytes_sw / (sw_ms * 1e-3) / 1e9 / peak, "traffic": (traffic["sw_dram_bytes_per_column"] * B if traffic else None),
                   "ms_per_launch": sw_ms}
    launches = args.steps * nblk * 4          # per block: LW mix + LW walk + SW mix + SW walk

    # ---- the fp32 variant of configs[3] ("fp64 and fp32"): same blocks, same kernels, CUDA events per launch ----------
    extras = {}
    roofline_f32 = roofline_sw_f32 = None
    try:
        atm32 = blocks[0][0].astype(torch.float32)
        g32 = ecckd_b200.PreparedGases(atm32.gas_list(), atm32.ncol, NLAY, np.float32, dev)
        o_lw = o_sw = None
        t_lw, t_sw = [], []
        for rep_ in range(3 + 5):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            o_lw = lw.gas_optics(None, atm32.plev, atm32.tlay, atm32.tsfc, g32, tlev=atm32.tlev, out=o_lw)
            ev[1].record()
            o_sw = sw.gas_optics(None, atm32.plev, atm32.tlay, None, g32, out=o_sw)
            ev[2].record()
            torch.cuda.synchronize()
            if rep_ >= 3:
                t_lw.append(ev[0].elapsed_time(ev[1])); t_sw.append(ev[1].elapsed_time(ev[2]))
        m_lw, m_sw = statistics.mean(t_lw), statistics.mean(t_sw)
        b_lw, b_sw = algorithmic_bytes(atm32.ncol, NLAY, NG_LW, True, 4), algorithmic_bytes(atm32.ncol, NLAY, NG_SW, False, 4)
        roofline_f32 = {"bound": "hbm", "kernel": "ecckd::quad_kernel<float, LW>", "achieved": b_lw / (m_lw * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": b_lw / (m_lw * 1e-3) / 1e9 / peak, "traffic": None, "ms_per_launch": m_lw, "columns_per_launch": atm32.ncol,
                        "algorithmic_bytes_per_launch": b_lw, "gptl_per_s": atm32.ncol * NLAY * NG_LW / (m_lw * 1e-3)}
        roofline_sw_f32 = {"bound": "hbm", "kernel": "ecckd::quad_kernel<float, SW>", "achieved": b_sw / (m_sw * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": b_sw / (m_sw * 1e-3) / 1e9 / peak, "traffic": None, "ms_per_launch": m_sw, "columns_per_launch": atm32.ncol,
                           "algorithmic_bytes_per_launch": b_sw, "gptl_per_s": atm32.ncol * NLAY * NG_SW / (m_sw * 1e-3)}
        extras["f32"] = {"columns": atm32.ncol, "lw_ms": m_lw, "sw_ms": m_sw, "gptl_per_s": atm32.ncol * NLAY * (NG_LW + NG_SW) / ((m_lw + m_sw) * 1e-3)}
        del atm32, g32, o_lw, o_sw
    except Exception as exc:   # extras must never break the headline line
        extras["f32_error"] = str(exc)

    # ---- extras: other instantiations and kernels on one launch-sized block (fp64, CUDA events, mean of 3) -----------
    def timed_call(go, atm, gases, lw_, kernel=ecckd_b200.KERNEL_AUTO, reps=3):
        out, ts = None, []
        for r_ in range(2 + reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = go.gas_optics(None, atm.plev, atm.tlay, atm.tsfc if lw_ else None, gases, tlev=atm.tlev if lw_ else None, kernel=kernel, out=out)
            e1.record(); torch.cuda.synchronize()
            if r_ >= 2:
                ts.append(e0.elapsed_time(e1))
        return statistics.mean(ts)
    if rank == 0 and not args.no_extras:
        try:
            atm0 = blocks[0][0]
            nb0 = atm0.ncol
            wob = 1.0 + 0.2 * torch.cos(torch.arange(nb0 * NLAY, device=dev, dtype=torch.float64)).reshape(NLAY, nb0).t()
            for label, names in (("co2_ch4_n2o_profiles", ("co2", "ch4", "n2o")), ("all_gases_as_fields", ("co2", "ch4", "n2o", "o2", "cfc11", "cfc12"))):
                gl = [(n, (v if torch.is_tensor(v) else (ecckd_b200.to_fortran(v * wob, torch.float64, dev) if n in names else v))) for n, v in atm0.gas_list()]
                gv = ecckd_b200.PreparedGases(gl, nb0, NLAY, np.float64, dev)
                extras[label] = {"lw_ms": timed_call(lw, atm0, gv, True), "sw_ms": timed_call(sw, atm0, gv, False), "columns": nb0,
                                 "note": "nvar = %d varying bilinear gases (+ the h2o look-up table)" % (len(names) + 1)}
                del gl, gv
            g0_ = blocks[0][1]
            extras["generic_kernel"] = {"lw_ms": timed_call(lw, atm0, g0_, True, ecckd_b200.KERNEL_GENERIC), "sw_ms": timed_call(sw, atm0, g0_, False, ecckd_b200.KERNEL_GENERIC), "columns": nb0}
            extras["column_walk_kernel"] = {"lw_ms": timed_call(lw, atm0, g0_, True, ecckd_b200.KERNEL_COLWALK), "sw_ms": timed_call(sw, atm0, g0_, False, ecckd_b200.KERNEL_COLWALK),
                                            "columns": nb0, "note": "round-1 product kernel, now the fallback for column counts without a 16-byte layer pitch"}
            atm_s = synthetic.make_atmosphere(nb0, NLAY, col0=col_base, device=dev, noise=0.0)
            g_s = ecckd_b200.PreparedGases(atm_s.gas_list(), nb0, NLAY, np.float64, dev)
            extras["smooth_atmosphere"] = {"lw_ms": timed_call(lw, atm_s, g_s, True), "sw_ms": timed_call(sw, atm_s, g_s, False), "columns": nb0,
                                           "note": "the same block without the specification's cell-by-cell white noise (N(0, 5 K) on tlay, LogN(0, 0.5) on h2o): "
                                                   "columns of a tile then share table rows and Planck-table banks; not the headline workload"}
            del atm_s, g_s
            nodd = nb0 - 1
            atm_odd = synthetic.make_atmosphere(nodd, NLAY, col0=col_base, device=dev)
            g_odd = ecckd_b200.PreparedGases(atm_odd.gas_list(), nodd, NLAY, np.float64, dev)
            extras["odd_ncol"] = {"lw_ms": timed_call(lw, atm_odd, g_odd, True), "sw_ms": timed_call(sw, atm_odd, g_odd, False), "columns": nodd,
                                  "note": "device-pointer call with an odd column count: no tensor map possible, column-walk kernel with writer warps"}
            del atm_odd, g_odd
        except Exception as exc:
            extras["kernels_error"] = str(exc)
        try:      # configs[2]: the other longwave file at the same block size (36 g-points: a full group of 32 and a group of 4)
            lw36 = ecckd_b200.GasOpticsEcckd.load(os.path.join(DATA, "ecckd-1.2_lw_ckd-definition_climate_rrtmgp-tol0.061.nc"))
            atm36, g36 = blocks[0][0], blocks[0][1]
            ms36, ng36 = timed_call(lw36, atm36, g36, True), lw36.get_ngpt()
            extras["lw_rrtmgp-tol0.061"] = {"lw_ms": ms36, "ngpt": ng36, "columns": atm36.ncol,
                                            "gptl_per_s": atm36.ncol * NLAY * ng36 / (ms36 * 1e-3),
                                            "frac_of_hbm_peak": algorithmic_bytes(atm36.ncol, NLAY, ng36, True, 8) / (ms36 * 1e-3) / 1e9 / peak,
                                            "note": "the g-points run as groups of 32 (grid.y): the second group carries 4 of 32"}
            del lw36
        except Exception as exc:
            extras["lw_rrtmgp_error"] = str(exc)
        try:
            extras["rfmip"] = rfmip_latencies(dev)
        except Exception as exc:
            extras["rfmip_error"] = str(exc)

    # free the device-resident workload before the host-array measurement
    del blocks, ring
    torch.cuda.empty_cache()

    # ---- e2e: host arrays through the C ABI's host entry points ------------------------------------
    e2e = None
    if not args.no_e2e:
        ne, EB = min(args.e2e_cols, ncol), min(args.e2e_block, args.e2e_cols)
        nhb = (ne + EB - 1) // EB
        pin = lambda t: t.cpu().pin_memory()
        hblocks = []
        for b in range(nhb):
            nb = min(EB, ne - b * EB)
            atm = synthetic.make_atmosphere(nb, NLAY, col0=col_base + b * EB, device=dev)
            # pinned host copies; numpy views of them are what a host model would hand to the shim
            h = {k: pin(getattr(atm, k).t().contiguous()).numpy().T for k in ("plev", "tlay", "tlev", "h2o", "o3")}
            h["tsfc"] = pin(atm.tsfc).numpy()
            gl = [(n, (h[n] if n in ("h2o", "o3") else v)) for n, v in atm.gas_list()]
            hblocks.append((nb, h, ecckd_b200.PreparedGases(gl, nb, NLAY, np.float64, None)))
        hp = lambda shape: torch.empty(shape[::-1], dtype=torch.float64).pin_memory().numpy().T
        hout = {"lw": {k: hp((EB, NLAY, NG_LW)) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec")},
                "sw": {k: hp((EB, NLAY, NG_SW)) for k in ("tau", "ssa", "g")}}
        hout["lw"]["sfc_source"] = hp((EB, NG_LW)); hout["sw"]["toa_src"] = hp((EB, NG_SW))
        h2d = d2h = 0
        for nb, h, _ in hblocks:
            h2d += 8 * nb * (2 * (NLAY + 1) + 3 * NLAY + 1) + 8 * nb * ((NLAY + 1) + 3 * NLAY) + 8 * 2 * 7
            d2h += 8 * nb * (4 * NLAY * NG_LW + NG_LW) + 8 * nb * (3 * NLAY * NG_SW + NG_SW)

        def e2e_step():
            for nb, h, gases in hblocks:
                ol = hout["lw"] if nb == EB else None
                osw = hout["sw"] if nb == EB else None
                lw.gas_optics(None, h["plev"], h["tlay"], h["tsfc"], gases, tlev=h["tlev"], out=ol, block_cols=EB)
                sw.gas_optics(None, h["plev"], h["tlay"], None, gases, out=osw, block_cols=EB)

        for _ in range(3):
            e2e_step()
        torch.cuda.synchronize()
        sharding.barrier()
        k_e2e = max(2, min(args.steps, 5))
        w0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()                       # host entry points are synchronous: results are in host memory on return
        torch.cuda.synchronize()
        dt = sharding.max_over_ranks((time.perf_counter() - w0) / k_e2e, dev)
        sharding.barrier()
        e2e = {"value": ne * NLAY * (NG_LW + NG_SW) * world / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "columns_per_step": ne, "block_cols": EB, "ms_per_step": dt * 1e3,
               "api": "ecckd_gas_optics_{lw,sw}_host_f64 (pinned host arrays, H2D + kernels + D2H per column block; the caller blocks "
                      "its columns like the reference driver, so every copy is contiguous)",
               "pcie_d2h_GBps": d2h / dt / 1e9, "host_binding": numa}
        if rank == 0 and not args.no_extras:
            try:    # the same step on PAGEABLE host arrays (what a Fortran host model passes; the library page-locks them per call)
                pg = lambda a: np.asfortranarray(np.array(a))
                pblocks = []
                for nb, h, _ in hblocks[:2]:
                    hpg = {k: pg(v) for k, v in h.items()}
                    gl = [(n, (hpg[n] if n in ("h2o", "o3") else v)) for n, v in synthetic.make_atmosphere(1, NLAY).gas_list()]
                    pblocks.append((nb, hpg, ecckd_b200.PreparedGases(gl, nb, NLAY, np.float64, None)))
                pout = {"lw": {k: pg(v) for k, v in hout["lw"].items()}, "sw": {k: pg(v) for k, v in hout["sw"].items()}}

                def pstep():
                    for nb, h, gases in pblocks:
                        lw.gas_optics(None, h["plev"], h["tlay"], h["tsfc"], gases, tlev=h["tlev"], out=pout["lw"] if nb == EB else None, block_cols=EB)
                        sw.gas_optics(None, h["plev"], h["tlay"], None, gases, out=pout["sw"] if nb == EB else None, block_cols=EB)
                pstep()
                w0 = time.perf_counter()
                for _ in range(2):
                    pstep()
                pdt = (time.perf_counter() - w0) / 2
                pcols = sum(nb for nb, _, _ in pblocks)
                extras["e2e_pageable"] = {"value": pcols * NLAY * (NG_LW + NG_SW) / pdt, "unit": UNIT, "columns_per_step": pcols,
                                          "note": "host entry points on pageable (malloc'd) arrays; arrays >= 8 MiB are page-locked in place for the call"}
                del pblocks, pout
            except Exception as exc:
                extras["e2e_pageable_error"] = str(exc)

    # ---- second e2e figure (row N1): the fluxes, not the optical properties, cross PCIe ----------------------------
    # pinned host inputs -> H2D -> gas optics + device flux solvers (ecckd_rte_{lw,sw}_f64) -> D2H of five (ncol, nlay+1)
    # flux arrays per block.
    e2e_fluxes = None
    if not args.no_e2e and not args.no_extras:
        try:
            nf = min(ncol, 1 << 18)                       # a quarter of the 1M-column step: it is a rate
            FB = min(B, nf)
            nfb = (nf + FB - 1) // FB
            keys = ("plev", "tlay", "tlev", "h2o", "o3")
            hin = []
            for b in range(nfb):
                nb = min(FB, nf - b * FB)
                atm = synthetic.make_atmosphere(nb, NLAY, col0=col_base + b * FB, device=dev)
                hc = {k: getattr(atm, k).t().contiguous().cpu().pin_memory() for k in keys}      # (n, ncol) contiguous == Fortran (ncol, n)
                hc["tsfc"] = atm.tsfc.cpu().pin_memory()
                hin.append((nb, hc, atm.scalars))
                del atm
            hflux = [torch.empty((5, NLAY + 1, FB), dtype=torch.float64).pin_memory() for _ in range(2)]
            o_lw = {k: FE((FB, NLAY, NG_LW), torch.float64, dev) for k in ("tau", "lay_source", "lev_source_inc", "lev_source_dec")}
            o_lw["sfc_source"] = FE((FB, NG_LW), torch.float64, dev)
            o_sw = {k: FE((FB, NLAY, NG_SW), torch.float64, dev) for k in ("tau", "ssa", "g")}
            o_sw["toa_src"] = FE((FB, NG_SW), torch.float64, dev)
            h2d_f = sum(8 * nb * (3 * (NLAY + 1) - 1 + 2 * NLAY + 1) for nb, _, _ in hin)
            d2h_f = sum(8 * nb * 5 * (NLAY + 1) for nb, _, _ in hin)

            # three streams, as a caller that cares about throughput would write it: inputs of block i+1 go up and fluxes of
            # block i-1 come down while block i computes (two input buffers, two pinned flux buffers; the optical
            # properties are one set of device arrays, reused block after block by the in-order compute stream)
            s_cmp, s_in, s_out = torch.cuda.current_stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            dbuf = [{k: torch.empty_like(v, device=dev) for k, v in hin[0][1].items()} for _ in range(2)]
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_cmp = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]

            def flux_step():
                for i, (nb, hc, scal) in enumerate(hin):
                    p_ = i & 1
                    if nb == FB:
                        with torch.cuda.stream(s_in):
                            s_in.wait_event(ev_cmp[p_])                  # block i-2 has finished reading this input buffer
                            for k, v in hc.items():
                                dbuf[p_][k].copy_(v, non_blocking=True)
                            ev_in[p_].record(s_in)
                        s_cmp.wait_event(ev_in[p_])
                        dv = {k: v.t() for k, v in dbuf[p_].items() if k != "tsfc"}
                        tsfc = dbuf[p_]["tsfc"]
                    else:
                        dv = {k: v.to(dev, non_blocking=True).t() for k, v in hc.items() if k != "tsfc"}
                        tsfc = hc["tsfc"].to(dev, non_blocking=True)
                    vals = dict(scal, h2o=dv["h2o"], o3=dv["o3"], no2=0.0)
                    gl = [(g_, vals[g_]) for g_ in synthetic.GAS_ORDER]
                    ol = o_lw if nb == FB else None
                    osw = o_sw if nb == FB else None
                    a_ = lw.gas_optics(None, dv["plev"], dv["tlay"], tsfc, gl, tlev=dv["tlev"], out=ol)
                    up, dn = lw.rte_lw(a_)
                    b_ = sw.gas_optics(None, dv["plev"], dv["tlay"], None, gl, out=osw)
                    sup, sdn, sdr = sw.rte_sw(b_)
                    ev_cmp[p_].record(s_cmp)
                    hf = hflux[p_]
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev_cmp[p_])
                        for q_, t_ in enumerate((up, dn, sup, sdn, sdr)):
                            hf[q_, :, :nb].copy_(t_.t(), non_blocking=True)
                            t_.record_stream(s_out)
                        ev_out[p_].record(s_out)
                torch.cuda.synchronize()

            for _ in range(2):
                flux_step()
            sharding.barrier()
            k_f = max(2, min(args.steps, 3))
            w0 = time.perf_counter()
            for _ in range(k_f):
                flux_step()
            dtf = sharding.max_over_ranks((time.perf_counter() - w0) / k_f, dev)
            sharding.barrier()
            e2e_fluxes = {"value": nf * world * NLAY * (NG_LW + NG_SW) / dtf, "unit": UNIT,
                          "h2d_bytes_per_step": h2d_f, "d2h_bytes_per_step": d2h_f, "columns_per_step": nf, "block_cols": FB, "ms_per_step": dtf * 1e3,
                          "api": "pinned host inputs -> H2D -> gas_optics (device arrays) -> ecckd_rte_{lw,sw}_f64 (device flux solvers: no-scattering LW, "
                                 "two-stream SW) -> D2H of flux_up/dn (LW) and flux_up/dn/dn_dir (SW), each (ncol, nlay+1); optical properties never leave HBM; "
                                 "copy-in, compute and copy-out of consecutive blocks overlap on three streams"}
            del hin, hflux, o_lw, o_sw, dbuf
            torch.cuda.empty_cache()
        except Exception as exc:
            extras["e2e_fluxes_error"] = str(exc)

    # ---- CPU baseline beside it (rank 0, N=1 only) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, secs, threads = cpu_port_rate(args.cpu_cols, 0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "fortran": fortran_probe(),
               "sample": f"{args.cpu_cols} of the same synthetic columns x {NLAY} layers, LW+SW, {secs:.1f} s on {threads} threads "
                         "(oracle/ecckd_oracle.c -O3 -march=native; Fortran reference not buildable offline)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.total_cols > 0 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("synthetic global grid, %d columns x 137 layers, LW+SW column-sharded over the GPUs, fp64 (BASELINE.json configs[4])" % args.total_cols)
                                   if args.total_cols > 0 else
                                   "synthetic GCM block, 1M columns x 137 layers, LW+SW on 1 B200, fp64 (BASELINE.json configs[3]); "
                                   "N GPUs = N x 1M columns column-sharded, no collective",
                       "columns_per_gpu": ncol, "nlay": NLAY, "ngpt_lw": NG_LW, "ngpt_sw": NG_SW, "columns_per_launch": B,
                       "ck_definitions": [os.path.basename(LW_FILE), os.path.basename(SW_FILE)],
                       "l2": "every launch streams >9 GB of outputs and 0.36 GB of inputs: far larger than the 126 MB L2, no flush needed",
                       "gases": "h2o, o3 (ncol,nlay) fields; co2, ch4, n2o, o2, cfc11, cfc12 scalars (RFMIP convention)"},
            "clocks": clocks, "e2e": e2e, "e2e_fluxes": e2e_fluxes, "gpu_launches": launches, "roofline": roofline, "roofline_sw": roofline_sw,
            "roofline_f32": roofline_f32, "roofline_sw_f32": roofline_sw_f32,
            "cpu_baseline": cpu, "extras": extras,
            "split": {"lw_gptl_per_s_per_gpu": B * NLAY * NG_LW / (lw_ms * 1e-3), "sw_gptl_per_s_per_gpu": B * NLAY * NG_SW / (sw_ms * 1e-3)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
