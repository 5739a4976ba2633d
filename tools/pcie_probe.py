"""PCIe ceiling of the host-array (e2e) path: pinned D2H / H2D bandwidth per GPU, alone and with all ranks copying at once.
  python tools/pcie_probe.py                                   # one GPU
  torchrun --nproc-per-node N tools/pcie_probe.py              # N GPUs concurrently (rank 0 prints the table)
Also prints the box's topology (nvidia-smi topo -m, NUMA nodes) -- what bounds bench.py's e2e figure at N > 1."""
import os, subprocess, time
import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 1 << 27  # 1 GiB of float64
g = torch.empty(n, dtype=torch.float64, device=dev)
h = torch.empty(n, dtype=torch.float64).pin_memory()


def bw(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return n * 8 * reps / (time.perf_counter() - t0) / 1e9


res = {"D2H": bw(lambda: h.copy_(g, non_blocking=True)), "H2D": bw(lambda: g.copy_(h, non_blocking=True))}
if world > 1:
    t = torch.tensor([res["D2H"], res["H2D"]], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        d2h = [float(x[0]) for x in allt]; h2d = [float(x[1]) for x in allt]
        print(f"{world} GPUs copying concurrently, 1 GiB pinned buffers, GB/s per GPU:")
        print("  D2H:", " ".join(f"{v:5.1f}" for v in d2h), f"  aggregate {sum(d2h):.0f}")
        print("  H2D:", " ".join(f"{v:5.1f}" for v in h2d), f"  aggregate {sum(h2d):.0f}")
else:
    print(f"1 GPU, 1 GiB pinned: D2H {res['D2H']:.1f} GB/s, H2D {res['H2D']:.1f} GB/s")
    rows, w = 16384, 8192                     # strided 2-D: rows of 64 KB with a larger pitch on the host side
    g2 = torch.empty(rows, w, dtype=torch.float64, device=dev)
    h2 = torch.empty(rows, 2 * w, dtype=torch.float64).pin_memory()
    h2[:, :w].copy_(g2, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        h2[:, :w].copy_(g2, non_blocking=True)
    torch.cuda.synchronize()
    print(f"D2H 2-D (64 KB rows, pitch 128 KB): {rows * w * 8 * 5 / (time.perf_counter() - t0) / 1e9:.1f} GB/s")
if rank == 0:
    run = lambda c: subprocess.run(c, capture_output=True, text=True).stdout
    print(run(["nvidia-smi", "--query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current", "--format=csv"]))
    print(run(["nvidia-smi", "topo", "-m"]))
    print(run(["bash", "-c", "lscpu | grep -i 'numa\\|socket\\|model name\\|^CPU(s)'; ls /sys/devices/system/node/ 2>/dev/null | tr '\\n' ' '"]))
if world > 1:
    dist.destroy_process_group()
