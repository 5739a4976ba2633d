import torch, time
dev = torch.device("cuda:0")
n = 1 << 27  # 1 GiB of float64
g = torch.empty(n, dtype=torch.float64, device=dev)
h = torch.empty(n, dtype=torch.float64).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(g, non_blocking=True)), ("H2D", lambda: g.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name} contiguous 1 GiB pinned: {n*8/dt/1e9:.1f} GB/s")
# strided 2D: rows of 64 KB with a larger pitch on the host side
rows, w = 16384, 8192
g2 = torch.empty(rows, w, dtype=torch.float64, device=dev)
h2 = torch.empty(rows, 2 * w, dtype=torch.float64).pin_memory()
h2[:, :w].copy_(g2, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): h2[:, :w].copy_(g2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"D2H 2-D (64 KB rows, pitch 128 KB): {rows*w*8/dt/1e9:.1f} GB/s")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
