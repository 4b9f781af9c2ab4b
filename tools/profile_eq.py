"""One EQ-geometry (triangular Q, dQ = E*Q) update of a 4096 x 4096 bf16 weight inside a cudaProfiler range, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`; also prints wall-clock vs device time of the update (host-bound or not)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd

dev = torch.device("cuda:0")
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
g = torch.Generator().manual_seed(0)
G = (0.05 * torch.randn(m, n, generator=g)).to(torch.bfloat16).to(dev)
QL, exprs = psgd.init_kron(torch.zeros(m, n, dtype=torch.bfloat16, device=dev), Scale=1.0, dQ="EQ")
for _ in range(4):
    psgd.update_precond_kron_whiten_eq(QL, exprs, G, lr=0.1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(10):
    psgd.update_precond_kron_whiten_eq(QL, exprs, G, lr=0.1)
e1.record()
t_issue = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
print(f"EQ update {m}x{n}: host issue time {t_issue:.3f} ms / update, device time {e0.elapsed_time(e1) / 10:.3f} ms / update")
torch.cuda.cudart().cudaProfilerStart()
psgd.update_precond_kron_whiten_eq(QL, exprs, G, lr=0.1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
