"""One LRA update + apply (n x 32, bf16) inside cudaProfilerStart/Stop, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd
n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 26)
r = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda:0")
torch.manual_seed(0)
sc = (0.1 / (n * r)) ** 0.5
U = (sc * torch.randn(n, r, device=dev)).bfloat16(); V = (sc * torch.randn(n, r, device=dev)).bfloat16()
d = torch.ones(n, 1, device=dev, dtype=torch.bfloat16)
L = [torch.zeros([], device=dev) for _ in range(3)]
g = (0.01 * torch.randn(n, 1, device=dev)).bfloat16()
for i in range(2):
    psgd.update_precond_lra_whiten([U, V, d], L, g, lr=0.1, noise={"v": torch.randn_like(g), "update_U": i % 2 == 0})
noise = {"v": torch.randn_like(g), "update_U": True}
torch.cuda.synchronize()
torch.cuda.profiler.start()
psgd.update_precond_lra_whiten([U, V, d], L, g, lr=0.1, noise=noise)
out = psgd.precond_grad_lra([U, V, d], g)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(out.float().norm()), [float(l) for l in L])
