"""One Kron update + apply of a single (m x n) tensor inside cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_kron.py 4096 4096
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd

m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
max_skew = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dev = torch.device("cuda:0")
torch.manual_seed(0)
G = (0.01 * torch.randn(m, n, device=dev)).bfloat16()
QL, exprs = psgd.init_kron(G, max_skew=max_skew)
for it in range(3):
    psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.5)
noise = psgd.draw_kron_noise(G, QL[0])
torch.cuda.synchronize()
torch.cuda.profiler.start()
psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1, noise=noise)
H = psgd.precond_grad_kron(QL, exprs, G)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(H.float().norm()), [float(l) for l in QL[1]])
