"""Bisect accuracy of one Kron shape: engine (various debug flags / paths) vs bf16 oracle vs fp64, step by step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib
from oracle import psgd_oracle as orc
m = int(sys.argv[1]); n = int(sys.argv[2])
dev = torch.device("cuda:0"); lib = _lib.load_library(); h = _lib.handle_for(dev)
torch.set_num_threads(os.cpu_count())
def rel(a, b): a, b = a.detach().double().cpu(), b.detach().double().cpu(); return float((a - b).norm() / b.norm())
def structured(m, n, seed):
    g = torch.Generator().manual_seed(seed)
    WL = torch.randn(m, m, generator=g) / m ** 0.5 + 0.5 * torch.eye(m); WR = torch.randn(n, n, generator=g) / n ** 0.5 + 0.5 * torch.eye(n)
    return (0.1 * WL @ torch.randn(m, n, generator=g) @ WR).bfloat16()
def to_dev(noise): return {"N": noise["N"].to(dev), "balance": noise["balance"], "spd": [None if v is None else v.to(dev) for v in noise["spd"]], "skh": [None if v is None else v.to(dev) for v in noise["skh"]]}
configs = [("default", 0, 0), ("no-sym", 0, 1), ("no-Pfirst", 0, 2), ("no-sym,no-Pfirst", 0, 3), ("no-tc-bound", 0, 4), ("all-off", 0, 7), ("simt", 1, 0)]
for name, path, flags in configs:
    lib.psgd_set_gemm_path(h, path); lib.psgd_debug_set_flags(h, flags)
    QLe, exprs = psgd.init_kron(torch.zeros(m, n, dtype=torch.bfloat16, device=dev))
    out = []
    for step in range(3):
        G = structured(m, n, 100 + step)
        torch.manual_seed(1234 + step)
        noise = orc.draw_kron_noise(G, [q.cpu() for q in QLe[0]]); noise["balance"] = False
        Qo = [q.detach().cpu().clone() for q in QLe[0]]; Lo = [l.detach().cpu().clone() for l in QLe[1]]
        Q64 = [q.detach().cpu().double() for q in QLe[0]]; L64 = [l.detach().cpu().double() for l in QLe[1]]
        n64 = {"N": noise["N"].double(), "balance": False, "spd": [None if v is None else v.double() for v in noise["spd"]], "skh": [None if v is None else v.double() for v in noise["skh"]]}
        orc.update_precond_kron_whiten_q0p5eq1p5([Qo, Lo], G, noise, lr=0.5)
        orc.update_precond_kron_whiten_q0p5eq1p5([Q64, L64], G.double(), n64, lr=0.5)
        psgd.update_precond_kron_whiten_q0p5eq1p5(QLe, exprs, G.to(dev), lr=0.5, noise=to_dev(noise))
        out.append(f"step{step}: Q e/64 {rel(QLe[0][0], Q64[0]):.2e} o/64 {rel(Qo[0], Q64[0]):.2e} | L e {float(QLe[1][0]):.5g} o {float(Lo[0]):.5g} 64 {float(L64[0]):.5g}")
    print(f"{name:18s} " + " || ".join(out))
lib.psgd_set_gemm_path(h, 0); lib.psgd_debug_set_flags(h, 0)
