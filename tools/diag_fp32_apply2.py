"""Replicates test_fp32_full_size_step: engine update then apply at 2048^2 fp32; fp64 on the GPU; localises the error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd
dev = torch.device("cuda:0")
n = 2048
g = torch.Generator().manual_seed(77)
WL = torch.randn(n, n, generator=g) / n ** 0.5 + 0.5 * torch.eye(n); WR = torch.randn(n, n, generator=g) / n ** 0.5 + 0.5 * torch.eye(n)
G = (0.1 * WL @ torch.randn(n, n, generator=g) @ WR).to(dev)
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
QL, exprs = psgd.init_kron(torch.zeros(n, n, device=dev))
for step in range(2):
    psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.5)
    Q1, Q2 = QL[0]
    print("Q1 diag mean", float(Q1.diagonal().mean()), "offdiag rms", float((Q1 - torch.diag(Q1.diagonal())).pow(2).mean().sqrt()), "asym", float((Q1 - Q1.T).abs().max()))
    P1_64, P2_64 = Q1.double().T @ Q1.double(), Q2.double().T @ Q2.double()
    ref = P1_64 @ G.double() @ P2_64
    out = psgd.precond_grad_kron(QL, exprs, G)
    print(f"step {step}: engine apply vs fp64 {rel(out, ref):.3e}; torch fp32 chain {rel(Q1.T @ (Q1 @ G @ Q2.T) @ Q2, ref):.3e}; torch fp32 P-first {rel((Q1.T @ Q1) @ G @ (Q2.T @ Q2), ref):.3e}")
    P1 = psgd.gemm(Q1, Q1, trans_a=True, path=1); P2 = psgd.gemm(Q2, Q2, trans_a=True, path=1)
    Y = psgd.gemm(P1, G, path=1); Z = psgd.gemm(Y, P2, path=1)
    print(f"   by hand with engine GEMMs: P1 {rel(P1, P1_64):.3e}  Y {rel(Y, P1_64 @ G.double()):.3e}  Z {rel(Z, ref):.3e}; out vs Z {rel(out, Z):.3e}")
    out2 = psgd.precond_grad_kron(QL, exprs, G)
    print("   repeat apply identical:", bool(torch.equal(out, out2)))
