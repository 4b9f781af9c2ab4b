"""Where does the fp32 apply lose accuracy? engine pieces vs fp64 (computed on the GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g = torch.Generator().manual_seed(1)
WL = torch.randn(n, n, generator=g) / n ** 0.5 + 0.5 * torch.eye(n); WR = torch.randn(n, n, generator=g) / n ** 0.5 + 0.5 * torch.eye(n)
G = (0.1 * WL @ torch.randn(n, n, generator=g) @ WR).to(dev)
Q1 = (1.5 * torch.eye(n) + 0.01 * torch.randn(n, n, generator=g)).to(dev)
Q2 = (1.4 * torch.eye(n) + 0.01 * torch.randn(n, n, generator=g)).to(dev)
rel = lambda a, b: float((a.double() - b).norm() / b.norm())
P1_64, P2_64 = Q1.double().T @ Q1.double(), Q2.double().T @ Q2.double()
ref = P1_64 @ G.double() @ P2_64
out = psgd.precond_grad_kron([[Q1, Q2], None], None, G)
print("engine apply (fp32) vs fp64:", rel(out, ref), " torch fp32 chain:", rel(Q1.T @ (Q1 @ G @ Q2.T) @ Q2, ref))
P1 = psgd.gemm(Q1, Q1, trans_a=True, path=1)
print("P1 = Q^T Q (simt):", rel(P1, P1_64), " torch:", rel(Q1.T @ Q1, P1_64))
Y = psgd.gemm(P1, G, path=1)
print("Y = P1 G (simt):", rel(Y, P1.double() @ G.double()), " torch:", rel(P1 @ G, P1.double() @ G.double()))
Z = psgd.gemm(Y, P1, path=1)
print("Z = Y P1 (simt):", rel(Z, Y.double() @ P1.double()))
A = torch.randn(n, n, device=dev); B = torch.randn(n, n, device=dev)
print("randn GEMM (simt):", rel(psgd.gemm(A, B, path=1), A.double() @ B.double()), " torch:", rel(A @ B, A.double() @ B.double()))
