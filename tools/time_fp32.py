"""fp32 preconditioner on a single 4096 x 4096 weight (SURVEY.md 8d config 2, fp32 leg): update + apply time with the fp32 products on the
tensor cores (bf16 triples concatenated along K) vs the CUDA-core kernel (the default), and one big GEMM both ways."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib

dev = torch.device("cuda:0")
lib = _lib.load_library(); h = _lib.handle_for(dev)
m = n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


def timeit(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


A = torch.randn(m, n, device=dev); B = torch.randn(n, n, device=dev)
ref = A.double() @ B.double()
for name, on in (("tensor cores (bf16 x 3)", 1), ("CUDA cores (default)", 0)):
    lib.psgd_set_fp32_tensor_cores(h, on)
    t = timeit(lambda: psgd.gemm(A, B))
    C = psgd.gemm(A, B)
    err = float((C.double() - ref).norm() / ref.norm())
    G = 0.01 * torch.randn(m, n, device=dev)
    QL, exprs = psgd.init_kron(G)
    for _ in range(2): psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1)
    tu = timeit(lambda: psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1), 2)
    ta = timeit(lambda: psgd.precond_grad_kron(QL, exprs, G), 2)
    print(f"fp32 {m}x{n}, {name:24s}: GEMM {t:8.3f} ms ({2*m*n*n/t/1e9:7.1f} TFLOP/s, rel err vs fp64 {err:.2e})   update {tu:8.2f} ms   apply {ta:8.2f} ms")
lib.psgd_set_fp32_tensor_cores(h, 0)
