"""Time the tcgen05 GEMM at the shapes of the Kron path, tile width 256 vs 128, against cuBLAS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib
dev = torch.device("cuda:0")
lib = _lib.load_library(); h = _lib.handle_for(dev)

def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

for (M, N, K, ta, tb) in ((4096, 4096, 4096, 0, 0), (4096, 4096, 4096, 1, 0), (4096, 4096, 4096, 0, 1), (4096, 14336, 4096, 0, 0), (14336, 4096, 4096, 0, 0),
                          (4096, 4096, 14336, 0, 1), (1024, 4096, 1024, 0, 0), (1024, 1024, 4096, 0, 1), (128256, 4096, 4096, 0, 0), (4096, 32, 4096, 1, 0)):
    A = torch.randn((K, M) if ta else (M, K), device=dev).bfloat16()
    B = torch.randn((N, K) if tb else (K, N), device=dev).bfloat16()
    res = []
    for name, fl in (("default", 0), ("2cta-256 only", 128), ("2cta-128 forced", 65536), ("1cta", 8)):
        lib.psgd_debug_set_flags(h, fl)
        try:
            ms = t(lambda: psgd.gemm(A, B, trans_a=bool(ta), trans_b=bool(tb), path=2))
            res.append(f"{name}: {ms*1e3:8.1f} us {2*M*N*K/ms/1e9:7.1f} TF/s")
        except Exception as ex:
            res.append(f"{name}: EXC {ex}")
    lib.psgd_debug_set_flags(h, 0)
    Ao = A.T if ta else A
    Bo = B.T if tb else B
    ms = t(lambda: Ao @ Bo)
    print(f"M={M:6d} N={N:6d} K={K:6d} ta={ta} tb={tb} | " + " | ".join(res) + f" | cuBLAS {ms*1e3:8.1f} us {2*M*N*K/ms/1e9:7.1f} TF/s")
