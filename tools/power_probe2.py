import os, sys, subprocess, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib
dev = torch.device("cuda:0"); lib = _lib.load_library(); h = _lib.handle_for(dev)
A = torch.randn(4096, 4096, device=dev).bfloat16(); B = torch.randn(4096, 4096, device=dev).bfloat16()
C = torch.empty(4096, 4096, device=dev, dtype=torch.bfloat16)
def run(name, fn, secs=2.5):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(200): fn()
        n += 200; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:14s}: {ms*1e3:7.1f} us/GEMM  {2*4096**3/ms/1e9:7.1f} TF/s")
run("cuBLAS", lambda: torch.matmul(A, B, out=C))
lib.psgd_debug_set_flags(h, 16)
lib.psgd_debug_set_mn_desc(h, 8192, 1024); run("relaxed-wait", lambda: psgd.gemm(A, B, path=2))
lib.psgd_debug_set_mn_desc(h, 8193, 1024); run("spin-wait", lambda: psgd.gemm(A, B, path=2))
lib.psgd_debug_set_mn_desc(h, 8192, 1024); run("relaxed-wait", lambda: psgd.gemm(A, B, path=2))
lib.psgd_debug_set_mn_desc(h, 8193, 1024); run("spin-wait", lambda: psgd.gemm(A, B, path=2))
lib.psgd_debug_set_mn_desc(h, 8192, 1024); lib.psgd_debug_set_flags(h, 0)
