"""cuBLAS vs 2-CTA (default) vs 1-CTA tcgen05 GEMM at 4096^3 inside cudaProfilerStart/Stop (for ncu --set full)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib
dev = torch.device("cuda:0"); lib = _lib.load_library(); h = _lib.handle_for(dev)
A = torch.randn(4096, 4096, device=dev).bfloat16(); B = torch.randn(4096, 4096, device=dev).bfloat16()
for _ in range(3):
    A @ B; psgd.gemm(A, B, path=2)
lib.psgd_debug_set_flags(h, 8); psgd.gemm(A, B, path=2); lib.psgd_debug_set_flags(h, 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
C0 = A @ B
C1 = psgd.gemm(A, B, path=2)                                     # default: the 2-CTA kernel (gemm_tc2_kernel<256>)
lib.psgd_debug_set_flags(h, 8); C2 = psgd.gemm(A, B, path=2)     # bit 3: 1-CTA kernel (gemm_tc_kernel<256>)
lib.psgd_debug_set_flags(h, 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float((C0.float() - C1.float()).abs().max()), float((C0.float() - C2.float()).abs().max()))
