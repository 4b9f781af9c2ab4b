"""Bring-up diagnostics on a real B200: GEMM kernels vs torch, descriptor variants, first timings.
Prints everything, never raises (so one gpurun call returns the whole picture)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib

dev = torch.device("cuda:0")
print("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
lib = _lib.load_library()
h = _lib.handle_for(dev)


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def ref_mm(A, B, ta, tb):
    Af, Bf = A.double(), B.double()
    return (Af.T if ta else Af) @ (Bf.T if tb else Bf)


def try_gemm(M, N, K, ta, tb, dtype, path, tag=""):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + ta * 2 + tb)
    A = torch.randn((K, M) if ta else (M, K), generator=g).to(dtype).to(dev)
    B = torch.randn((N, K) if tb else (K, N), generator=g).to(dtype).to(dev)
    try:
        Cm = psgd.gemm(A, B, trans_a=ta, trans_b=tb, path=path)
        torch.cuda.synchronize()
        e = rel(Cm, ref_mm(A, B, ta, tb))
    except Exception as ex:  # noqa
        e = f"EXC {type(ex).__name__}: {ex}"
    print(f"  gemm{tag} path={path} {str(dtype)[6:]:9s} M={M} N={N} K={K} ta={int(ta)} tb={int(tb)} rel={e}")
    return e


print("== SIMT GEMM ==")
for dt in (torch.float32, torch.bfloat16):
    for (M, N, K) in ((5, 7, 3), (64, 64, 16), (100, 130, 77), (32, 300, 300)):
        for ta in (False, True):
            for tb in (False, True):
                try_gemm(M, N, K, ta, tb, dt, 1)

print("== tcgen05 GEMM (default MN descriptor LBO=8192 SBO=1024) ==")
ok = {}
for (ta, tb) in ((False, True), (False, False), (True, True), (True, False)):
    e = try_gemm(256, 512, 192, ta, tb, torch.bfloat16, 2)
    ok[(ta, tb)] = isinstance(e, float) and e < 1e-2
print("tc ok map:", ok)
if not all(ok.values()):
    for (lbo, sbo) in ((1024, 8192), (8192, 128), (128, 8192), (16, 1024), (1024, 16)):
        print(f"-- trying MN descriptor LBO={lbo} SBO={sbo}")
        lib.psgd_debug_set_mn_desc(h, lbo, sbo)
        for (ta, tb) in ((False, False), (True, True), (True, False)):
            try_gemm(256, 512, 192, ta, tb, torch.bfloat16, 2)
    lib.psgd_debug_set_mn_desc(h, 8192, 1024)

print("== tcgen05 edge shapes ==")
for (M, N, K) in ((128, 128, 64), (384, 640, 1000), (1000, 136, 264), (130, 136, 72), (4096, 4096, 4096)):
    for (ta, tb) in ((False, True), (False, False), (True, False)):
        try_gemm(M, N, K, ta, tb, torch.bfloat16, 2)

print("== timing 4096^3 bf16 ==")
for (ta, tb) in ((False, True), (False, False), (True, False), (True, True)):
    try:
        A = torch.randn(4096, 4096, device=dev).bfloat16(); B = torch.randn(4096, 4096, device=dev).bfloat16()
        for _ in range(3): psgd.gemm(A, B, trans_a=ta, trans_b=tb, path=2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): psgd.gemm(A, B, trans_a=ta, trans_b=tb, path=2)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"  tc ta={int(ta)} tb={int(tb)}: {ms:.3f} ms  {2*4096**3/ms/1e9:.1f} TFLOP/s")
    except Exception as ex:
        print("  timing EXC", ex)
try:
    A = torch.randn(4096, 4096, device=dev).bfloat16(); B = torch.randn(4096, 4096, device=dev).bfloat16()
    for _ in range(3): A @ B
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): A @ B
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"  cuBLAS (torch.matmul): {ms:.3f} ms  {2*4096**3/ms/1e9:.1f} TFLOP/s")
except Exception as ex:
    print("  cublas EXC", ex)

print("== Kron update 4096x4096 bf16 timing (whole update + apply) ==")
try:
    m = n = 4096
    G = (0.01 * torch.randn(m, n, device=dev)).bfloat16()
    QL, exprs = psgd.init_kron(G)
    for it in range(3):
        psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.5)
    torch.cuda.synchronize()
    noise = psgd.draw_kron_noise(G, QL[0])
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(5): psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1, noise=noise)
    e1.record()
    for _ in range(5): Hh = psgd.precond_grad_kron(QL, exprs, G)
    e2.record(); torch.cuda.synchronize()
    print(f"  update {e0.elapsed_time(e1)/5:.3f} ms   apply {e1.elapsed_time(e2)/5:.3f} ms   |Q|max={float(QL[0][0].abs().max()):.3f} L={[float(l) for l in QL[1]]}")
except Exception as ex:
    traceback.print_exc()
print("launches:", _lib.launch_count())
