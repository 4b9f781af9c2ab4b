"""Device time (CUDA events) of one preconditioner update per geometry on a single m x n bf16 weight (BASELINE.json configs[1]:
4096 x 4096, the "triangular" EQ geometry included), whitening and Newton-pair form, plus the apply.  Usage:
    python tools/time_geometries.py [m n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib

dev = torch.device("cuda:0")
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
dtype = torch.bfloat16
FN = {"EQ": "eq", "QEP": "qep", "QEQ": "qeq", "Q0.5EQ1.5": "q0p5eq1p5", "PRO4P": "pro4p", "QUAD": "quad", "QUAD4P": "quad4p"}


def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count(dev)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (_lib.launch_count(dev) - l0) // iters


g = torch.Generator().manual_seed(0)
G = (0.05 * torch.randn(m, n, generator=g)).to(dtype).to(dev)
V = torch.randn(m, n, generator=g).to(dtype).to(dev)
print(f"single {m} x {n} {dtype} weight, update time per geometry (ms, device) and engine kernels per update")
for dq in ("Q0.5EQ1.5", "EQ", "QEP", "QEQ", "QUAD", "QUAD4P", "PRO4P"):
    QL, exprs = psgd.init_kron(torch.zeros(m, n, dtype=dtype, device=dev), Scale=1.0, dQ=dq)
    for _ in range(5):   # leave the identity: Q becomes a generic (upper-triangular for EQ) matrix
        if dq == "Q0.5EQ1.5":
            psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1)
        else:
            getattr(psgd, f"update_precond_kron_whiten_{FN[dq]}")(QL, exprs, G, lr=0.1)
    if dq == "Q0.5EQ1.5":
        tw, lw = timeit(lambda: psgd.update_precond_kron_whiten_q0p5eq1p5(QL, exprs, G, lr=0.1))
    else:
        f = getattr(psgd, f"update_precond_kron_whiten_{FN[dq]}")
        tw, lw = timeit(lambda: f(QL, exprs, G, lr=0.1))
    fn_ = getattr(psgd, f"update_precond_kron_newton_{FN[dq]}")
    tn, ln = timeit(lambda: fn_(QL, exprs, V, G, lr=0.02))
    if dq in ("PRO4P", "QUAD4P"):
        ta, la = timeit(lambda: exprs[0](*QL[0], G))
    else:
        ta, la = timeit(lambda: psgd.precond_grad_kron(QL, exprs, G))
    ok = all(bool(torch.isfinite(q.float()).all()) for q in QL[0])
    print(f"  {dq:10s} whiten {tw:7.3f} ms ({lw:3d} kernels)   newton {tn:7.3f} ms ({ln:3d})   apply {ta:6.3f} ms ({la:2d})   finite={ok}")
if m == n:
    QL, exprs = psgd.init_kron(torch.zeros(m, n, dtype=dtype, device=dev), Scale=1.0, dQ="EQ")
    for _ in range(3):
        psgd.update_precond_kron_whiten_eq(QL, exprs, G, lr=0.1)
    t, l = timeit(lambda: psgd.solve_kron_factors(QL[0], V))
    print(f"  conjB = QL^-T V QR^-1 alone (two blocked triangular inverses + four products): {t:.3f} ms ({l} kernels)")
