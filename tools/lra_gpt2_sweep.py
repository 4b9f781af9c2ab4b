"""BASELINE.json configs[4] at full size: one global LRA preconditioner over the GPT-2-small parameter vector (n = 124 439 808, misc/gpt2.py:216-247
with the unpadded 50257 vocabulary), r in {4, 16, 64}, fp32 and bf16: device time of update / apply, HBM traffic rate against the compulsory
bytes (update 6 n r e, apply 3 n r e) and finiteness after 10 steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd

dev = torch.device("cuda:0")
n = 124439808


def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for dtype in (torch.bfloat16, torch.float32):
    for r in (4, 16, 64):
        es = 2 if dtype == torch.bfloat16 else 4
        gen = torch.Generator(device=dev).manual_seed(r)
        sc = (0.1 / (n * r)) ** 0.5
        U = (sc * torch.randn(n, r, device=dev, generator=gen)).to(dtype)
        V = (sc * torch.randn(n, r, device=dev, generator=gen)).to(dtype)
        UVd = [U, V, torch.ones(n, 1, device=dev, dtype=dtype)]
        L = [torch.zeros([], device=dev) for _ in range(3)]
        g = (0.01 * torch.randn(n, 1, device=dev, generator=gen)).to(dtype)
        for _ in range(10):
            psgd.update_precond_lra_whiten(UVd, L, g, lr=0.1)
        noise = {"v": torch.randn_like(g), "update_U": True}
        tu = timeit(lambda: psgd.update_precond_lra_whiten(UVd, L, g, lr=0.1, noise=noise))
        ta = timeit(lambda: psgd.precond_grad_lra(UVd, g))
        ok = all(bool(torch.isfinite(x.float()).all()) for x in UVd)
        print(f"GPT-2-small LRA n={n} r={r:2d} {str(dtype):15s}: update {tu:7.2f} ms ({6 * n * r * es / tu / 1e9:5.2f} TB/s of compulsory traffic)  "
              f"apply {ta:7.2f} ms ({3 * n * r * es / ta / 1e9:5.2f} TB/s)  finite={ok}  L={[round(float(l), 3) for l in L]}")
        del U, V, UVd, g
        torch.cuda.empty_cache()
