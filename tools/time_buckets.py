"""Per-bucket device time of one update and one apply (CUDA events), plus torch.randn noise cost. For prioritising kernel work."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from psgd_torch_b200 import psgd

dev = torch.device("cuda:0")
only = [a for a in sys.argv[1:] if not a.startswith("--")]  # optional bucket names
BATCHED = "--batched" in sys.argv     # time the groups bench.py runs: same-shape units per engine call, in-kernel Philox noise


def timeit(fn, iters):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

tot = 0.0
for name, count, shape, kind in bench.LLAMA3_8B_SET:
    if only and name not in only: continue
    if BATCHED:
        nb = min(count, bench.BATCH.get(name, 1))
        units = bench.build_units([(name, shape, kind)] * nb, dev)
        groups = bench.make_groups(units)
        psgd.set_noise_mode("philox")
        t = timeit(lambda: bench.step_resident(units, groups, psgd), 5) / nb
        print(f"{name:22s} x{count:3d} shape={shape}: batch of {nb:2d}: {t:8.3f} ms per unit  -> bucket {t*count:8.1f} ms")
        tot += t * count
        del units
        torch.cuda.empty_cache()
        continue
    u = bench.build_units([(name, shape, kind)], dev)[0]
    iters = 3 if u.numel > 1e8 else 10
    if kind == "kron":
        noise = psgd.draw_kron_noise(u.G, u.QL[0])
        t_noise = timeit(lambda: psgd.draw_kron_noise(u.G, u.QL[0]), iters)
        t_upd = timeit(lambda: psgd.update_precond_kron_whiten_q0p5eq1p5(u.QL, u.exprs, u.G, lr=0.1, noise=noise), iters)
        t_app = timeit(lambda: psgd.precond_grad_kron(u.QL, u.exprs, u.G), iters)
    else:
        noise = {"v": torch.randn_like(u.G), "update_U": True}
        t_noise = timeit(lambda: torch.randn_like(u.G), iters)
        t_upd = timeit(lambda: psgd.update_precond_lra_whiten(u.UVd, u.Luvd, u.G, lr=0.1, noise=noise), iters)
        t_app = timeit(lambda: psgd.precond_grad_lra(u.UVd, u.G), iters)
    t_all = timeit(lambda: bench.run_unit(u, u.G, psgd), iters)
    print(f"{name:22s} x{count:3d} shape={shape}: noise {t_noise:8.3f} ms  update {t_upd:8.3f} ms  apply {t_app:8.3f} ms  unit(all) {t_all:8.3f} ms  -> bucket {t_all*count:8.1f} ms")
    tot += t_all * count
    del u
    torch.cuda.empty_cache()
print(f"sum over buckets: {tot:.1f} ms / step")
