"""Row-sharded LRA (psgd_torch_b200/lra_sharded.py) on ONE GPU: two row shards driven stage by stage, the all-reduces emulated by summing /
maximising the shards' cross-row buffers, against the unsharded engine on the whole preconditioner.  (The NCCL path itself runs in
bench.py --gpus N and tools/check_sharded_lra.py under torchrun.)"""
import pytest
import torch

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,r,dtype,split", [(70000, 32, torch.bfloat16, 256 * 100), (5000, 16, torch.bfloat16, 2304), (3001, 8, torch.float32, 1000)])
def test_two_row_shards_equal_the_whole(n, r, dtype, split):
    from psgd_torch_b200 import psgd
    from psgd_torch_b200.lra_sharded import ShardedLRA, ST_SWEEP1, ST_SWEEP2, ST_FINISH
    dev = torch.device("cuda:0")
    g0 = torch.Generator().manual_seed(n + r)
    sc = (0.1 / (n * r)) ** 0.5
    U = (sc * torch.randn(n, r, generator=g0)).to(dtype).to(dev)
    V = (sc * torch.randn(n, r, generator=g0)).to(dtype).to(dev)
    d = (1.0 + 0.1 * torch.rand(n, 1, generator=g0)).to(dtype).to(dev)
    whole = [U.clone(), V.clone(), d.clone()]
    Lw = [torch.zeros([], device=dev) for _ in range(3)]
    rows = [(0, split), (split, n)]
    shards = []
    for lo, hi in rows:
        UVd = [U[lo:hi].clone(), V[lo:hi].clone(), d[lo:hi].clone()]
        shards.append(ShardedLRA(UVd, [torch.zeros([], device=dev) for _ in range(3)]))
    tol = 2e-6 if dtype == torch.float32 else 4e-3
    for step in range(3):
        g = (0.01 * torch.randn(n, 1, generator=g0) * (1 + torch.arange(n).reshape(n, 1) % 5)).to(dtype).to(dev)
        v = torch.randn(n, 1, generator=g0).to(dtype).to(dev)
        upd_U = step % 2 == 0
        psgd.update_precond_lra_whiten(whole, Lw, g, lr=0.1, noise={"v": v, "update_U": upd_U})
        gs = [g[lo:hi].contiguous() for lo, hi in rows]
        vs = [v[lo:hi].contiguous() for lo, hi in rows]
        for sh, gi, vi in zip(shards, gs, vs):
            sh.update_stage(ST_SWEEP1, gi, vi, 0.1, 0.9, 1e-9, True, upd_U)
        tot = shards[0].sums + shards[1].sums                     # all-reduce SUM
        for sh in shards:
            sh.sums.copy_(tot)
        for sh, gi, vi in zip(shards, gs, vs):
            sh.update_stage(ST_SWEEP2, gi, vi, 0.1, 0.9, 1e-9, True, upd_U)
        mx = torch.maximum(shards[0].maxima, shards[1].maxima)    # all-reduce MAX
        for sh in shards:
            sh.maxima.copy_(mx)
        for sh, gi, vi in zip(shards, gs, vs):
            sh.update_stage(ST_FINISH, gi, vi, 0.1, 0.9, 1e-9, True, upd_U)
        for k in range(3):
            got = torch.cat([sh.UVd[k] for sh in shards])
            assert relerr(got, whole[k]) < tol, (step, k, relerr(got, whole[k]))
        for sh in shards:
            for a, b in zip(sh.Luvd, Lw):
                assert relerr(a, b) < 1e-4
        # apply
        x = (0.01 * torch.randn(n, 1, generator=g0)).to(dtype).to(dev)
        ssq_w = torch.zeros(1, device=dev)
        want = psgd.precond_grad_lra(whole, x, sumsq_out=ssq_w)
        xs = [x[lo:hi].contiguous() for lo, hi in rows]
        outs = [torch.empty_like(xi) for xi in xs]
        for mode, view in ((1, "proj1"), (2, "proj2"), (4, None)):
            for sh, xi, oi in zip(shards, xs, outs):
                sh.apply_stage(mode, xi, oi)
            if view:
                tot = getattr(shards[0], view) + getattr(shards[1], view)
                for sh in shards:
                    getattr(sh, view).copy_(tot)
        assert relerr(torch.cat(outs), want) < tol
        assert relerr(shards[0]._sumsq + shards[1]._sumsq, ssq_w) < 1e-3
