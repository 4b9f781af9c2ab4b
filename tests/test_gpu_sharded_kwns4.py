"""KWNS4(shard_preconditioners=True) -- BASELINE configs[3], "sharded per-parameter ... via DDP wrapper" -- as a test that runs in the
one-GPU tier: two processes share cuda:0 and talk over gloo (which carries CUDA tensors by staging; NCCL refuses two ranks on one device;
the same script runs over NCCL on 2 GPUs through tools/check_sharded_kwns4.py).  Every rank sees the same gradients (as after DDP's
all-reduce).  Checked, for the parameter-by-parameter and the batched (batch_same_shape) forms:
  * the parameters stay bit-identical on all ranks (the owner's result is what everybody holds),
  * each rank keeps preconditioner state for the parameters it owns only, and together they cover every parameter once,
  * the loss goes down like in the replicated mode,
  * state_dict() is collective and returns the COMPLETE state on every rank; a fresh sharded optimizer that loads it continues with
    bit-identical parameters."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(256, 384), (384,), (128, 2048), (64, 64), (128, 2048), (1, 96, 1, 40), (384,), (256, 384)]


def _worker(rank, world, port, q):
    import faulthandler
    import traceback
    faulthandler.dump_traceback_later(240, exit=True)      # a dead-locked collective must not eat the GPU lease
    try:
        _worker_body(rank, world, port, q)
    except Exception:
        q.put((rank, {"error": traceback.format_exc()}))


def _worker_body(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from psgd_torch_b200 import KWNS4
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        g0 = torch.Generator().manual_seed(3)
        targets = [torch.randn(*s, generator=g0).to(dev) for s in SHAPES]
        mixers = [(torch.randn(s[0], s[0], generator=g0) / s[0] ** 0.5 + torch.eye(s[0])).to(dev) if len(s) == 2 else None for s in SHAPES]

        def loss_fn(ps):
            tot = 0.0
            for p, t, m in zip(ps, targets, mixers):
                e = p - t
                tot = tot + ((m @ e) ** 2).sum() if m is not None else tot + (e ** 2).sum()
            return tot

        def make(sharded, batched):
            torch.manual_seed(11)
            ps = [torch.nn.Parameter(torch.zeros(*s, device=dev)) for s in SHAPES]
            opt = KWNS4(ps, lr_params=0.05, lr_preconditioner=0.3, weight_decay=0.0, preconditioner_dtype=torch.float32,
                        shard_preconditioners=sharded, batch_same_shape=batched)
            return ps, opt

        def steps(ps, opt, n):
            losses = []
            for _ in range(n):
                loss = loss_fn(ps)
                losses.append(float(loss.detach()))
                for p, g in zip(ps, torch.autograd.grad(loss, ps)):
                    p.grad = g
                opt.step()
            return losses

        def cross_rank_diff(ps):
            worst = 0.0
            for p in ps:
                ref = p.detach().clone()
                dist.broadcast(ref, src=0)
                worst = max(worst, float((p.detach() - ref).abs().max()))
            return worst

        out = {}
        ps, opt = make(False, False)
        out["rep"] = steps(ps, opt, 40)
        for name, batched in (("sharded", False), ("sharded_batched", True)):
            ps, opt = make(True, batched)
            losses = steps(ps, opt, 40)
            owned = [i for i, p in enumerate(ps) if len(opt.state[p]) > 0]
            # collective checkpoint: complete on every rank; resume into a fresh optimizer and keep going
            sd = opt.state_dict()
            complete = len(sd["state"]) == len(SHAPES)
            ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
            torch.manual_seed(999 + rank)
            opt2 = KWNS4(ps2, lr_params=0.05, lr_preconditioner=0.3, weight_decay=0.0, preconditioner_dtype=torch.float32,
                         shard_preconditioners=True, batch_same_shape=batched)
            opt2.load_state_dict(sd)
            owned2 = [i for i, p in enumerate(ps2) if len(opt2.state[p]) > 0]
            more = steps(ps2, opt2, 5)
            out[name] = dict(losses=losses, diff=cross_rank_diff(ps), owned=owned, complete=complete, owned2=owned2, more=more,
                             diff2=cross_rank_diff(ps2))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_sharded_kwns4_two_ranks_on_one_gpu():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 35500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = dict(q.get(timeout=300) for _ in range(2))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for r in (0, 1):
        assert "error" not in res[r], res[r]["error"]
    rep = res[0]["rep"]
    assert rep[-1] < 0.2 * rep[0]
    for name in ("sharded", "sharded_batched"):
        a, b = res[0][name], res[1][name]
        assert a["diff"] == 0.0 and b["diff"] == 0.0, "parameters must be bit-identical on every rank"
        assert sorted(a["owned"] + b["owned"]) == list(range(len(SHAPES))), (a["owned"], b["owned"])     # a partition of the parameters
        assert a["owned"] and b["owned"]
        assert a["complete"] and b["complete"], "state_dict() must return every parameter's state on every rank"
        assert a["owned2"] == a["owned"] and b["owned2"] == b["owned"], "a resumed rank keeps the state of its own parameters only"
        assert a["diff2"] == 0.0 and b["diff2"] == 0.0
        la = a["losses"]
        assert la[-1] < 0.2 * la[0]
        assert abs(la[-1] - rep[-1]) < 0.3 * max(la[-1], rep[-1]) + 1e-6, (name, la[-1], rep[-1])
        assert a["more"][-1] <= 1.05 * a["more"][0]          # the resumed run keeps descending from where the checkpoint left off
