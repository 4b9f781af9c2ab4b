"""CPU test (gloo, world_size 4, 2 x 2 device mesh) of KWNS4DTensor on REAL DTensors: the host logic of
wrapped_as_torch_optimizer_for_dtensor.py:98-185 that a single process cannot show --

  * every rank preconditions its LOCAL shard (to_local, dtensor.py:123) -- compared with the oracle's per-parameter step run on the
    same shard with the same draws;
  * a rank whose shard is empty skips the parameter (dtensor.py:124-125) and keeps no state for it;
  * the private RNG states are synchronised at construction (dtensor.py:89-96) so that ranks holding the same shard (the replicas
    along the "dp" mesh dim) draw the same numbers and stay bit-identical;
  * resync_every: parameter, momentum and (Q, L) are broadcast along every mesh dim on which the parameter is replicated, from the
    first rank of that group (dtensor.py:167-179) -- a deliberately perturbed replica is pulled back.

The CUDA engine is replaced by an oracle-backed CPU stand-in with the same call signatures (the checker standing in for the library;
the product classes never import it): what is under test is the wrapper, the engine arithmetic is covered on the GPU."""
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stand_ins(log):
    """(_lib stand-in, psgd stand-in) for kwns4.py on CPU tensors: `ptr` hands the tensor itself through."""
    from oracle import psgd_oracle as orc
    from psgd_torch_b200 import psgd as real_psgd

    class Lib:
        @staticmethod
        def psgd_kwns4_head(h, numel, p, p_dt, grad, g_dt, wd, lr, decoupled, ema, g_out, pre_dt, beta, stream):   # ddp.py:117-143
            g = grad
            if wd > 0.0:
                if decoupled:
                    p.mul_(1.0 - wd * lr)
                else:
                    g = g.add(p.reshape(g.shape), alpha=wd)
            gq = g.reshape(-1).to(ema.dtype if ema is not None else g_out.dtype)
            if g_out is not None:
                g_out.reshape(-1).copy_(gq)
            if ema is not None:
                ema.reshape(-1).mul_(beta).add_(gq, alpha=1.0 - beta)
            return 0

        @staticmethod
        def psgd_kwns4_tail(h, numel, amp_numel, p, p_dt, hbuf, h_dt, sumsq, max_avg, max_elem, lr, stream):   # ddp.py:153-157
            avg = torch.sqrt(sumsq.reshape(()) / amp_numel)
            if avg > max_avg:
                hbuf.mul_(max_avg / avg)
            hbuf.clamp_(min=-max_elem, max=max_elem)
            p.sub_(hbuf.reshape(p.shape).to(p.dtype), alpha=lr)
            return 0

    lib = types.SimpleNamespace(
        load_library=lambda: Lib, ptr=lambda t: t, dtype_code=lambda t: 0, _DTYPES={torch.bfloat16: 0, torch.float32: 1},
        stream_ptr=lambda d: None, handle_for=lambda d: None, check=lambda h, rc, what: None, EngineError=RuntimeError)

    def update(QL, exprs, G, lr=0.1, betaL=0.9, damping=1e-9):
        noise = orc.draw_kron_noise(G, QL[0])
        log.append(("update", tuple(G.shape), float(noise["N"].reshape(-1)[0])))
        orc.update_precond_kron_whiten_q0p5eq1p5(QL, G, noise, lr=lr, betaL=betaL, damping=damping)

    def apply(QL, exprs, G, sumsq_out=None):
        out = orc.precond_grad_kron(QL[0], G)
        if sumsq_out is not None:
            sumsq_out.copy_((out.float() ** 2).sum().reshape(sumsq_out.shape))
        return out

    def update_batched(QLs, exprs, Gs, lr=0.1, betaL=0.9, damping=1e-9):
        for QL, G in zip(QLs, Gs):
            update(QL, None, G, lr=lr, betaL=betaL, damping=damping)

    def apply_batched(QLs, exprs, Gs, sumsq_out=None):
        outs = [apply(QL, None, G) for QL, G in zip(QLs, Gs)]
        if sumsq_out is not None:
            sumsq_out.copy_(torch.stack([(o.float() ** 2).sum() for o in outs]))
        return outs

    lib.MAX_BATCH = 16
    ps = types.SimpleNamespace(init_kron=real_psgd.init_kron, update_precond_kron_whiten_q0p5eq1p5=update, precond_grad_kron=apply,
                               update_precond_kron_whiten_q0p5eq1p5_batched=update_batched, precond_grad_kron_batched=apply_batched,
                               exprs_for_state=real_psgd.exprs_for_state)
    return lib, ps


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.tensor import Replicate, Shard, distribute_tensor
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from psgd_torch_b200 import kwns4
        log = []
        kwns4._lib, kwns4.psgd = _stand_ins(log)
        mesh = init_device_mesh("cpu", (2, 2), mesh_dim_names=("dp", "tp"))
        dp_rank, tp_rank = mesh.get_local_rank("dp"), mesh.get_local_rank("tp")
        torch.manual_seed(500 + rank)                       # ranks start from DIFFERENT generator states
        g0 = torch.Generator().manual_seed(3)               # ... but build the same full tensors
        full = {"w1": torch.randn(8, 6, generator=g0), "w2": torch.randn(3, 5, generator=g0), "w3": torch.randn(1, 4, generator=g0),
                "b": torch.randn(5, generator=g0)}
        place = {"w1": [Replicate(), Shard(0)], "w2": [Replicate(), Shard(0)], "w3": [Replicate(), Shard(0)], "b": [Replicate(), Replicate()]}
        params = {k: torch.nn.Parameter(distribute_tensor(v.clone(), mesh, place[k])) for k, v in full.items()}
        opt = kwns4.KWNS4DTensor(list(params.values()), preconditioner_dtype=torch.float32, lr_params=1e-2, lr_preconditioner=0.3, resync_every=2)
        # the private generator states are identical on every rank after construction (dtensor.py:89-96)
        st = opt.cpu_rng_state.clone()
        gathered = [torch.zeros_like(st) for _ in range(world)]
        dist.all_gather(gathered, st)
        same_rng = all(torch.equal(x, gathered[0]) for x in gathered)

        # oracle replica of this rank's local shards, stepped with the draws the wrapper consumed
        from oracle import psgd_oracle as orc
        local0 = {k: p.to_local().detach().clone() for k, p in params.items()}
        oracle_p = {k: v.clone() for k, v in local0.items()}
        oracle_state = {k: {} for k in params}
        worst, empty_skipped, perturbed_restored = 0.0, True, None
        for step in range(4):
            for k, p in params.items():
                gfull = torch.randn(full[k].shape, generator=g0)
                p.grad = distribute_tensor(gfull, mesh, place[k])
            if step == 3 and dp_rank == 1:                  # break a replica on purpose: the resync at the end of this step (state step 4) repairs it
                with torch.no_grad():
                    params["w1"].to_local().add_(0.123)
            ext = torch.get_rng_state()
            # replay for the oracle: the wrapper draws from the private state, in parameter order, skipping empty shards
            torch.set_rng_state(opt.cpu_rng_state)
            coin = torch.rand([])
            assert coin < 1.0
            for k, p in params.items():
                gl = p.grad.to_local()
                if gl.numel() == 0:
                    continue
                Q = oracle_state[k]["QL"][0] if oracle_state[k] else orc.init_kron(gl.squeeze())[0]
                noise = orc.draw_kron_noise(gl.squeeze(), Q)
                if not (step == 3 and k == "w1"):
                    orc.kwns4_param_step(oracle_p[k], gl.clone(), oracle_state[k], noise, preconditioner_dtype=torch.float32,
                                         lr_params=1e-2, lr_preconditioner=0.3)
            torch.set_rng_state(ext)
            opt.step()
            assert torch.equal(torch.get_rng_state(), ext)  # the caller's generator is untouched (dtensor.py:181-185)
            for k, p in params.items():
                lp = p.to_local()
                if lp.numel() == 0:
                    empty_skipped = empty_skipped and (p not in opt.state or len(opt.state[p]) == 0)
                    continue
                if step == 3 and k == "w1":
                    continue                                # perturbed step: compared across replicas below instead
                d = float((lp - oracle_p[k]).norm() / oracle_p[k].norm())
                worst = max(worst, d)
            if step == 3:                                   # after the resync every dp replica of w1 equals the dp_rank-0 copy again
                lp = params["w1"].to_local().detach().clone()
                both = [torch.zeros_like(lp) for _ in range(2)]
                dist.all_gather(both, lp, group=mesh.get_group("dp"))
                perturbed_restored = bool(torch.equal(both[0], both[1]))
        # replicas along dp hold bit-identical shards of everything
        identical = True
        for k, p in params.items():
            lp = p.to_local().detach().clone()
            if lp.numel() == 0:
                continue
            both = [torch.zeros_like(lp) for _ in range(2)]
            dist.all_gather(both, lp, group=mesh.get_group("dp"))
            identical = identical and bool(torch.equal(both[0], both[1]))
        q.put((rank, same_rng, worst, empty_skipped, perturbed_restored, identical, int(params["w3"].to_local().numel()), len(log)))
    finally:
        dist.destroy_process_group()


def test_kwns4_dtensor_on_a_2x2_mesh():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(4)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert any(r[6] == 0 for r in res), "one tp rank must hold an empty shard of the 1 x 4 parameter"
    for rank, same_rng, worst, empty_skipped, restored, identical, w3_numel, ncalls in res:
        assert same_rng, "private RNG states must be identical on every rank after construction"
        assert worst < 1e-6, (rank, worst)                  # local shards follow the oracle's per-parameter step
        assert empty_skipped, "no state may be created for an empty shard"
        assert restored, "resync_every must pull a diverged dp replica back (broadcast along the replicated mesh dim)"
        assert identical, "replicas along the dp mesh dim must stay bit-identical"
        assert ncalls == 4 * (4 if w3_numel else 3)         # one preconditioner update per non-empty local shard and step
