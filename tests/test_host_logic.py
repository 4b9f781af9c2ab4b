"""CPU tests of the host-side logic: owner-computes partition and the world_size-2 (gloo) RNG discipline of the wrappers."""
import os
import sys

import pytest
import torch

from conftest import ROOT


def _llama_costs():
    sys.path.insert(0, ROOT)
    import bench
    from psgd_torch_b200 import partition
    costs = []
    for name, shape, kind in bench.unit_list():
        if kind == "lra":
            costs.append(partition.lra_unit_cost(shape[0], bench.LRA_RANK))
        elif len(shape) == 1:
            costs.append(partition.kron_unit_cost(shape[0], 1, False, False))
        else:
            dl, dr = bench.dense_flags(shape)
            costs.append(partition.kron_unit_cost(shape[0], shape[1], dl, dr))
    return costs


def test_unit_set_matches_survey():
    import bench
    units = bench.unit_list()
    assert len(units) == 291  # 64+64+64+32+2+65
    assert bench.dense_flags((4096, 4096)) == [True, True]
    assert bench.dense_flags((1024, 4096)) == [True, False]
    assert bench.dense_flags((14336, 4096)) == [False, True]
    assert bench.dense_flags((4096, 14336)) == [True, False]
    assert bench.dense_flags((128256, 4096)) == [False, True]
    assert bench.dense_flags((4096,)) == [False]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_lpt_partition_covers_every_unit_once_and_balances(world):
    from psgd_torch_b200 import partition
    costs = _llama_costs()
    parts = partition.lpt_partition(costs, world)
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(len(costs)))
    imb, _ = partition.imbalance(costs, parts)
    assert imb < (1.02 if world <= 4 else 1.35)   # at 8 ranks the single LRA unit is the critical path
    assert partition.lpt_partition(costs, world) == parts  # deterministic: every rank computes the same assignment
    own = partition.owner_of(costs, world)
    assert all(i in parts[own[i]] for i in range(len(costs)))


def test_kron_unit_flops_match_survey_figures():
    from oracle.psgd_oracle import kron_unit_flops
    up, ap = kron_unit_flops(4096, 4096, True, True)
    assert abs(up - 1.6493e12) / 1.6493e12 < 1e-3 and abs(ap - 5.4976e11) / 5.4976e11 < 1e-3
    up, ap = kron_unit_flops(4096, 14336, True, False)
    assert abs(up - 1.5118e12) / 1.5118e12 < 1e-3 and abs(ap - 6.1848e11) / 6.1848e11 < 1e-3


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from psgd_torch_b200 import KWNS4
    torch.manual_seed(100 + rank)  # ranks start with DIFFERENT generator states
    p = torch.nn.Parameter(torch.zeros(4, 6))
    opt = KWNS4([p])
    assert opt.is_distributed
    st = opt.cpu_rng_state.clone()
    gathered = [torch.zeros_like(st) for _ in range(world)]
    dist.all_gather(gathered, st)
    same_state = all(torch.equal(g, gathered[0]) for g in gathered)
    ext_before = torch.get_rng_state()
    opt.step()  # p.grad is None -> no engine call; still swaps the private RNG state in and out (ddp.py:100-104,172-176)
    ext_after = torch.get_rng_state()
    drew = not torch.equal(opt.cpu_rng_state, st)   # the group coin flip (ddp.py:110) advanced the PRIVATE state
    # owner-computes partition is identical on every rank
    from psgd_torch_b200 import partition
    parts = partition.lpt_partition([5.0, 1.0, 3.0, 3.0, 2.0, 8.0], world)
    t = torch.tensor([sum((r + 1) * (i + 1) * 7919 for r, pp in enumerate(parts) for i in pp)], dtype=torch.int64)
    lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    q.put((rank, same_state, torch.equal(ext_before, ext_after), drew, all(int(x) == int(lst[0]) for x in lst)))
    dist.destroy_process_group()


def test_world_size_2_gloo_rng_sync_and_partition():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, same_state, ext_restored, drew, same_parts in res:
        assert same_state, "private RNG states must be identical across ranks after construction (ddp.py:88-96)"
        assert ext_restored, "step() must restore the caller's RNG state (ddp.py:172-176)"
        assert drew and same_parts


def test_row_shard_covers_all_rows_in_aligned_blocks():
    from psgd_torch_b200 import partition
    for n in (525336576, 1000, 70000, 255):
        for world in (1, 2, 4, 8):
            spans = [partition.row_shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (lo, hi), (lo2, _) in zip(spans, spans[1:]):
                assert hi == lo2 and lo <= hi
            for lo, hi in spans[:-1]:
                assert (hi - lo) % 256 == 0 or hi == n    # whole bulk-copy blocks, except where the rows run out


def test_bench_multi_gpu_partition_is_balanced():
    """bench.py at N > 1: Kron units by LPT on the measured unit times with the LRA row shard as every rank's initial load."""
    import bench
    from psgd_torch_b200 import partition
    units = bench.unit_list()
    kron = [u for u in units if u[2] != "lra"]
    lra_ms = bench.UNIT_MS["embed_tokens_lra32"]
    for world in (2, 4, 8):
        parts = partition.lpt_partition([bench.UNIT_MS[u[0]] for u in kron], world, initial_loads=[lra_ms / world] * world)
        assert sorted(i for p in parts for i in p) == list(range(len(kron)))
        loads = [lra_ms / world + sum(bench.UNIT_MS[kron[i][0]] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.03, loads


def test_kwns4_owner_assignment_is_a_partition(monkeypatch):
    """shard_preconditioners=True: the owner map is computed from shapes only (identical on every rank), covers every parameter once."""
    import torch
    from psgd_torch_b200 import kwns4
    ps = [torch.nn.Parameter(torch.zeros(*s)) for s in [(64, 96), (96,), (32, 512), (1, 8, 1, 12), (48, 48)]]
    opt = kwns4.KWNS4(ps, shard_preconditioners=True)
    opt.cpu_rng_state = torch.get_rng_state()
    monkeypatch.setattr(torch.distributed, "get_world_size", lambda *a, **k: 2)
    opt._assign_owners()
    owners = [opt._owner[id(p)] for p in ps]
    assert set(owners) == {0, 1} and len(owners) == len(ps)
    opt2 = kwns4.KWNS4(ps, shard_preconditioners=True)
    opt2.cpu_rng_state = opt.cpu_rng_state
    opt2._assign_owners()
    assert [opt2._owner[id(p)] for p in ps] == owners
    assert float(torch.rand([], generator=opt._coin_gen)) == float(torch.rand([], generator=opt2._coin_gen))


def test_kwns4_checkpoint_carries_private_rng_states():
    """state_dict() of a distributed KWNS4 carries the private generator states the reference forgets (ddp.py:92,96); a reference-style
    checkpoint without them still loads."""
    import io
    import torch
    from psgd_torch_b200 import kwns4
    p = torch.nn.Parameter(torch.zeros(6, 5))
    opt = kwns4.KWNS4([p])
    assert "psgd_rng" not in opt.state_dict()          # single process: no private states, exactly the reference's keys
    opt.is_distributed, opt.cpu_rng_state, opt.cuda_rng_state = True, torch.get_rng_state(), None
    sd = opt.state_dict()
    buf = io.BytesIO(); torch.save(sd, buf); buf.seek(0)
    torch.manual_seed(987)
    opt2 = kwns4.KWNS4([torch.nn.Parameter(torch.zeros(6, 5))])
    opt2.is_distributed, opt2.cpu_rng_state, opt2.cuda_rng_state = True, torch.get_rng_state(), None
    opt2.load_state_dict(torch.load(buf, weights_only=False))
    assert torch.equal(opt2.cpu_rng_state, opt.cpu_rng_state)
    ref_style = {k: v for k, v in sd.items() if k != "psgd_rng"}
    opt2.load_state_dict(ref_style)


# ---- the collective choreography of the row-sharded LRA preconditioner (lra_sharded.ShardedLRA) under gloo, world_size 2 ----
# The engine stages are replaced by a CPU stand-in with the same data flow (stage 1: per-row sums that must be SUM-reduced; stage 2: a row
# update that needs the reduced sums and leaves two per-shard maxima that must be MAX-reduced; stage 4: a row update that needs the
# reduced maxima; apply: two projections SUM-reduced between three row passes, and a global sum of squares).  What is under test is the
# host logic the GPU box cannot show on one device: which buffers are reduced, with which operation, in which order, and the agreement
# on the CPU coin.  (The arithmetic of the real stages is checked on the GPU: tests/test_gpu_sharded_lra.py, tools/check_sharded_lra.py.)
def _make_cpu_stage_class():
    from psgd_torch_b200.lra_sharded import ShardedLRA, ST_SWEEP1, ST_SWEEP2, ST_FINISH

    class CpuStages(ShardedLRA):
        def __init__(self, UVd, group=None):
            self.UVd, self.Luvd, self.group = UVd, None, group
            self.dev = torch.device("cpu")
            r = UVd[0].shape[1]
            self._views = [torch.zeros(2 * r + 1), torch.zeros(2), torch.zeros(r), torch.zeros(r)]
            self._sumsq = torch.zeros(1)
            self.coin_seen = None

        def update_stage(self, stages, gh, v, lr, betaL, damping, whiten, update_U):
            U, V, d = self.UVd
            if stages == ST_SWEEP1:
                self.coin_seen = update_U
                self.sums.copy_(torch.cat([(U * gh).sum(0), (V * v).sum(0), (gh * gh).sum().reshape(1)]))
            elif stages == ST_SWEEP2:
                r = U.shape[1]
                su, sv, sq = self.sums[:r], self.sums[r:2 * r], self.sums[2 * r]
                (U if update_U else V).add_(lr * gh * (su if update_U else sv) / (1.0 + sq))
                self._dd = (U * su).sum(1, keepdim=True) - (V * sv).sum(1, keepdim=True)
                self.maxima.copy_(torch.stack([self._dd.abs().max(), (v * d).abs().max()]))
            elif stages == ST_FINISH:
                d.sub_(lr / (self.maxima[0] + self.maxima[1]) * self._dd * d)

        def apply_stage(self, modes, g, out):
            U, V, d = self.UVd
            if modes == 1:
                self.proj1.copy_((V * (d * g)).sum(0)); self.proj2.zero_(); self._sumsq.zero_()
            elif modes == 2:
                self._y = d * g + U @ self.proj1.reshape(-1, 1)
                self.proj2.copy_((U * self._y).sum(0))
            else:
                out.copy_(d * (self._y + V @ self.proj2.reshape(-1, 1)))
                self._sumsq.copy_((out * out).sum().reshape(1))

    return CpuStages


def _lra_inputs(n, r):
    g = torch.Generator().manual_seed(17)
    U, V = 0.1 * torch.randn(n, r, generator=g), 0.1 * torch.randn(n, r, generator=g)
    d = 1.0 + 0.1 * torch.rand(n, 1, generator=g)
    steps = [(torch.randn(n, 1, generator=g), torch.randn(n, 1, generator=g)) for _ in range(3)]
    return U, V, d, steps


def _lra_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from psgd_torch_b200 import partition
    Cls = _make_cpu_stage_class()
    n, r = 1000, 4
    U, V, d, steps = _lra_inputs(n, r)
    lo, hi = partition.row_shard(n, world, rank, align=8)
    sh = Cls([U[lo:hi].clone(), V[lo:hi].clone(), d[lo:hi].clone()])
    torch.manual_seed(1000 + rank)          # different CPU generators: the coin must still agree (rank 0's draw)
    outs, coins = [], []
    for g, v in steps:
        coin = sh._agree(bool(torch.rand([]) < 0.5))
        coins.append(coin)
        sh._update(g[lo:hi], v[lo:hi], 0.1, 0.9, 0.0, True, coin)
        ssq = torch.zeros(1)
        outs.append((sh.precond_grad_lra(g[lo:hi], sumsq_out=ssq), float(ssq)))
    # numpy arrays travel through the queue by value (tensors would travel as shared-memory handles served by this process, which may
    # have exited by the time the parent unpickles them)
    q.put((rank, lo, hi, [x.numpy().copy() for x in sh.UVd], [(o.numpy().copy(), s_) for o, s_ in outs], coins))
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_lra_choreography():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_lra_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][5] == res[1][5], "every rank must use rank 0's coin (psgd.py:1035)"
    coins = res[0][5]
    # single-process run of the same stand-in on all rows with the same coins
    Cls = _make_cpu_stage_class()
    n, r = 1000, 4
    U, V, d, steps = _lra_inputs(n, r)
    whole = Cls([U.clone(), V.clone(), d.clone()])
    wouts = []
    for (g, v), coin in zip(steps, coins):
        whole._update(g, v, 0.1, 0.9, 0.0, True, coin)
        ssq = torch.zeros(1)
        wouts.append((whole.precond_grad_lra(g, sumsq_out=ssq), float(ssq)))
    for k in range(3):
        got = torch.cat([torch.from_numpy(res[0][3][k]), torch.from_numpy(res[1][3][k])])
        assert torch.allclose(got, whole.UVd[k], rtol=1e-5, atol=1e-6), k
    for i, (wo, ws_) in enumerate(wouts):
        got = torch.cat([torch.from_numpy(res[0][4][i][0]), torch.from_numpy(res[1][4][i][0])])
        assert torch.allclose(got, wo, rtol=1e-4, atol=1e-4)     # fp32 partial sums in a different order
        assert abs(res[0][4][i][1] - ws_) < 1e-4 * abs(ws_) and abs(res[1][4][i][1] - ws_) < 1e-4 * abs(ws_)   # global sum of squares on every rank


# ---- checkpoints ----
def test_kwns4_loads_a_checkpoint_written_by_the_reference_wrapper():
    """tests/golden/refckpt_*.pt hold state_dict()s written by the unmodified reference KWNS4 (ddp.py:131-137 keys) mid-run.  Loading must
    keep every tensor bit-exact in the dtype the reference held it in -- in particular the fp32 Lipschitz constants and an fp32
    preconditioner of a bf16 parameter, which torch.optim.Optimizer.load_state_dict would truncate to the parameter's dtype."""
    import torch
    from conftest import load_golden
    from psgd_torch_b200 import kwns4
    for name in ("refckpt_f32.pt", "refckpt_bf16_param.pt"):
        case = load_golden(name)
        ptype = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16}[case["ptype"]]
        pdtype = {"torch.float32": torch.float32, "torch.bfloat16": torch.bfloat16}[case["pdtype"]]
        p = torch.nn.Parameter(case["p_at_save"].clone().to(ptype))
        opt = kwns4.KWNS4([p], preconditioner_dtype=pdtype, **case["kw"])
        opt.load_state_dict(case["checkpoint"])
        st, ref = opt.state[p], case["checkpoint"]["state"][0]
        assert st["step"] == ref["step"] == case["save_at"]
        for q, qr in zip(st["QL"][0], ref["QL"][0]):
            assert q.dtype == pdtype and torch.equal(q, qr)
        for l, lr_ in zip(st["QL"][1], ref["QL"][1]):
            assert l.dtype == torch.float32 and torch.equal(l, lr_) and float(l) > 0
        assert st["ema"].dtype == pdtype and torch.equal(st["ema"], ref["ema"])
        assert callable(st["exprs"][0]) and len(st["exprs"][1]) == len(st["QL"][0])     # rebuilt from the factor shapes


def test_kwns4_checkpoint_holds_tensors_and_scalars_only():
    """state_dict() replaces the `exprs` callables by a tag, so torch.load(weights_only=True) accepts the file; a round trip restores
    Q / L / ema exactly and the private RNG states from param_groups[0]."""
    import io
    import torch
    from psgd_torch_b200 import kwns4, psgd
    p = torch.nn.Parameter(torch.zeros(6, 5, dtype=torch.bfloat16))
    opt = kwns4.KWNS4([p])
    QL, exprs = psgd.init_kron(torch.zeros(6, 5, dtype=torch.bfloat16))
    QL[0][0] += 0.01 * torch.randn_like(QL[0][0])      # 6^2 > 30: diagonal factor; QL[0][1] is the dense 5 x 5 one
    QL[0][1] += 0.01 * torch.randn(5, 5).bfloat16()
    QL[1][0].fill_(1.2345678)                        # not representable in bf16
    opt.state[p] = {"QL": QL, "exprs": exprs, "step": 7, "ema": torch.randn(6, 5).bfloat16()}
    opt.is_distributed, opt.cpu_rng_state, opt.cuda_rng_state = True, torch.get_rng_state(), None
    buf = io.BytesIO(); torch.save(opt.state_dict(), buf); buf.seek(0)
    sd = torch.load(buf, weights_only=True)
    assert set(sd.keys()) == {"state", "param_groups"}
    p2 = torch.nn.Parameter(torch.zeros(6, 5, dtype=torch.bfloat16))
    opt2 = kwns4.KWNS4([p2])
    opt2.is_distributed, opt2.cpu_rng_state, opt2.cuda_rng_state = True, torch.zeros_like(opt.cpu_rng_state), None
    opt2.load_state_dict(sd)
    st = opt2.state[p2]
    assert st["step"] == 7 and torch.equal(st["QL"][0][0], QL[0][0]) and st["QL"][0][0].dtype == torch.bfloat16
    assert st["QL"][1][0].dtype == torch.float32 and float(st["QL"][1][0]) == float(QL[1][0])
    assert torch.equal(opt2.cpu_rng_state, opt.cpu_rng_state) and "psgd_rng" not in opt2.param_groups[0]
    assert type(st["exprs"][0]).__name__ == "_ExprP"


def test_lra_optimizer_checkpoint_round_trip():
    """LRAWhitenOptimizer keeps its global preconditioner (U, V, d), Lipschitz constants, momentum buffer and counter under the reserved
    state key "psgd_lra": a resumed run continues from the fitted preconditioner instead of a fresh random one."""
    import io
    import torch
    from psgd_torch_b200 import LRAWhitenOptimizer
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(5))]
    opt = LRAWhitenOptimizer(ps, rank_of_approximation=4, preconditioner_init_scale=0.5, momentum=0.9)
    opt._m, opt._counter_m = torch.randn(26, 1), 11
    opt._Luvd[0].fill_(3.25)
    buf = io.BytesIO(); torch.save(opt.state_dict(), buf); buf.seek(0)
    torch.manual_seed(2)
    ps2 = [torch.nn.Parameter(torch.randn(7, 3)), torch.nn.Parameter(torch.randn(5))]
    opt2 = LRAWhitenOptimizer(ps2, rank_of_approximation=4, preconditioner_init_scale=0.5, momentum=0.9)
    assert not torch.equal(opt2._UVd[0], opt._UVd[0])
    opt2.load_state_dict(torch.load(buf, weights_only=True))
    for a, b in zip(opt2._UVd + opt2._Luvd + [opt2._m], opt._UVd + opt._Luvd + [opt._m]):
        assert torch.equal(a, b)
    assert opt2._counter_m == 11


# ---- KWNS4(shard_preconditioners=True): schedule, ownership, collective checkpoint -- world_size 2 over gloo, oracle-backed engine stand-in ----
def _sharded_kwns4_worker(rank, world, port, q):
    import traceback
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from psgd_torch_b200 import kwns4
        from test_dtensor_host_logic import _stand_ins
        kwns4._lib, kwns4.psgd = _stand_ins([])
        shapes = [(12, 16), (16,), (6, 40), (8, 8), (6, 40), (1, 5, 1, 4), (16,), (12, 16)]
        g0 = torch.Generator().manual_seed(3)
        targets = [torch.randn(*s, generator=g0) for s in shapes]
        out = {}
        for name, batched, exchange in (("plain", False, "broadcast"), ("batched", True, "all_gather"), ("batched_bcast", True, "broadcast")):
            torch.manual_seed(11)
            ps = [torch.nn.Parameter(torch.zeros(*s)) for s in shapes]
            opt = kwns4.KWNS4(ps, lr_params=0.05, lr_preconditioner=0.3, weight_decay=0.0, preconditioner_dtype=torch.float32,
                              shard_preconditioners=True, batch_same_shape=batched, exchange=exchange)

            def steps(ps, opt, n):
                losses = []
                for _ in range(n):
                    loss = sum(((p - t) ** 2).sum() for p, t in zip(ps, targets))
                    losses.append(float(loss.detach()))
                    for p, g in zip(ps, torch.autograd.grad(loss, ps)):
                        p.grad = g
                    opt.step()
                return losses

            losses = steps(ps, opt, 12)
            owned = [i for i, p in enumerate(ps) if len(opt.state[p]) > 0]
            sd = opt.state_dict()                       # collective
            ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
            torch.manual_seed(999 + rank)
            opt2 = kwns4.KWNS4(ps2, lr_params=0.05, lr_preconditioner=0.3, weight_decay=0.0, preconditioner_dtype=torch.float32,
                               shard_preconditioners=True, batch_same_shape=batched, exchange=exchange)
            opt2.load_state_dict(sd)
            owned2 = [i for i, p in enumerate(ps2) if len(opt2.state[p]) > 0]
            more = steps(ps2, opt2, 3)
            diffs = []
            for p in list(ps) + list(ps2):
                ref = p.detach().clone()
                dist.broadcast(ref, src=0)
                diffs.append(float((p.detach() - ref).abs().max()))
            out[name] = dict(losses=losses, owned=owned, owned2=owned2, n_saved=len(sd["state"]), more=more, diff=max(diffs),
                             coin=float(torch.rand([], generator=opt2._coin_gen)))
        q.put((rank, out))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, {"error": traceback.format_exc()}))


def test_world_size_2_gloo_sharded_kwns4_schedule_and_checkpoint():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 37500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sharded_kwns4_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = dict(q.get(timeout=180) for _ in range(2))
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.kill()
    for r in (0, 1):
        assert "error" not in res[r], res[r]["error"]
    # (the all-gather exchange visits each rank's batches largest first, so its random draws come in another order than with broadcasts:
    # the trajectories differ in the noise only)
    assert abs(res[0]["batched"]["losses"][-1] - res[0]["batched_bcast"]["losses"][-1]) < 0.1 * res[0]["batched"]["losses"][0]
    for name in ("plain", "batched", "batched_bcast"):
        a, b = res[0][name], res[1][name]
        assert a["diff"] == 0.0 and b["diff"] == 0.0, "parameters must be bit-identical on every rank"
        assert sorted(a["owned"] + b["owned"]) == list(range(8)) and a["owned"] and b["owned"]
        assert a["n_saved"] == 8 and b["n_saved"] == 8, "the collective state_dict() holds every parameter's state on every rank"
        assert a["owned2"] == a["owned"] and b["owned2"] == b["owned"]
        assert a["losses"][-1] < 0.5 * a["losses"][0] and a["more"][-1] < a["more"][0]
        assert a["coin"] == b["coin"], "the resumed update-coin generators must agree across ranks"
