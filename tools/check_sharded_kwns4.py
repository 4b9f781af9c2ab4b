"""torchrun check of KWNS4(shard_preconditioners=True): owner-computes sharding of the per-parameter preconditioners with a broadcast of the
updated parameters (BASELINE configs[3]).  Every rank sees the same gradients (as after DDP's all-reduce); checks: parameters stay
identical on all ranks, each rank keeps state for its own parameters only, and the loss goes down like in the replicated mode.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_sharded_kwns4.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from psgd_torch_b200 import KWNS4

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shapes = [(256, 384), (384,), (128, 2048), (64, 64), (512, 512), (1, 96, 1, 40)]
g0 = torch.Generator().manual_seed(3)
targets = [torch.randn(*s, generator=g0).to(dev) for s in shapes]
mixers = [(torch.randn(s[0], s[0], generator=g0) / s[0] ** 0.5 + torch.eye(s[0])).to(dev) if len(s) == 2 else None for s in shapes]


def loss_fn(ps):
    tot = 0.0
    for p, t, m in zip(ps, targets, mixers):
        e = p - t
        tot = tot + ((m @ e) ** 2).sum() if m is not None else tot + (e ** 2).sum()
    return tot


def run(sharded, batched=False, exchange="all_gather"):
    torch.manual_seed(11)
    ps = [torch.nn.Parameter(torch.zeros(*s, device=dev)) for s in shapes]
    opt = KWNS4(ps, lr_params=0.05, lr_preconditioner=0.3, weight_decay=0.0, preconditioner_dtype=torch.float32, shard_preconditioners=sharded,
                batch_same_shape=batched, exchange=exchange)
    losses = []
    for step in range(60):
        loss = loss_fn(ps)
        losses.append(float(loss.detach()))
        grads = torch.autograd.grad(loss, ps)
        for p, g in zip(ps, grads):
            p.grad = g
        opt.step()
    # parameters must be identical on every rank
    worst = 0.0
    for p in ps:
        ref = p.detach().clone()
        dist.broadcast(ref, src=0)
        worst = max(worst, float((p.detach() - ref).abs().max()))
    n_state = sum(1 for p in ps if len(opt.state[p]) > 0)
    return losses, worst, n_state


l_rep, w_rep, n_rep = run(False)
l_sh, w_sh, n_sh = run(True)
l_ag, w_ag, n_ag = run(True, batched=True)      # same-shape batches, one all-gather per round of batches
l_pp, w_pp, n_pp = run(True, batched=True, exchange="p2p")   # owners push into the peers' parameters (CUDA IPC + peer copies)
counts = [torch.zeros(1, device=dev) for _ in range(world)]
dist.all_gather(counts, torch.tensor([float(n_sh)], device=dev))
if rank == 0:
    print(f"replicated: loss {l_rep[0]:.4e} -> {l_rep[-1]:.4e}, max cross-rank param diff {w_rep:.2e}, params with state on rank 0: {n_rep}/{len(shapes)}")
    print(f"sharded   : loss {l_sh[0]:.4e} -> {l_sh[-1]:.4e}, max cross-rank param diff {w_sh:.2e}, params with state per rank: {[int(c.item()) for c in counts]} of {len(shapes)}")
    ok = w_sh == 0.0 and l_sh[-1] < 0.2 * l_sh[0] and abs(l_sh[-1] - l_rep[-1]) < 0.3 * max(l_rep[-1], l_sh[-1]) + 1e-6 \
        and sum(int(c.item()) for c in counts) == len(shapes)
    print(f"sharded + batched (all-gather exchange): loss {l_ag[0]:.4e} -> {l_ag[-1]:.4e}, max cross-rank param diff {w_ag:.2e}")
    ok = ok and w_ag == 0.0 and l_ag[-1] < 0.2 * l_ag[0] and abs(l_ag[-1] - l_rep[-1]) < 0.3 * max(l_rep[-1], l_ag[-1]) + 1e-6
    print(f"sharded + batched (p2p push exchange): loss {l_pp[0]:.4e} -> {l_pp[-1]:.4e}, max cross-rank param diff {w_pp:.2e}")
    ok = ok and w_pp == 0.0 and l_pp[-1] < 0.2 * l_pp[0] and abs(l_pp[-1] - l_rep[-1]) < 0.3 * max(l_rep[-1], l_pp[-1]) + 1e-6
    print("sharded KWNS4:", "OK" if ok else "MISMATCH")
dist.barrier()
dist.destroy_process_group()
