"""torchrun check of the NCCL path of the row-sharded LRA preconditioner: every rank builds the same full (U, V, d, g) from one seed, keeps
its row shard, runs collective updates + applies; rank 0 also runs the unsharded engine on the whole thing and compares its rows.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_lra.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from psgd_torch_b200 import psgd, partition
from psgd_torch_b200.lra_sharded import ShardedLRA

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, r = 1 << 22, 32
g0 = torch.Generator().manual_seed(5)
sc = (0.1 / (n * r)) ** 0.5
U = (sc * torch.randn(n, r, generator=g0)).bfloat16().to(dev)
V = (sc * torch.randn(n, r, generator=g0)).bfloat16().to(dev)
d = torch.ones(n, 1).bfloat16().to(dev)
lo, hi = partition.row_shard(n, world, rank)
sh = ShardedLRA([U[lo:hi].clone(), V[lo:hi].clone(), d[lo:hi].clone()], [torch.zeros([], device=dev) for _ in range(3)])
whole, Lw = [U, V, d], [torch.zeros([], device=dev) for _ in range(3)]
worst = 0.0
for step in range(3):
    g = (0.01 * torch.randn(n, 1, generator=g0)).bfloat16().to(dev)
    v = torch.randn(n, 1, generator=g0).bfloat16().to(dev)
    noise = {"v": v[lo:hi].contiguous(), "update_U": step % 2 == 0}
    sh.update_precond_lra_whiten(g[lo:hi].contiguous(), lr=0.1, noise=noise)
    ssq = torch.zeros(1, device=dev)
    out = sh.precond_grad_lra(g[lo:hi].contiguous(), sumsq_out=ssq)
    if rank == 0:
        psgd.update_precond_lra_whiten(whole, Lw, g, lr=0.1, noise={"v": v, "update_U": step % 2 == 0})
        ssq_w = torch.zeros(1, device=dev)
        want = psgd.precond_grad_lra(whole, g, sumsq_out=ssq_w)
        rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
        errs = [rel(sh.UVd[k], whole[k][lo:hi]) for k in range(3)] + [rel(out, want[lo:hi]), rel(ssq, ssq_w)] + [rel(a, b) for a, b in zip(sh.Luvd, Lw)]
        worst = max(worst, max(errs))
        print(f"step {step}: rel err U,V,d,out,sumsq,Lu,Lv,Ld = " + " ".join(f"{e:.2e}" for e in errs), flush=True)
dist.barrier()
if rank == 0:
    print("sharded LRA over NCCL:", "OK" if worst < 5e-3 else "MISMATCH", f"(world {world}, worst {worst:.2e})")
dist.destroy_process_group()
