"""Peer-copy probe (torchrun, N >= 2): bandwidth of torch's cross-device copy_ from this rank's GPU into (a) a tensor this process
allocated on the next GPU, (b) the next rank's tensor mapped through CUDA IPC -- alone and beside a matmul loop."""
import os, sys, time
import torch
import torch.distributed as dist
from torch.multiprocessing import reductions

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nxt = (local + 1) % world
n = 512 * 1024 * 1024
src = torch.ones(n, dtype=torch.bfloat16, device=f"cuda:{local}")
mine = torch.zeros(n, dtype=torch.bfloat16, device=f"cuda:{local}")
args = reductions.reduce_tensor(mine)[1]
gathered = [None] * world
dist.all_gather_object(gathered, args)
ipc_view = reductions.rebuild_cuda_tensor(*gathered[nxt])
own_remote = torch.zeros(n, dtype=torch.bfloat16, device=f"cuda:{nxt}")
print(f"rank {rank}: can_access_peer({local},{nxt}) = {torch.cuda.can_device_access_peer(local, nxt)}, ipc view on {ipc_view.device}", flush=True)
side = torch.cuda.Stream()
a = torch.randn(8192, 8192, device=f"cuda:{local}", dtype=torch.bfloat16)


def timed(dst, with_compute):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side):
        e0.record()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        e1.record()
    if with_compute:
        for _ in range(40):
            a @ a
    torch.cuda.synchronize()
    return 4 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9


for name, dst in (("own allocation on the peer GPU", own_remote), ("peer's tensor through CUDA IPC", ipc_view)):
    timed(dst, False)
    bw0, bw1 = timed(dst, False), timed(dst, True)
    print(f"rank {rank}: {name}: {bw0:.0f} GB/s alone, {bw1:.0f} GB/s beside matmuls", flush=True)
dist.barrier()
ok = bool((mine[:1024] == 1).all())
print(f"rank {rank}: peer wrote into my tensor: {ok}", flush=True)
dist.barrier()
del ipc_view
dist.destroy_process_group()
