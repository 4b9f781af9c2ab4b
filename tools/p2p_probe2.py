"""Peer-copy probe 2 (torchrun, N >= 2): rank r allocates a staging tensor ON THE NEXT RANK'S GPU, pushes into it (peer copy), and the
next rank maps that staging tensor through CUDA IPC (device-local memory of another process) and copies out of it."""
import os
import torch
import torch.distributed as dist
from torch.multiprocessing import reductions

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nxt, prv = (local + 1) % world, (local - 1) % world
n = 512 * 1024 * 1024
src = torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device=f"cuda:{local}")
stage_on_next = torch.zeros(n, dtype=torch.bfloat16, device=f"cuda:{nxt}")          # my allocation, on the next rank's GPU
gathered = [None] * world
dist.all_gather_object(gathered, reductions.reduce_tensor(stage_on_next)[1])
from_prev = reductions.rebuild_cuda_tensor(*gathered[prv])                          # the previous rank's staging tensor on MY GPU
dst = torch.zeros(n, dtype=torch.bfloat16, device=f"cuda:{local}")
print(f"rank {rank}: staging of rank {prv} mapped on {from_prev.device}", flush=True)


def timed(fn, reps=4):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return reps * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9


timed(lambda: stage_on_next.copy_(src, non_blocking=True))
push = timed(lambda: stage_on_next.copy_(src, non_blocking=True))
dist.barrier(); torch.cuda.synchronize()
timed(lambda: dst.copy_(from_prev, non_blocking=True))
pull = timed(lambda: dst.copy_(from_prev, non_blocking=True))
ok = bool((dst[:4096] == float(prv + 1)).all() and (dst[-4096:] == float(prv + 1)).all())
print(f"rank {rank}: push into own allocation on the peer GPU {push:.0f} GB/s; local copy out of the IPC-mapped staging {pull:.0f} GB/s; data correct: {ok}", flush=True)
dist.barrier()
del from_prev
dist.destroy_process_group()
