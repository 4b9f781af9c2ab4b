"""Pinned-host <-> device copy bandwidth with 1 ... N ranks copying at once (the e2e leg of bench.py stops scaling
at N > 1; this names the host-side limiter).  Each rank copies `mb` MiB each way `reps` times on its own GPU, first alone (ranks take
turns), then all together.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/h2d_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mb, reps = 1024, 8
h_in = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
h_out = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(both):
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h_out.copy_(d, non_blocking=True)
    torch.cuda.synchronize(dev)
    return reps * mb / 1024 / (time.perf_counter() - t0)   # GiB/s per direction


run(True)
alone = []
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        alone = [run(False), run(True)]
if world > 1:
    dist.barrier()
together = [run(False), run(True)]
res = torch.tensor(alone + together, device=dev)
allr = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    print(f"pinned copy probe, {world} rank(s), {mb} MiB x {reps}; GiB/s per rank and direction")
    print("rank  alone:H2D  alone:H2D+D2H  together:H2D  together:H2D+D2H")
    for r, x in enumerate(allr):
        print(f"{r:4d} {x[0]:10.1f} {x[1]:14.1f} {x[2]:13.1f} {x[3]:17.1f}")
    tot = torch.stack(allr).sum(0)
    print(f" sum {float(tot[0]):10.1f} {float(tot[1]):14.1f} {float(tot[2]):13.1f} {float(tot[3]):17.1f}")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
