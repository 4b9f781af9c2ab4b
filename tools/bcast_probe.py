"""NCCL broadcast probe for the sharded KWNS4 step (run under torchrun): the Llama-3-8B parameter list (bf16, 16 GB) is broadcast
parameter by parameter from round-robin owners, async on NCCL's stream, exactly as KWNS4(shard_preconditioners=True).step() posts
them -- with no compute beside them.  Prints ms per pass and the per-rank ingress rate; NCCL_MAX_NCHANNELS comes from the environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
params = []
for name, count, shape, kind in bench.LLAMA3_8B_SET:
    numel = 1
    for s in (shape if kind != "lra" else (shape[0],)):
        numel *= s
    for _ in range(count):
        params.append(torch.zeros(numel, dtype=torch.bfloat16, device="cuda"))
total = sum(p.numel() * 2 for p in params)


def one_pass():
    pend = [dist.broadcast(p, src=i % world, async_op=True) for i, p in enumerate(params)]
    for w in pend:
        w.wait()


for _ in range(2):
    one_pass()
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    one_pass()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 3], device="cuda")
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"NCCL_MAX_NCHANNELS={os.environ.get('NCCL_MAX_NCHANNELS', 'default')}: {world} ranks, {len(params)} broadcasts, {total/1e9:.2f} GB: "
          f"{ms.item():.1f} ms per pass, {total * (world - 1) / world / ms.item() / 1e6:.0f} GB/s ingress per rank")
dist.destroy_process_group()
