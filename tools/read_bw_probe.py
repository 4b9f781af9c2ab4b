"""What can a pure-read HBM stream reach on this B200?  (yardstick for the read-only LRA sweeps; MEASURED_PEAKS.json's figure is a copy,
i.e. half reads and half writes).  Sweeps grid size / threads / loads in flight of a trivial read kernel and compares with torch reductions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import _lib

dev = torch.device("cuda:0")
lib = _lib.load_library()
h = _lib.handle_for(dev)
nbytes = 8 << 30
x = torch.empty(nbytes, dtype=torch.uint8, device=dev)
x.view(torch.int32).random_(0, 1 << 30)
scratch = torch.zeros(16, dtype=torch.int32, device=dev)
sms = torch.cuda.get_device_properties(0).multi_processor_count


def timeit(fn, iters=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


y = torch.empty_like(x)
t = timeit(lambda: y.copy_(x))
print(f"torch copy (read + write bytes): {2 * nbytes / t / 1e6:.0f} GB/s")
del y
t = timeit(lambda: x.view(torch.bfloat16).sum(dtype=torch.float32))
print(f"torch bf16 sum (read only):      {nbytes / t / 1e6:.0f} GB/s")
t = timeit(lambda: x.view(torch.int32).max())
print(f"torch int32 max (read only):     {nbytes / t / 1e6:.0f} GB/s")
best = 0
for noalloc in (0, 1):
    for threads in (256, 512):
        for bps in (1, 2, 4, 8):
            for unroll in (1, 4, 8, 16):
                if bps * threads > 2048: continue
                fn = lambda: _lib.check(h, lib.psgd_debug_read_probe(h, _lib.ptr(x), nbytes, sms * bps, threads, unroll, noalloc, _lib.ptr(scratch),
                                                                     _lib.stream_ptr(dev)), "probe")
                t = timeit(fn)
                bw = nbytes / t / 1e6
                best = max(best, bw)
                print(f"  noalloc={noalloc} threads={threads} blocks/SM={bps} loads in flight/thread={unroll:2d} -> {threads*bps*unroll*16/1024:6.0f} KB/SM in flight: {bw:6.0f} GB/s")
print(f"best pure-read stream: {best:.0f} GB/s")
