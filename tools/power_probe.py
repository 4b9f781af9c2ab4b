"""Sustained 4096^3 GEMM loops (cuBLAS, 1-CTA, 2-CTA) with nvidia-smi power / clock sampling: is the kernel power-capped or latency-bound?"""
import os, sys, subprocess, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from psgd_torch_b200 import psgd, _lib
dev = torch.device("cuda:0"); lib = _lib.load_library(); h = _lib.handle_for(dev)
A = torch.randn(4096, 4096, device=dev).bfloat16(); B = torch.randn(4096, 4096, device=dev).bfloat16()
C = torch.empty(4096, 4096, device=dev, dtype=torch.bfloat16)

def run(name, fn, secs=3.0):
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=f)
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(200): fn()
        n += 200
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    p.terminate(); p.wait(); f.flush()
    rows = [l.strip().split(", ") for l in open(f.name) if l.strip()][5:]
    clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
    cap = sum(1 for r in rows if r[2].startswith("Active")) / max(1, len(rows))
    print(f"{name:10s}: {ms*1e3:7.1f} us/GEMM  {2*4096**3/ms/1e9:7.1f} TF/s | sm clock median {clk[len(clk)//2]:.0f} MHz | power median {pw[len(pw)//2]:.0f} W max {pw[-1]:.0f} W | sw_power_cap active {cap*100:.0f}% of samples")

run("cuBLAS", lambda: torch.matmul(A, B, out=C))
lib.psgd_debug_set_flags(h, 16); run("1cta", lambda: psgd.gemm(A, B, path=2))
lib.psgd_debug_set_flags(h, 8); run("2cta", lambda: psgd.gemm(A, B, path=2))
lib.psgd_debug_set_flags(h, 0); run("1cta-NT", lambda: psgd.gemm(A, B, trans_b=True, path=2))
