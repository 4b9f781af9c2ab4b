#!/bin/bash
# multi-GPU batch (run under gpurun --gpus N): kwns4-mode bench, functional bench, pinned-copy probe
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 bench.py --mode kwns4 --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_kwns4_n$N.json 2> gpurun_out/r02_bench_kwns4_n$N.err
tail -c 300 gpurun_out/r02_bench_kwns4_n$N.err
timeout 400 $TR --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 300 gpurun_out/r02_bench_n$N.err
timeout 200 $TR --master-port 29613 tools/h2d_probe.py > gpurun_out/r02_h2d_probe_n$N.log 2>&1
if [ "$N" -le 2 ]; then
  timeout 200 $TR --master-port 29614 tools/check_sharded_kwns4.py > gpurun_out/r02_sharded_kwns4_n$N.log 2>&1
  tail -4 gpurun_out/r02_sharded_kwns4_n$N.log
fi
python - <<PY
import json
for f in ("gpurun_out/r02_bench_kwns4_n$N.json", "gpurun_out/r02_bench_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/r02_h2d_probe_n$N.log
