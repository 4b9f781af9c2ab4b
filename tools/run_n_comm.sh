#!/bin/bash
# multi-GPU (run under gpurun --gpus N): the sharded-KWNS4 bench against the SMs left to NCCL, and the broadcasts alone
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for C in default 8 16 32; do
  if [ "$C" = default ]; then E=""; else E="NCCL_MAX_NCHANNELS=$C"; fi
  env $E timeout 120 $TR --master-port 29621 tools/bcast_probe.py 2>&1 | grep NCCL_MAX >> gpurun_out/r02_bcast_probe_n$N.log
done
cat gpurun_out/r02_bcast_probe_n$N.log
for C in 16 32; do
  timeout 300 $TR --master-port 29622 bench.py --mode kwns4 --gpus $N --steps 5 --warmup 3 --comm-sms $C > gpurun_out/r02_bench_kwns4_n${N}_c$C.json 2> gpurun_out/r02_bench_kwns4_n${N}_c$C.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_kwns4_n${N}_c$C.json").read().strip().splitlines()[-1])
    print("comm_sms $C: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1))
except Exception as e:
    print("comm_sms $C unreadable", e)
PY
done
