#!/bin/bash
# multi-GPU (run under gpurun --gpus N): the sharded-KWNS4 bench (BASELINE configs[3]) per exchange scheme / SMs left to NCCL, the sharded
# correctness check, and the parameter broadcasts alone
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29614 tools/check_sharded_kwns4.py > gpurun_out/r02_sharded_kwns4_n$N.log 2>&1
tail -4 gpurun_out/r02_sharded_kwns4_n$N.log
for CFG in ${CFGS:-p2p_0}; do
  X=${CFG%_*}; C=${CFG##*_}
  F=gpurun_out/r02_bench_kwns4_n${N}_${X}_c$C
  timeout 300 $TR --master-port 29622 bench.py --mode kwns4 --gpus $N --steps 5 --warmup 3 --comm-sms $C --exchange $X > $F.json 2> $F.err
  python - <<PY
import json
try:
    d = json.loads(open("$F.json").read().strip().splitlines()[-1])
    print("$X comm_sms $C: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1), "all-gather bytes", d["config"].get("nccl_all_gather_buffer_bytes_per_step"))
except Exception as e:
    print("$X comm_sms $C unreadable", e); print(open("$F.err").read()[-1500:])
PY
done
