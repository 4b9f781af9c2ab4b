#!/bin/bash
# round-2 profiling batch (run under gpurun): bucket timings, launch list and ncu full of the bounds kernel
mkdir -p gpurun_out
python tools/time_buckets.py q_o_proj gate_up_proj k_v_proj > gpurun_out/r02_buckets_fused.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_kron4096.csv python tools/profile_kron.py 4096 4096 > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_kron14336x4096.csv python tools/profile_kron.py 14336 4096 > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_norm_bounds -c 1 -o gpurun_out/r02_ncu_bounds python tools/profile_kron.py 4096 4096 > gpurun_out/ncu_bounds.log 2>&1
cat gpurun_out/r02_buckets_fused.log
grep -E "k_norm_bounds" gpurun_out/r02_launches_kron4096.csv gpurun_out/r02_launches_kron14336x4096.csv | awk -F'","' '{print $5, $NF}'
