#!/bin/bash
# round-2 closing batch on one B200 (run under gpurun): full GPU test suite (writes gpurun_out/parity_errors.log), smoke, the three bench
# lines, the ncu launch list of the bench's timed region and one ncu --set full of the LRA sweep-1 kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
timeout 600 python bench.py --mode kwns4 --steps 5 > gpurun_out/r02_bench_kwns4_n1.json 2> gpurun_out/r02_bench_kwns4_n1.err
python tools/time_buckets.py --batched > gpurun_out/r02_buckets_batched.log 2>&1
python tools/time_buckets.py > gpurun_out/r02_buckets_unbatched.log 2>&1
timeout 300 python tools/lra_gpt2_sweep.py > gpurun_out/r02_lra_gpt2_sweep.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_value.csv \
  python bench.py --steps 1 --warmup 3 --profile-range --no-cpu-baseline --no-gpu-reference > gpurun_out/r02_bench_under_ncu.json 2> /dev/null
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_lra_gram_tc -c 1 -o gpurun_out/r02_ncu_lra_gram_tc \
  python tools/profile_lra.py 67108864 32 > /dev/null 2>&1
python - <<'PY'
import json
for f in ("r02_bench_n1", "r02_bench_reference", "r02_bench_kwns4_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 2),
              "launches", d.get("gpu_launches"), "roofline", (d.get("roofline") or {}).get("frac"), "gpu_ref", (d.get("gpu_reference") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
