#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r02f_pytest_gpu.log 2>&1; tail -12 gpurun_out/r02f_pytest_gpu.log | cut -c1-300
for wlk in causvid_540p_fp8 causvid_540p_fp8_dynamic causvid_540p_int8_dynamic; do
  echo "=== bench $wlk"; timeout 300 python bench.py --workload $wlk --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_$wlk.json 2> gpurun_out/r02f_bench_$wlk.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02f_bench_$wlk.json").read().strip().splitlines()[-1])
    print("$wlk", round(d["value"], 3), "frames/s", round(d["ms_per_step"], 1), "ms/block", d["dtype"], d["roofline"]["share_of_step"] if d.get("roofline") else None)
except Exception as e:
    print("$wlk failed", e); import subprocess; print(subprocess.run("tail -5 gpurun_out/r02f_bench_$wlk.err", shell=True, capture_output=True, text=True).stdout)
PY
done
