#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02d_pytest_gpu.log 2>&1; tail -8 gpurun_out/r02d_pytest_gpu.log
echo "=== sp shape (plan on)"; timeout 300 python tools/sp_shape_bench.py --rows 10800,2700,1350 > gpurun_out/r02d_sp_shape.jsonl 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/r02d_sp_shape.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["M"], d["layer_us"], "attn", d["attn_self_us"], d["attn_tflops"], "gemm", d["gemm_us"])
PY
echo "=== sp shape (plan off)"; IFX_ATTN_PLAN=0 timeout 300 python tools/sp_shape_bench.py --rows 2700,1350 > gpurun_out/r02d_sp_shape_noplan.jsonl 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/r02d_sp_shape_noplan.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["M"], d["layer_us"], "attn", d["attn_self_us"], d["attn_tflops"], "gemm", d["gemm_us"])
PY
echo "=== bench N=1"; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/r02d_bench_720p.json 2> gpurun_out/r02d_bench_720p.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02d_bench_720p.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["kv_hbm"]["append_norm_rope"], d["launches_per_step"])
PY
