#!/bin/bash
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02k_pytest_gpu.log | cut -c1-250
echo "=== sp shape (slack fill on)"; timeout 300 python tools/sp_shape_bench.py --rows 2700,1350 > gpurun_out/r02k_sp_shape_slackfill.jsonl 2>&1
echo "=== sp shape (slack fill off)"; IFX_ATTN_SLACK_FILL=0 timeout 300 python tools/sp_shape_bench.py --rows 2700,1350 > gpurun_out/r02k_sp_shape_plain.jsonl 2>&1
python - <<'PY'
import json
for f in ("slackfill", "plain"):
    for l in open(f"gpurun_out/r02k_sp_shape_{f}.jsonl"):
        if l.startswith("{"):
            d = json.loads(l); print(f, d["M"], "layer", d["layer_us"], "attn", d["attn_self_us"], d["attn_tflops"], "gemm", d["gemm_us"])
PY
