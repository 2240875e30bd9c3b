#!/bin/bash
# Round-2 multi-GPU visit #2: fused exchange.  SP parity tests, then bench overlap vs store on the same box, then N=1.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_gpu.log 2>&1; tail -6 gpurun_out/r02c_pytest_gpu.log
echo "=== sp shape"; timeout 200 python tools/sp_shape_bench.py --rows 10800,1350 > gpurun_out/r02c_sp_shape.jsonl 2>&1; python - <<'PY'
import json
for l in open("gpurun_out/r02c_sp_shape.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["M"], d["layer_us"], "attn", d["attn_self_us"], "gemm", d["gemm_us"], {k.split("[")[0][:28]: v for k, v in d["kernels_us"].items()})
PY
for variant in overlap store; do
  echo "=== bench $variant"
  env IFX_SP_MODE=$variant timeout 400 $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02c_sp${N}_${variant}.json 2> gpurun_out/r02c_sp${N}_${variant}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02c_sp${N}_${variant}.json").read().strip().splitlines()[-1])
    print("$variant", round(d["value"], 4), "frames/s", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 4), d.get("sp_parity"), {k: v for k, v in d["kv_hbm"].items() if k.startswith("peer") or k.startswith("fused") or k.startswith("append")}, d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"])
except Exception as e:
    print("$variant failed:", e)
    import subprocess; print(subprocess.run("tail -8 gpurun_out/r02c_sp${N}_${variant}.err", shell=True, capture_output=True, text=True).stdout)
PY
done
echo "=== bench N=1"; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-gpu > gpurun_out/r02c_bench_720p.json 2> gpurun_out/r02c_bench_720p.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02c_bench_720p.json").read().strip().splitlines()[-1])
print("N=1", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["share_of_step"], d["kv_hbm"]["append_norm_rope"], d["launches_per_step"])
PY
