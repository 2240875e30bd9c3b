#!/bin/bash
# Round-2 8-GPU visit: SP parity at 8 ranks, then the bench with the fused exchange vs the store+wait exchange, same box.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== sp parity (tiny pipeline, $N ranks, default mode)"; timeout 300 $TR --master-port 29533 tools/sp_check.py 2>gpurun_out/r02e_spcheck.err | grep "^{" | tee gpurun_out/r02e_sp_check_${N}.json; tail -2 gpurun_out/r02e_spcheck.err
echo "=== interface tests"; timeout 300 python -m pytest tests/test_gpu_interfaces.py tests/test_gpu_paged_attention.py -q 2>&1 | tail -3
for variant in overlap store; do
  echo "=== bench $variant"
  env IFX_SP_MODE=$variant timeout 400 $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r02e_sp${N}_${variant}.json 2> gpurun_out/r02e_sp${N}_${variant}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02e_sp${N}_${variant}.json").read().strip().splitlines()[-1])
    sp = d.get("sp_parity") or {}
    print("$variant", round(d["value"], 4), "frames/s", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 4), {k: sp.get(k) for k in ("rel_l2", "bit_equal", "index_trace_equal")}, {k: v for k, v in d["kv_hbm"].items() if k.startswith("peer") or k.startswith("fused") or k.startswith("append")}, d["roofline"]["share_of_step"], d["roofline"]["avg_launch_ms"], d["clocks"])
except Exception as e:
    print("$variant failed:", e)
    import subprocess; print(subprocess.run("tail -8 gpurun_out/r02e_sp${N}_${variant}.err", shell=True, capture_output=True, text=True).stdout)
PY
done
