#!/bin/bash
# Round-2 4-GPU visit: config 5 as specified (256 blocks, 720p, 4-step list, 4 ranks), SP parity tests at 2 / 4 ranks and
# the MAGI Ulysses-CP parity test, and the 4-rank bench line.
set -u
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== sp / cp parity tests"; timeout 600 python -m pytest tests/test_gpu_sp.py -q -k "not 8" 2>&1 | tail -3
echo "=== config 5: long horizon"; timeout 900 $TR --master-port 29541 tools/long_horizon.py --out gpurun_out/r02g_long_horizon_sp${N}.json 2> gpurun_out/r02g_long.err | grep "^{" | cut -c1-1200; tail -2 gpurun_out/r02g_long.err
echo "=== bench N=$N"; timeout 400 $TR --master-port 29561 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_sp${N}.json 2> gpurun_out/r02g_sp${N}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02g_sp${N}.json").read().strip().splitlines()[-1])
    sp = d.get("sp_parity") or {}
    print("N=$N", round(d["value"], 4), "frames/s", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 4), {k: sp.get(k) for k in ("rel_l2", "index_trace_equal")}, d["roofline"]["avg_launch_ms"], d["clocks"])
except Exception as e:
    print("bench failed:", e); import subprocess; print(subprocess.run("tail -8 gpurun_out/r02g_sp${N}.err", shell=True, capture_output=True, text=True).stdout)
PY
