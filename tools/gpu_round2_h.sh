#!/bin/bash
# Round-2 8-GPU visit #2: BASELINE config 4 shapes (MAGI-1 1080p, Ulysses CP over 8 ranks, 32-chunk cache) on one layer at
# 4.5B and 24B widths, then the 720p bench with fewer CTAs sharing the fused exchange.
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for m in 4.5b 24b; do
  echo "=== MAGI-1 $m layer, 1080p, cp_ulysses $N, 28 cached + 4 denoising chunks"
  timeout 300 $TR --master-port $((29700 + RANDOM % 100)) tools/magi_layer_bench.py --model $m --clip-tokens 48240 --ranges 4 --history 28 --reps 3 \
      > gpurun_out/r02h_magi_layer_${m}_1080p_cp${N}.json 2> gpurun_out/r02h_magi_${m}.err
  grep "^{" gpurun_out/r02h_magi_layer_${m}_1080p_cp${N}.json | cut -c1-900; tail -2 gpurun_out/r02h_magi_${m}.err | cut -c1-300
done
for ctas in 48 16; do
  echo "=== bench fused exchange, push CTAs $ctas"
  env IFX_SP_PUSH_CTAS=$ctas timeout 400 $TR --master-port $((29600 + RANDOM % 100)) bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --no-sp-parity \
      > gpurun_out/r02h_sp${N}_push${ctas}.json 2> gpurun_out/r02h_sp${N}_push${ctas}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02h_sp${N}_push${ctas}.json").read().strip().splitlines()[-1])
    print("push_ctas=$ctas", round(d["value"], 4), "frames/s", round(d["ms_per_step"], 1), "ms e2e", round(d["e2e"]["value"], 4), d["roofline"]["avg_launch_ms"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("failed:", e); import subprocess; print(subprocess.run("tail -8 gpurun_out/r02h_sp${N}_push${ctas}.err", shell=True, capture_output=True, text=True).stdout)
PY
done
