#!/bin/bash
# 2-GPU A/B of the exchange variants with the final kernels: gpurun --gpus 2 -- bash tools/gpu_round2_sp2ab.sh
set -x
mkdir -p gpurun_out
O=gpurun_out
for mode in overlap store; do
  IFX_SP_MODE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-sp-parity > $O/r02m_sp2_$mode.json 2> $O/r02m_sp2_$mode.err
  tail -c 600 $O/r02m_sp2_$mode.err
done
python - <<'P'
import json
for m in ("overlap", "store"):
    for line in open(f"gpurun_out/r02m_sp2_{m}.json"):
        if line.startswith("{"):
            d = json.loads(line)
            print(m, d["ms_per_step"], d["value"], d["roofline"]["avg_launch_ms"], d["clocks"]["sm_mhz"], d["sp_exchange"][:30])
P
