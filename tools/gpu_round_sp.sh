#!/bin/bash
# One multi-GPU box visit (gpurun --gpus N -- 'bash tools/gpu_round_sp.sh N'): SP / CP parity, then the 720p bench with
# the three exchange variants back to back on the same box.  Outputs in gpurun_out/.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "=== parity"; timeout 400 python -m pytest tests/test_gpu_sp.py -q 2>&1 | tail -3
for variant in peer push_overlap nccl; do
  case $variant in
    peer) ENVV="";;
    nccl) ENVV="IFX_SP_PEER=0";;
  esac
  echo "=== bench $variant"
  env $ENVV timeout 300 $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 2 --warmup 3 --no-cpu-baseline \
      > gpurun_out/sp${N}_${variant}.json 2> gpurun_out/sp${N}_${variant}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sp${N}_${variant}.json").read().strip().splitlines()[-1])
    print("$variant", round(d["value"], 4), "frames/s", round(d["ms_per_step"], 1), "ms", d.get("sp_exchange"), d["kv_hbm"].get("peer_wait"))
except Exception as e:
    print("$variant failed:", e)
PY
done
if [ "$N" -ge 8 ]; then
  echo "=== MAGI 24B-width layer, Ulysses CP"
  timeout 200 $TR --master-port 29811 tools/magi_layer_bench.py --model 24b --clip-tokens 48240 --ranges 2 --history 2 \
      > gpurun_out/magi_layer_24b_cp${N}.json 2> gpurun_out/magi_cp.err; tail -c 600 gpurun_out/magi_layer_24b_cp${N}.json
fi
