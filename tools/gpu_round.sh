#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, ncu launch list + full captures.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench tiny"; timeout 300 python bench.py --workload tiny --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
echo "=== bench 720p"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_720p.json 2> gpurun_out/bench_720p.err; tail -c 3000 gpurun_out/bench_720p.json; tail -5 gpurun_out/bench_720p.err
if [ "${1:-}" = "ncu" ]; then
  echo "=== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 13500 -c 420 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log | cut -c1-300
  echo "=== ncu full attn"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 2 -c 1 -f -o gpurun_out/attn_full \
      python tools/gpu_probe.py attn_once > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
  echo "=== ncu full gemm"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/gemm_full \
      python tools/gpu_probe.py gemm_once > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
fi
