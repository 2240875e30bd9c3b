#!/bin/bash
# End-of-round validation on ONE GPU: GPU parity tests, build + smoke, the default bench line (what the driver runs),
# ncu --set full of the MAGI row kernels.  gpurun --timeout 1500 -- bash tools/gpu_round2_final.sh
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s -x > $O/r02n_pytest_gpu.log 2>&1; tail -3 $O/r02n_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/r02n_smoke.log 2>&1; tail -3 $O/r02n_smoke.log
timeout 900 python bench.py > $O/r02n_bench_720p.json 2> $O/r02n_bench.err; tail -c 400 $O/r02n_bench.err; cut -c1-600 $O/r02n_bench_720p.json
for k in magi_qkv_post_kernel head_layernorm_kernel gate_norm_residual_kernel silu_mul_kernel; do
  timeout 300 ncu --clock-control none --set full --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r02n_${k}_full \
      python tools/magi_layer_bench.py --model 24b --reps 1 > $O/r02n_ncu_$k.log 2>&1
  tail -1 $O/r02n_ncu_$k.log | cut -c1-200
done
ls -la $O/r02n_*
