#!/bin/bash
# Round-2 profiler visit (ONE GPU): launch list of a steady-state slice of the 720p block + `--set full` captures of the
# kernels the roofline discussion names.  Numbers printed by anything under ncu are NOT bench values.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
echo "=== launch list"
timeout 900 $NCU --metrics gpu__time_duration.sum -s 40500 -c 450 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --profile-steps 1 --no-cpu-baseline --no-ref-gpu > gpurun_out/r02_ncu_bench.log 2>&1
tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-200; wc -l gpurun_out/r02_launches.csv
cap() {  # name, kernel regex, probe mode, skip
  echo "=== ncu full $1"
  timeout 600 $NCU --set full --import-source on -k regex:$2 -s $4 -c 1 -f -o gpurun_out/r02_$1_full \
      python tools/gpu_probe.py $3 > gpurun_out/r02_ncu_$1.log 2>&1; tail -1 gpurun_out/r02_ncu_$1.log | cut -c1-160
}
cap attn attn_fwd attn_once 2
cap gemm2_ffn1 gemm2_tn gemm_once 2
cap gemm2_fp8 gemm2_tn gemm_fp8_once 2
cap gemm2_small gemm2_tn gemm_small_once 2
cap qk_norm_rope qk_norm_rope rows_once 2
cap ln_modulate ln_modulate rows_once 2
ls -la gpurun_out/*.ncu-rep
