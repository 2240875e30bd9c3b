#!/bin/bash
# Round-2 single-GPU visit: parity (incl. full-shape + paged attention), SP-shape microbench, bench line, reference-on-GPU.
set -u
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02a_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02a_pytest_gpu.log
echo "=== sp shape"; timeout 300 python tools/sp_shape_bench.py > gpurun_out/r02a_sp_shape.jsonl 2> gpurun_out/r02a_sp_shape.err; cut -c1-400 gpurun_out/r02a_sp_shape.jsonl; tail -3 gpurun_out/r02a_sp_shape.err
echo "=== bench 720p"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench_720p.json 2> gpurun_out/r02a_bench_720p.err; tail -c 4000 gpurun_out/r02a_bench_720p.json; tail -5 gpurun_out/r02a_bench_720p.err
echo "=== ref gpu offload=1"; timeout 400 python tools/ref_gpu_bench.py --offload 1 --out gpurun_out/r02a_ref_gpu_offload1.json 2>&1 | tail -3
