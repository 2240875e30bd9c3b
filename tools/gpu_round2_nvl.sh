#!/bin/bash
# 2-GPU visit (single process): NVLink bytes of the K/V exchange under ncu + device times of rank 0's layer with the
# exchange going to a real peer GPU.  gpurun --gpus 2 -- bash tools/gpu_round2_nvl.sh
set -x
mkdir -p gpurun_out
O=gpurun_out
M=gpu__time_duration.sum,nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
timeout 300 python tools/nvlink_probe.py --mode overlap > $O/r02l_nvlink_probe_overlap.json 2> $O/r02l_probe_overlap.err
timeout 300 python tools/nvlink_probe.py --mode store > $O/r02l_nvlink_probe_store.json 2> $O/r02l_probe_store.err
timeout 300 python tools/sp_shape_bench.py --rows 1350 > $O/r02l_sp_shape_1350.jsonl 2> $O/r02l_sp_shape.err
timeout 600 ncu --metrics $M --clock-control none -k regex:'attn_fwd_kernel|qk_norm_rope' --csv \
    --log-file $O/r02l_nvlink_ncu_overlap.csv python tools/nvlink_probe.py --mode overlap --iters 1 > $O/r02l_ncu_overlap.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'attn_fwd_kernel|qk_norm_rope' --csv \
    --log-file $O/r02l_nvlink_ncu_store.csv python tools/nvlink_probe.py --mode store --iters 1 > $O/r02l_ncu_store.log 2>&1
timeout 600 python -m pytest tests/test_gpu_sp.py -m gpu -x -q -s > $O/r02l_pytest_sp2.log 2>&1
tail -3 $O/r02l_pytest_sp2.log
cat $O/r02l_nvlink_probe_overlap.json $O/r02l_nvlink_probe_store.json
tail -5 $O/r02l_probe_overlap.err
head -c 1500 $O/r02l_nvlink_ncu_overlap.csv
