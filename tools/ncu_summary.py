"""Summarise an .ncu-rep (first kernel) into a small JSON: the metrics the roofline discussion needs."""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "launch__shared_mem_per_block_dynamic",
]


def main(rep, out=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in txt.splitlines() if l.startswith('"')]))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"report": rep, "kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": vals[i], "unit": units[i]}
    extra = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_elapsed" in h and h not in d]
    for h in extra:
        i = hdr.index(h)
        d[h] = {"value": vals[i], "unit": units[i]}
    print(json.dumps(d, indent=1))
    if out:
        json.dump(d, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:])
