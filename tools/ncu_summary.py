#!/usr/bin/env python3
"""Compact per-kernel summary of an `ncu -i X.ncu-rep --page raw --csv` export (CPU only).
Usage: ncu_summary.py raw.csv > summary.csv"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
w = csv.writer(sys.stdout)
w.writerow(["kernel", "metric", "value", "unit"])
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = d["Kernel Name"].replace("void ", "").replace("foto::", "").replace("<unnamed>::", "").split("(")[0]
    for k in KEYS:
        if k in d:
            w.writerow([name, k, d[k], u[k]])
