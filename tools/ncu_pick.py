#!/usr/bin/env python3
"""Prints selected metrics of the first kernel in an .ncu-rep (via `ncu -i ... --page raw --csv`).
usage: ncu_pick.py report.ncu-rep [substring ...]"""
import csv
import io
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum [", "dram__bytes_write.sum [",
           "pipe_tensor_cycles_active_realtime.avg.pct", "sm__ops_path_tensor_src_fp64", "lts__t_sector_hit_rate.pct",
           "launch__registers_per_thread [", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "lts__throughput.avg.pct", "l1tex__throughput.avg.pct", "dram__throughput.avg.pct",
           "sm__throughput.avg.pct", "smsp__inst_executed.sum [", "l1tex__data_bank_conflicts_pipe_lsu.sum [",
           "smsp__warp_issue_stalled", "sm__warps_active.avg.pct", "smsp__issue_active.avg.pct"]
rep = sys.argv[1]
want = sys.argv[2:] or DEFAULT
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:3]:
    for h, u, v in zip(hdr, units, vals):
        key = f"{h} [{u}]"
        if any(w in key for w in want) and v != "":
            print(f"{key} = {v}")
