#!/usr/bin/env python3
"""Summarises the per-panel timeline written by LA_LU_TRACE=<file> (lu.cu): chain segments vs bulk segments, in ms."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/lu_trace.csv")))[1:]
R = [[float(x) for x in r] for r in rows]
every = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print("  p panel_done  perm+wait  head  headupd  panel_next | bulk_start  swaps   gemm  bulk_len |  step")
for i, r in enumerate(R):
    p, pd, pb, hr, hu, bs, bw, bd = r
    nxt = R[i + 1][1] if i + 1 < len(R) else float("nan")
    if i % every == 0 or i > len(R) - 4:
        print(f"{int(p):3d} {pd:9.3f} {pb - pd:9.3f} {hr - pb:6.3f} {hu - hr:7.3f} {nxt - hu:10.3f} | {bs:9.3f} "
              f"{bw - bs:6.3f} {bd - bw:6.3f} {bd - bs:8.3f} | {nxt - pd:6.3f}")
