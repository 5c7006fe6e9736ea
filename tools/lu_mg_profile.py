#!/usr/bin/env python3
"""Times the multi-device LU (la_lu_mg_*) on the seeded n x n matrix: device time from the context's own events.
usage: lu_mg_profile.py n reps dev[,dev...] [dev[,dev...] ...]   e.g.  lu_mg_profile.py 16384 3 0 0,0 0,1 0,1,2,3"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
from la import sharding  # noqa: E402

n = int(sys.argv[1])
reps = int(sys.argv[2])
for spec in sys.argv[3:]:
    devs = [int(x) for x in spec.split(",")]
    ctx = sharding.LuMgContext(devs, n)
    times = []
    for _ in range(reps):
        ctx.fill_hash(1)
        ctx.sync()
        ctx.factor()
        times.append(ctx.last_ms())
    ctx.destroy()
    best = min(times)
    print(f"lu_mg n={n} devices={devs}: " + " ".join(f"{t:.2f}" for t in times) +
          f" ms; best {best:.2f} ms = {2 / 3 * n ** 3 / best / 1e9:.2f} TFLOP/s", flush=True)
