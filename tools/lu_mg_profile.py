#!/usr/bin/env python3
"""Times the multi-device LU (la_lu_mg_*) on the seeded n x n matrix: device time from the context's own events.
usage: lu_mg_profile.py n reps dev[,dev...] [dev[,dev...] ...] [--check]   e.g.  lu_mg_profile.py 16384 3 0 0,0 0,1 0,1,2,3
--check (n = 16384 only): the pivot permutation of the last factorisation against the committed oracle fixture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import numpy as np  # noqa: E402
from la import sharding  # noqa: E402

n = int(sys.argv[1])
reps = int(sys.argv[2])
check = "--check" in sys.argv
for spec in [a for a in sys.argv[3:] if not a.startswith("--")]:
    devs = [int(x) for x in spec.split(",")]
    ctx = sharding.LuMgContext(devs, n)
    times = []
    for _ in range(reps):
        ctx.fill_hash(1)
        ctx.sync()
        ctx.factor()
        times.append(ctx.last_ms())
    verdict = ""
    if check and n == 16384:
        fx = np.load(os.path.join(ROOT, "tests", "golden", "lu16384_f64.npz"))
        _, piv, sign = ctx.download(want_lu=False)
        same = bool(np.array_equal(piv.astype(np.int64), fx["piv"].astype(np.int64))) and sign == bool(fx["pospivsign"])
        verdict = f"; piv_identical_to_oracle_fixture={same}"
    ctx.destroy()
    best = min(times)
    print(f"lu_mg n={n} devices={devs}: " + " ".join(f"{t:.2f}" for t in times) +
          f" ms; best {best:.2f} ms = {2 / 3 * n ** 3 / best / 1e9:.2f} TFLOP/s{verdict}", flush=True)
