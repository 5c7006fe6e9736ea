#!/usr/bin/env python3
"""Generates tests/golden/lu16384_f64.npz: the reference-order LU (oracle, order-preserving parallel form, proven
bit-identical to the canonical loop nest by tests/test_oracle_forms.py) of the BASELINE config-2 input
(n = 16384, f64, counter-hash seed 1).  Takes ~10-15 minutes on 8 cores; run once, commit the small fixture.

Stored: piv (uint16), pospivsign, the diagonal of U (every 8th entry), three packed-LU rows (every 16th column),
log|det| pieces, and the oracle's backward error ||PA-LU||_F/||A||_F.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
orc.build()
a = orc.fill((n, n), 1)
t0 = time.time()
lu, piv, sign = orc.lu(a, form="fast")
t1 = time.time()
print(f"oracle LU n={n}: {t1 - t0:.1f} s on {orc.num_threads()} threads", flush=True)
be = orc.lu_backward_error(a, lu, piv)
print(f"backward error {be:.3e} ({time.time() - t1:.1f} s)", flush=True)
rows = np.array([1, n // 2 + 3, n - 1])
np.savez_compressed(
    os.path.join(HERE, f"lu{n}_f64.npz"),
    n=n, seed=1, piv=piv.astype(np.uint16 if n <= 65536 else np.uint32), pospivsign=sign,
    diag_every8=np.ascontiguousarray(np.diagonal(lu)[::8]), rows=rows,
    row_samples_every16=np.ascontiguousarray(lu[rows][:, ::16]),
    backward_error=be, oracle_seconds=t1 - t0, oracle_threads=orc.num_threads())
print("saved")
