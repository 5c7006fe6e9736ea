#!/usr/bin/env python3
"""Times la_qr_factor_f64_dev (and optionally get_q) on the seeded n x n matrix.  usage: qr_profile.py [m] [n] [reps] [--q]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from gpu_util import DevBuf, fill_hash, sync  # noqa: E402
from la._cabi import check, lib  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
m = int(args[0]) if len(args) > 0 else 16384
n = int(args[1]) if len(args) > 1 else m
reps = int(args[2]) if len(args) > 2 else 2
L = lib()
a = DevBuf(m * n * 8)
te = ctypes.c_size_t(0)
check(L.la_qr_tmat_elems(m, n, 0, 8, ctypes.byref(te)))
rd = DevBuf(min(m, n) * 8)
tm = DevBuf(te.value * 8)
flops = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3 if m >= n else 2.0 * n * m * m - 2.0 / 3.0 * m ** 3
for r in range(reps):
    fill_hash(a, m * n, 1, np.float64)
    sync()
    t0 = time.perf_counter()
    check(L.la_qr_factor_f64_dev(a.ptr(), m, n, rd.ptr(), tm.ptr(), None))
    sync()
    t1 = time.perf_counter()
    print(f"qr {m}x{n}: {1e3 * (t1 - t0):.2f} ms ({flops / (t1 - t0) / 1e12:.2f} TFLOP/s, flops = 2mn^2 - 2/3 n^3)", flush=True)
if "--q" in sys.argv:
    q = DevBuf(m * m * 8)
    qh = ctypes.c_void_p()
    for r in range(2):
        sync()
        t0 = time.perf_counter()
        check(L.la_qr_get_q_f64(a.h, m, n, tm.h, q.h))
        sync()
        t1 = time.perf_counter()
        print(f"get_q {m}x{m}: {1e3 * (t1 - t0):.2f} ms")
