#!/usr/bin/env python3
"""Runs one LU factorisation (+ optional solve) of the seeded n x n matrix through the C ABI; meant to be wrapped in
ncu for the per-launch time list.  usage: lu_profile.py [n] [reps] [--solve] [--f32]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
do_solve = "--solve" in sys.argv
f32 = "--f32" in sys.argv
dt = np.float32 if f32 else np.float64
es = 4 if f32 else 8
L = lib()
a = DevBuf(n * n * es)
piv = DevBuf(n * 8 + 64)
b = DevBuf(n * 16 * es)
x = DevBuf(n * 16 * es)
fill_hash(b, n * 16, 3, dt)
factor = L.la_lu_factor_f32_dev if f32 else L.la_lu_factor_f64_dev
solve = L.la_lu_solve_f32_dev if f32 else L.la_lu_solve_f64_dev
for r in range(reps):
    fill_hash(a, n * n, 1, dt)
    sync()
    t0 = time.perf_counter()
    check(factor(a.ptr(), n, n, piv.ptr(), piv.ptr(n * 8), None))
    sync()
    t1 = time.perf_counter()
    if do_solve:
        check(solve(a.ptr(), n, piv.ptr(), b.ptr(), 16, x.ptr(), None))
        sync()
    t2 = time.perf_counter()
    print(f"n={n} lu {1e3 * (t1 - t0):.2f} ms ({2 / 3 * n ** 3 / (t1 - t0) / 1e12:.2f} TFLOP/s) solve {1e3 * (t2 - t1):.2f} ms")
if "--inverse" in sys.argv:
    # many right-hand sides: X = A^-1 for the last factorisation (nx = n), the path Matrix::inverse takes
    eye = DevBuf(n * n * es)
    inv = DevBuf(n * n * es)
    check((L.la_identity_f32 if f32 else L.la_identity_f64)(eye.h, n))
    for r in range(2):
        sync()
        t0 = time.perf_counter()
        check(solve(a.ptr(), n, piv.ptr(), eye.ptr(), n, inv.ptr(), None))
        sync()
        t1 = time.perf_counter()
        print(f"n={n} inverse (solve with n right-hand sides) {1e3 * (t1 - t0):.2f} ms ({2.0 * n ** 3 / (t1 - t0) / 1e12:.2f} TFLOP/s)")
