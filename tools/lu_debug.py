#!/usr/bin/env python3
"""Debug aid: compares the CUDA LU pivots/values with the oracle for several n and reports the first mismatch."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import numpy as np  # noqa: E402
from la._cabi import check, lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402

L = lib()
for n in [int(x) for x in sys.argv[1:]] or [1024, 2048]:
    a = orc.fill((n, n), 1)
    ref_lu, ref_piv, ref_sign = orc.lu(a)
    for rep in range(3):
        lu = np.empty_like(a)
        piv = np.empty(n, dtype=np.uint64)
        sign = ctypes.c_int(0)
        check(L.la_lu_factor_f64_host(a.ctypes.data, lu.ctypes.data, n, n, piv.ctypes.data, ctypes.byref(sign)))
        bad = np.nonzero(piv != ref_piv)[0]
        err = np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), 1.0)
        rows_bad = np.nonzero(err.max(axis=1) > 1e-9)[0]
        cols_bad = np.nonzero(err.max(axis=0) > 1e-9)[0]
        print(f"n={n} rep={rep} LA_LU_DEBUG={os.environ.get('LA_LU_DEBUG')}: piv mismatches {bad.size} first {bad[:3]}, "
              f"max err {err.max():.2e}, bad rows {rows_bad[:4]} ({rows_bad.size}), bad cols {cols_bad[:4]} ({cols_bad.size})",
              flush=True)
