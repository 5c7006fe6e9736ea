#!/usr/bin/env python3
"""Debug aid: truncates the LU loop after k iterations in serial (LA_LU_DEBUG=1) and look-ahead (0) mode and reports where
the partially factored matrices first differ."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import numpy as np  # noqa: E402

if len(sys.argv) > 1 and sys.argv[1] == "child":
    n, out = int(sys.argv[2]), sys.argv[3]
    from la._cabi import check, lib
    from oracle import oracle as orc
    a = orc.fill((n, n), 1)
    lu = np.empty_like(a)
    piv = np.empty(n, dtype=np.uint64)
    sign = ctypes.c_int(0)
    check(lib().la_lu_factor_f64_host(a.ctypes.data, lu.ctypes.data, n, n, piv.ctypes.data, ctypes.byref(sign)))
    np.save(out, lu)
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
for stop in range(1, n // 128 + 1):
    res = {}
    bad = None
    for trial in range(3):
        for dbg in (1, 0):
            env = dict(os.environ, LA_LU_DEBUG=str(dbg), LA_LU_STOP=str(stop))
            out = f"/tmp/lu_{dbg}.npy"
            subprocess.check_call([sys.executable, __file__, "child", str(n), out], env=env)
            res[dbg] = np.load(out)
        d = res[0] != res[1]
        if d.any():
            bad = d
            break
    if bad is None:
        print(f"stop={stop}: identical", flush=True)
        continue
    rows = np.nonzero(bad.any(axis=1))[0]
    cols = np.nonzero(bad.any(axis=0))[0]
    j0 = (stop - 1) * 128
    print(f"stop={stop} (last j0={j0}, c1={j0 + 128}, c2={j0 + 256}): {bad.sum()} differing elements, rows {rows.min()}..{rows.max()} "
          f"({rows.size}), cols {cols.min()}..{cols.max()} ({cols.size}); col histogram by 64: "
          f"{np.bincount(cols // 64)[cols.min() // 64:][:12].tolist()}", flush=True)
    # classify: is the difference +P (update missing), -P (applied twice) or something else?  P = L21 * U12 of the last step
    ref = res[1]
    for t0 in sorted(set((cols // 64).tolist()))[:4]:
        cs = slice(t0 * 64, t0 * 64 + 64)
        rws = np.nonzero(bad[:, cs].any(axis=1))[0]
        rb = sorted(set((rws // 128).tolist()))
        for b in rb[:3]:
            rs = slice(b * 128, b * 128 + 128)
            rs2 = slice(max(b * 128, j0 + 128), b * 128 + 128)
            P = ref[rs2, j0:j0 + 128] @ ref[j0:j0 + 128, cs]
            diff = res[0][rs2, cs] - ref[rs2, cs]
            nz = np.abs(diff) > 0
            print(f"   tile rows {rs2.start}..{rs2.stop} cols {cs.start}..{cs.stop}: differing {nz.sum()} of {diff.size}; "
                  f"|diff-P|max={np.abs(diff - P).max():.3e} |diff+P|max={np.abs(diff + P).max():.3e} |diff|max={np.abs(diff).max():.3e} "
                  f"rows-with-diff {np.nonzero(nz.any(axis=1))[0][[0, -1]].tolist()} cols-with-diff {np.nonzero(nz.any(axis=0))[0][[0, -1]].tolist()}")
    # bitmap of the first bad tile
    t0 = sorted(set((cols // 64).tolist()))[0]
    cs = slice(t0 * 64, t0 * 64 + 64)
    rws = np.nonzero(bad[:, cs].any(axis=1))[0]
    b = rws[0] // 128
    print(f"   bitmap of tile rows {b * 128}.. cols {cs.start}.. ('#' = differs)")
    for r in range(b * 128, b * 128 + 128):
        print("   %4d " % r + "".join("#" if v else "." for v in bad[r, cs]))
    # details of first few
    rr, cc = np.nonzero(bad)
    for k in range(min(5, rr.size)):
        print("   ", rr[k], cc[k], res[0][rr[k], cc[k]], res[1][rr[k], cc[k]])
    break
