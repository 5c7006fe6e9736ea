#!/usr/bin/env python3
"""Times la_chol_factor_f64 (+ solve with 16 right-hand sides) on a device-resident SPD matrix A = G G' + n I.
The input is built with torch (not part of the product path).  usage: chol_profile.py [n] [reps]"""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from gpu_util import DevBuf, sync  # noqa: E402
from la._cabi import check, lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = lib()
dev = torch.device("cuda", 0)
g = torch.rand((n, n), dtype=torch.float64, device=dev)
a0 = g @ g.T
a0 = (a0 + a0.T) * 0.5 + n * torch.eye(n, dtype=torch.float64, device=dev)
del g
a = DevBuf(n * n * 8)
b = DevBuf(n * 16 * 8)
x = DevBuf(n * 16 * 8)
rhs = torch.rand((n, 16), dtype=torch.float64, device=dev)
rhs_host = rhs.cpu().numpy()  # keep the array alive while ctypes reads it
check(L.la_buf_upload(b.h, 0, ctypes.c_void_p(rhs_host.ctypes.data), n * 16 * 8))
ok = ctypes.c_int(0)
host = a0.cpu().numpy()
for r in range(reps):
    check(L.la_buf_upload(a.h, 0, ctypes.c_void_p(host.ctypes.data), n * n * 8))
    sync()
    t0 = time.perf_counter()
    check(L.la_chol_factor_f64(a.h, n, ctypes.byref(ok)))
    t1 = time.perf_counter()
    check(L.la_chol_solve_f64(a.h, n, b.h, 16, x.h))
    t2 = time.perf_counter()
    print(f"n={n} cholesky ok={ok.value} {1e3 * (t1 - t0):.2f} ms ({n ** 3 / 3 / (t1 - t0) / 1e12:.2f} TFLOP/s)  "
          f"solve(16 rhs) {1e3 * (t2 - t1):.2f} ms", flush=True)
lt = torch.from_numpy(DevBuf.to_array(a, (n, n), "float64")).to(dev)
xs = torch.from_numpy(DevBuf.to_array(x, (n, 16), "float64")).to(dev)
print("||L L' - A|| / ||A|| =", float((lt @ lt.T - a0).norm() / a0.norm()),
      " solve residual =", float((a0 @ xs - rhs).norm() / (a0.norm() * xs.norm())))
