#!/usr/bin/env python3
"""Times la_gemm_f64_dev on device-resident seeded inputs for a list of shapes / kernel paths (tuning aid).
usage: gemm_bench.py m,k,n[,mode[,path]] ...   (path: 0 auto, 3 BN=64, 4 BN=128)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import torch  # noqa: E402
from la._cabi import check, lib  # noqa: E402

L = lib()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
sp = ctypes.c_void_p(st.cuda_stream)
for spec in sys.argv[1:]:
    parts = [int(x) for x in spec.split(",")]
    m, k, n = parts[:3]
    mode = parts[3] if len(parts) > 3 else 0
    path = parts[4] if len(parts) > 4 else 0
    A = torch.empty((m, k), dtype=torch.float64, device=dev)
    B = torch.empty((k, n), dtype=torch.float64, device=dev)
    C = torch.zeros((m, n), dtype=torch.float64, device=dev)
    check(L.la_fill_hash_f64_dev(A.data_ptr(), A.numel(), 1, 0, sp))
    check(L.la_fill_hash_f64_dev(B.data_ptr(), B.numel(), 2, 0, sp))
    check(L.la_debug_set_gemm_path(path))
    reps = max(3, min(50, int(2e12 / (2.0 * m * n * k))))
    for _ in range(3):
        check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, mode, sp))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, mode, sp))
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"m={m} k={k} n={n} mode={mode} path={path}: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.2f} TFLOP/s", flush=True)
    check(L.la_debug_set_gemm_path(0))
