#!/usr/bin/env python3
"""Times la_gemm_f32_dev (tcgen05 TF32 or CUDA-core path) on device-resident seeded inputs.  usage: m,k,n[,path] ..."""
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
    path = parts[3] if len(parts) > 3 else 0
    A = torch.empty((m, k), dtype=torch.float32, device=dev)
    B = torch.empty((k, n), dtype=torch.float32, device=dev)
    C = torch.zeros((m, n), dtype=torch.float32, device=dev)
    check(L.la_fill_hash_f32_dev(A.data_ptr(), A.numel(), 1, 0, sp))
    check(L.la_fill_hash_f32_dev(B.data_ptr(), B.numel(), 2, 0, sp))
    check(L.la_debug_set_gemm_f32_path(path))
    reps = max(3, min(30, int(4e12 / (2.0 * m * n * k)))) if path != 1 else 2
    for _ in range(2):
        check(L.la_gemm_f32_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, 0, sp))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps):
        check(L.la_gemm_f32_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, 0, sp))
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # spot check against torch fp32 (cuBLAS) on a slice: harness only
    ref = (A[:256].double() @ B.double()).float()
    err = float(((C[:256] - ref).abs() / ref.abs().clamp_min(1e-30)).max())
    print(f"m={m} k={k} n={n} path={path}: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s  max rel err vs fp64 ref {err:.2e}", flush=True)
    check(L.la_debug_set_gemm_f32_path(0))
