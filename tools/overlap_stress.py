#!/usr/bin/env python3
"""Stress: repeated C -= A*B (K=128) on stream 1 while an unrelated LU (panel kernels, serial pipeline) runs on stream 2.
No data is shared; any difference from the overlap-free result is kernel-level interference."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import torch  # noqa: E402
from la._cabi import check, lib  # noqa: E402

L = lib()
dev = torch.device("cuda", 0)
path = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m, k, n = 2432, 128, 2304
f64 = torch.float64
A = torch.rand((m, k), dtype=f64, device=dev)
B = torch.rand((k, n), dtype=f64, device=dev)
C0 = torch.rand((m, n), dtype=f64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
p1, p2 = ctypes.c_void_p(s1.cuda_stream), ctypes.c_void_p(s2.cuda_stream)
check(L.la_debug_set_gemm_path(path))

def gemm(C):
    check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, mode, p1))

ref = C0.clone()
torch.cuda.synchronize()
gemm(ref)
torch.cuda.synchronize()
N2 = 3072
M2 = torch.rand((N2, N2), dtype=f64, device=dev)
piv = torch.empty((N2 + 8,), dtype=torch.int64, device=dev)
outs = [C0.clone() for _ in range(int(os.environ.get("STRESS_N", "40")))]
torch.cuda.synchronize()
check(L.la_lu_factor_f64_dev(M2.data_ptr(), N2, N2, piv.data_ptr(), piv.data_ptr() + N2 * 8, p2))
for C in outs:
    gemm(C)
torch.cuda.synchronize()
bad = 0
for i, C in enumerate(outs):
    d = (C != ref)
    if bool(d.any()):
        bad += 1
        idx = d.nonzero()
        print(f"gemm #{i}: {int(d.sum())} differing elements; rows {int(idx[:,0].min())}..{int(idx[:,0].max())} cols {int(idx[:,1].min())}..{int(idx[:,1].max())}")
print(f"path={path} mode={mode} extra={os.environ.get('LA_GEMM_EXTRA_SMEM')} LA_LU_DEBUG={os.environ.get('LA_LU_DEBUG')}: {bad} of {len(outs)} overlapped GEMMs differ from the overlap-free result")
