#!/usr/bin/env python3
"""Stress 2: C -= A*B (K=128, BN=64 or 128) on stream 1 while stream 2 runs simple torch kernels of a chosen flavour:
  int  : int32 elementwise adds (no FP64 pipe, no shared memory)
  f64  : float64 elementwise fused multiply-adds (FP64 pipe, no shared memory)
  f32  : float32 elementwise
  copy : device-to-device copies
Reports how many overlapped GEMMs differ from the overlap-free result."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import torch  # noqa: E402
from la._cabi import check, lib  # noqa: E402

L = lib()
dev = torch.device("cuda", 0)
path = int(sys.argv[1])
flavour = sys.argv[2]
small = len(sys.argv) > 3 and sys.argv[3] == "small"
m, k, n = 2432, 128, 2304
f64 = torch.float64
A = torch.rand((m, k), dtype=f64, device=dev)
B = torch.rand((k, n), dtype=f64, device=dev)
C0 = torch.rand((m, n), dtype=f64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
p1 = ctypes.c_void_p(s1.cuda_stream)
check(L.la_debug_set_gemm_path(path))
ref = C0.clone()
torch.cuda.synchronize()
check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, ref.data_ptr(), n, m, k, n, 1, p1))
torch.cuda.synchronize()
NG = 100
outs = [C0.clone() for _ in range(NG)]
cnt = (1 << 14) if small else (1 << 22)
xi = torch.zeros(cnt, dtype=torch.int32, device=dev)
xd = torch.rand(cnt, dtype=f64, device=dev)
xf = torch.rand(cnt, dtype=torch.float32, device=dev)
yd = torch.empty_like(xd)
torch.cuda.synchronize()
with torch.cuda.stream(s2):
    for _ in range(3000 if small else 600):
        if flavour == "int":
            xi.add_(1)
        elif flavour == "f64":
            torch.addcmul(xd, xd, xd, value=0.5, out=yd)
        elif flavour == "f32":
            xf.mul_(1.0001)
        elif flavour == "copy":
            yd.copy_(xd)
for C in outs:
    check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, 1, p1))
torch.cuda.synchronize()
bad = sum(1 for C in outs if bool((C != ref).any()))
print(f"path={path} foreign={flavour}{' small' if small else ''}: {bad} of {NG} overlapped GEMMs differ", flush=True)
