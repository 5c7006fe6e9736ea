"""BASELINE config 2 at full size: f64 LU with partial pivoting, n = 16384, + solve with 16 RHS, on device-generated
seeded input.  The oracle's result for this exact input (order-preserving parallel form, ~8 minutes on 8 cores) is
committed as tests/golden/lu16384_f64.npz (see tests/golden/make_lu16384_fixture.py): the pivot permutation must be
IDENTICAL, sampled U-diagonal / packed rows must match to 1e-12*n, the backward error must be within 10x of the oracle's.
Size-independent properties close the loop: P*A == L*U through the CUDA GEMM itself, and A*x == b for the solve."""
import ctypes
import os

import numpy as np
import pytest

from la._cabi import check, lib

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "lu16384_f64.npz")


def test_lu_16384_matches_committed_oracle_fixture():
    torch = pytest.importorskip("torch")
    fx = np.load(FIX)
    n = int(fx["n"])
    L = lib()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream()
    sp = ctypes.c_void_p(st.cuda_stream)
    f64 = torch.float64
    A0 = torch.empty((n, n), dtype=f64, device=dev)
    check(L.la_fill_hash_f64_dev(A0.data_ptr(), A0.numel(), int(fx["seed"]), 0, sp))
    LU = A0.clone()
    piv = torch.empty((n,), dtype=torch.int64, device=dev)
    sign = torch.empty((1,), dtype=torch.int32, device=dev)
    check(L.la_lu_factor_f64_dev(LU.data_ptr(), n, n, piv.data_ptr(), sign.data_ptr(), sp))
    torch.cuda.synchronize()
    pv = piv.cpu().numpy()
    ref_piv = fx["piv"].astype(np.int64)
    mism = np.nonzero(pv != ref_piv)[0]
    assert mism.size == 0, f"pivot permutation differs from the reference's at {mism[:5]} ({mism.size} rows)"
    assert bool(sign.item()) == bool(fx["pospivsign"])
    # sampled values of the packed factors
    diag = torch.diagonal(LU)[::8].cpu().numpy()
    ref_diag = fx["diag_every8"]
    assert np.max(np.abs(diag - ref_diag) / np.maximum(np.abs(ref_diag), 1.0)) <= 1e-12 * n
    rows = fx["rows"]
    got = LU[torch.as_tensor(rows, device=dev)][:, ::16].cpu().numpy()
    ref = fx["row_samples_every16"]
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-12 * n
    # backward error through the CUDA GEMM: || A(piv,:) - L*U ||_F / ||A||_F   (torch only unpacks and takes norms)
    Lm = torch.tril(LU, -1)
    Lm.diagonal().fill_(1.0)
    Um = torch.triu(LU)
    del LU
    P = torch.empty((n, n), dtype=f64, device=dev)
    check(L.la_gemm_f64_dev(Lm.data_ptr(), n, Um.data_ptr(), n, P.data_ptr(), n, n, n, n, 0, sp))
    torch.cuda.synchronize()
    del Lm, Um
    PA = A0[piv]
    be = float((PA - P).norm() / A0.norm())
    assert be <= 10 * float(fx["backward_error"]), f"backward error {be} vs reference {float(fx['backward_error'])}"


def test_lu_16384_solve_16_rhs_residual():
    torch = pytest.importorskip("torch")
    n, nx = 16384, 16
    L = lib()
    dev = torch.device("cuda", 0)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    f64 = torch.float64
    A0 = torch.empty((n, n), dtype=f64, device=dev)
    B = torch.empty((n, nx), dtype=f64, device=dev)
    check(L.la_fill_hash_f64_dev(A0.data_ptr(), A0.numel(), 1, 0, sp))
    check(L.la_fill_hash_f64_dev(B.data_ptr(), B.numel(), 3, 0, sp))
    LU = A0.clone()
    piv = torch.empty((n,), dtype=torch.int64, device=dev)
    sign = torch.empty((1,), dtype=torch.int32, device=dev)
    X = torch.empty((n, nx), dtype=f64, device=dev)
    check(L.la_lu_factor_f64_dev(LU.data_ptr(), n, n, piv.data_ptr(), sign.data_ptr(), sp))
    check(L.la_lu_solve_f64_dev(LU.data_ptr(), n, piv.data_ptr(), B.data_ptr(), nx, X.data_ptr(), sp))
    R = torch.empty((n, nx), dtype=f64, device=dev)
    check(L.la_gemm_f64_dev(A0.data_ptr(), n, X.data_ptr(), nx, R.data_ptr(), nx, n, n, nx, 0, sp))
    torch.cuda.synchronize()
    res = float((R - B).norm() / (A0.norm() * X.norm()))
    assert res <= 1e-14, res


@pytest.mark.parametrize("n,nx", [(19200, 5), (19454, 16)])
def test_solve_with_more_block_rows_than_sms(n, nx):
    """n > 148 * 128 = 18944: the sweep kernels run one CTA per 128-row block and per SM, so larger systems are solved
    in leading parts with a GEMM update of the remaining right-hand sides in between (lu_solve.cu).  Property check:
    A x == b through the CUDA GEMM, and agreement with the many-right-hand-side path (a different code path, GEMM sweeps)
    on the same factors."""
    torch = pytest.importorskip("torch")
    L = lib()
    dev = torch.device("cuda", 0)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    f64 = torch.float64
    A0 = torch.empty((n, n), dtype=f64, device=dev)
    B = torch.empty((n, nx), dtype=f64, device=dev)
    check(L.la_fill_hash_f64_dev(A0.data_ptr(), A0.numel(), 1, 0, sp))
    check(L.la_fill_hash_f64_dev(B.data_ptr(), B.numel(), 3, 0, sp))
    LU = A0.clone()
    piv = torch.empty((n,), dtype=torch.int64, device=dev)
    sign = torch.empty((1,), dtype=torch.int32, device=dev)
    X = torch.full((n, nx), float("nan"), dtype=f64, device=dev)
    check(L.la_lu_factor_f64_dev(LU.data_ptr(), n, n, piv.data_ptr(), sign.data_ptr(), sp))
    check(L.la_lu_solve_f64_dev(LU.data_ptr(), n, piv.data_ptr(), B.data_ptr(), nx, X.data_ptr(), sp))
    R = torch.empty((n, nx), dtype=f64, device=dev)
    check(L.la_gemm_f64_dev(A0.data_ptr(), n, X.data_ptr(), nx, R.data_ptr(), nx, n, n, nx, 0, sp))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(X).all())
    res = float((R - B).norm() / (A0.norm() * X.norm()))
    assert res <= 1e-14, res
    # the same system with 18 right-hand sides (the first nx are B) takes the GEMM-sweep path
    B2 = torch.zeros((n, 18), dtype=f64, device=dev)
    B2[:, :nx] = B
    X2 = torch.empty((n, 18), dtype=f64, device=dev)
    check(L.la_lu_solve_f64_dev(LU.data_ptr(), n, piv.data_ptr(), B2.data_ptr(), 18, X2.data_ptr(), sp))
    torch.cuda.synchronize()
    diff = float((X2[:, :nx] - X).abs().max() / X.abs().max())
    assert diff <= 1e-9, diff
