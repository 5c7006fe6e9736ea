"""Regression for a timing-dependent corruption found while building the LU look-ahead: an earlier GEMM main loop
(fragments consumed right after their LDS) produced wrong tiles whenever unrelated work (another stream's kernels or
plain device-to-device copies) loaded the memory system.  The kernel must be bit-identical with and without neighbours."""
import ctypes

import numpy as np
import pytest

from la._cabi import check, lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", [3, 4])   # BN = 64 (2 CTA/SM) and BN = 128 (1 CTA/SM) tile configurations
@pytest.mark.parametrize("mode", [0, 1])
def test_gemm_is_bit_stable_under_concurrent_traffic(path, mode):
    torch = pytest.importorskip("torch")
    L = lib()
    dev = torch.device("cuda", 0)
    m, k, n = 2432, 128, 2304
    f64 = torch.float64
    g = torch.Generator(device=dev).manual_seed(5)
    A = torch.rand((m, k), dtype=f64, device=dev, generator=g)
    B = torch.rand((k, n), dtype=f64, device=dev, generator=g)
    C0 = torch.rand((m, n), dtype=f64, device=dev, generator=g)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    p1 = ctypes.c_void_p(s1.cuda_stream)
    check(L.la_debug_set_gemm_path(path))
    try:
        ref = C0.clone()
        torch.cuda.synchronize()
        check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, ref.data_ptr(), n, m, k, n, mode, p1))
        torch.cuda.synchronize()
        outs = [C0.clone() for _ in range(60)]
        src = torch.rand(1 << 22, dtype=f64, device=dev)
        dst = torch.empty_like(src)
        torch.cuda.synchronize()
        with torch.cuda.stream(s2):
            for _ in range(400):
                dst.copy_(src)
        for C in outs:
            check(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, k, n, mode, p1))
        torch.cuda.synchronize()
    finally:
        L.la_debug_set_gemm_path(0)
    bad = [i for i, C in enumerate(outs) if not torch.equal(C, ref)]
    assert not bad, f"{len(bad)} of {len(outs)} overlapped GEMMs differ from the quiet run: {bad[:5]}"


def test_lu_lookahead_is_deterministic_and_correct(oracle):
    """n = 3072: 24 panels, look-ahead pipeline on two streams; pivots identical to the oracle, run-to-run bit-stable."""
    n = 3072
    a = oracle.fill((n, n), 1)
    ref_lu, ref_piv, ref_sign = oracle.lu(a)
    prev = None
    for rep in range(3):
        lu = np.empty_like(a)
        piv = np.empty(n, dtype=np.uint64)
        sign = ctypes.c_int(0)
        check(lib().la_lu_factor_f64_host(a.ctypes.data, lu.ctypes.data, n, n, piv.ctypes.data, ctypes.byref(sign)))
        assert np.array_equal(piv, ref_piv), f"rep {rep}: first pivot mismatch at {np.nonzero(piv != ref_piv)[0][:3]}"
        assert bool(sign.value) == ref_sign
        err = np.max(np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), 1.0))
        assert err <= 1e-12 * n
        if prev is not None:
            assert np.array_equal(prev, lu)
        prev = lu
