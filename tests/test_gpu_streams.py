"""Two asynchronous `*_dev` calls from ONE host thread on DIFFERENT streams (round-1 advisor finding): the library's
scratch (the fp32 B^T / split buffers, the LU exchange workspace, W, the side streams) is per (thread, device), so the
second call must wait for the first one's tail instead of overwriting what it still reads."""
import ctypes

import numpy as np
import pytest

from gpu_util import max_rel_err
from la._cabi import check, lib

pytestmark = pytest.mark.gpu


def test_f32_gemms_on_two_streams_share_scratch_safely(oracle):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    m, k, n = 4096, 512, 4096
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a = torch.from_numpy(oracle.fill((m, k), 1, np.float32)).to(dev)
    b1 = torch.from_numpy(oracle.fill((k, n), 2, np.float32)).to(dev)
    b2 = torch.from_numpy(oracle.fill((k, n), 7, np.float32)).to(dev)
    c1 = torch.empty((m, n), dtype=torch.float32, device=dev)
    c2 = torch.empty((m, n), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    L = lib()
    for _ in range(3):
        check(L.la_gemm_f32_dev(a.data_ptr(), k, b1.data_ptr(), n, c1.data_ptr(), n, m, k, n, 0, ctypes.c_void_p(s1.cuda_stream)))
        check(L.la_gemm_f32_dev(a.data_ptr(), k, b2.data_ptr(), n, c2.data_ptr(), n, m, k, n, 0, ctypes.c_void_p(s2.cuda_stream)))
    torch.cuda.synchronize()
    rows = [0, 1, 2047, 4095]
    an = a.cpu().numpy()
    for c, b in ((c1, b1), (c2, b2)):
        bn = b.cpu().numpy()
        cn = c.cpu().numpy()
        for r in rows:
            assert max_rel_err(cn[r:r + 1], oracle.gemm_rows(an, bn, r, r + 1)) <= 4e-6 + 1.2e-7 * k


def test_lu_factorisations_on_two_streams(oracle):
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    n = 1024
    mats = [oracle.fill((n, n), s) for s in (1, 5)]
    refs = [oracle.lu(a) for a in mats]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    lus = [torch.from_numpy(a.copy()).to(dev) for a in mats]
    pivs = [torch.empty((n,), dtype=torch.int64, device=dev) for _ in mats]
    signs = [torch.empty((1,), dtype=torch.int32, device=dev) for _ in mats]
    torch.cuda.synchronize()
    L = lib()
    for i in range(2):
        check(L.la_lu_factor_f64_dev(lus[i].data_ptr(), n, n, pivs[i].data_ptr(), signs[i].data_ptr(),
                                     ctypes.c_void_p(streams[i].cuda_stream)))
    torch.cuda.synchronize()
    for i in range(2):
        ref_lu, ref_piv, ref_sign = refs[i]
        assert np.array_equal(pivs[i].cpu().numpy().astype(np.uint64), ref_piv)
        assert bool(signs[i].item()) == ref_sign
        got = lus[i].cpu().numpy()
        assert float(np.max(np.abs(got - ref_lu) / np.maximum(np.abs(ref_lu), 1.0))) <= 1e-12 * n
