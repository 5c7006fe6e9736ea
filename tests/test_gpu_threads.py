"""The reference's Matrix<T>/LUDecomposition<T> are Send + Sync (plain Vecs), so concurrent `&a * &b` and factorisations
from many host threads are legal Rust ("Threading").  The C ABI promises per-thread streams, scratch
and error text: hammer it from several Python threads at once (ctypes drops the GIL during the calls) and check every
result against the single-threaded one."""
import threading

import numpy as np
import pytest

from la._cabi import check, lib

pytestmark = pytest.mark.gpu


def _gemm(a, b):
    m, k = a.shape
    n = b.shape[1]
    c = np.empty((m, n), dtype=a.dtype)
    fn = lib().la_gemm_f64_host if a.dtype == np.float64 else lib().la_gemm_f32_host
    check(fn(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
    return c


def _lu_solve(a, b):
    n = a.shape[0]
    nx = b.shape[1]
    lu = np.empty_like(a)
    piv = np.empty(n, dtype=np.uint64)
    import ctypes
    sign = ctypes.c_int(0)
    check(lib().la_lu_factor_f64_host(a.ctypes.data, lu.ctypes.data, n, n, piv.ctypes.data, ctypes.byref(sign)))
    x = np.empty((n, nx))
    check(lib().la_lu_solve_f64_host(lu.ctypes.data, n, n, piv.ctypes.data, b.ctypes.data, nx, x.ctypes.data))
    return lu, piv, x


def test_concurrent_calls_from_host_threads(oracle):
    jobs = []
    # mixed sizes: exact SIMT path, DMMA path, two-phase host pipeline, multi-panel LU + sweep solve, f32 tensor path
    for i, (m, k, n) in enumerate([(64, 64, 64), (700, 300, 500), (4224, 2304, 4100), (1024, 512, 768)]):
        jobs.append(("gemm", oracle.fill((m, k), 10 + i), oracle.fill((k, n), 20 + i)))
    jobs.append(("gemm", oracle.fill((1024, 256, ), 30, np.float32).reshape(1024, 256),
                 oracle.fill((256, 512), 31, np.float32)))
    for i, (n, nx) in enumerate([(96, 3), (700, 16), (1536, 8)]):
        jobs.append(("lu", oracle.fill((n, n), 40 + i), oracle.fill((n, nx), 50 + i)))

    def run(job):
        kind, a, b = job
        return _gemm(a, b) if kind == "gemm" else _lu_solve(a, b)

    expected = [run(j) for j in jobs]  # single-threaded first
    results = {}
    errors = []

    def worker(tid):
        try:
            for rep in range(3):
                for ji in range(len(jobs)):
                    j = (ji + tid) % len(jobs)  # different threads work on different jobs at the same moment
                    results[(tid, rep, j)] = run(jobs[j])
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (tid, rep, j), got in results.items():
        want = expected[j]
        if isinstance(want, tuple):
            for g, w in zip(got, want):
                assert np.array_equal(g, w), f"thread {tid} rep {rep} job {j}: LU/solve differs from the single-threaded run"
        else:
            assert np.array_equal(got, want), f"thread {tid} rep {rep} job {j}: product differs from the single-threaded run"
