"""Multi-GPU Mul behind the C ABI (la_gemm_*_mg*) against the oracle.

* two RANKS ON ONE DEVICE in one process (two host threads, two contexts): exercises the whole protocol -- replica,
  ready/ack flags, pull kernels, column-range GEMMs, host pipeline -- on a 1-GPU box;
* one process, two devices (la_gemm_*_mg) and one process per device over CUDA IPC handles (tests/mg_worker.py under
  torch.distributed/gloo): skipped below 2 GPUs.
Rows of a product are independent, so sampled full rows against oracle.gemm_rows are an exact check."""
import ctypes
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from gpu_util import DevBuf, fill_hash, max_rel_err, sync
from la import _cabi, sharding
from la._cabi import check, lib

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ran_in_child(request):
    """Ranks that share ONE device depend on each other's kernels making progress side by side; a scheduling hazard there
    would hang the whole pytest process.  Each such test therefore re-runs itself in a child process with a hard timeout:
    a hang fails this test only."""
    if os.environ.get("LA_MG_CHILD") == "1":
        return False
    env = dict(os.environ, LA_MG_CHILD="1", CUDA_MODULE_LOADING="EAGER")
    out = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", request.node.nodeid], cwd=ROOT,
                         env=env, capture_output=True, text=True, timeout=90)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    return True


def _pinned(shape, dtype):
    """numpy view of page-locked host memory (la_host_alloc).  Pageable buffers are staged by the driver, and a staged
    copy issued while another rank's kernel is spinning on the SAME device can block behind it."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    check(lib().la_host_alloc(n, ctypes.byref(p)))
    buf = (ctypes.c_ubyte * n).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    return arr, p


def _run_ranks(fn, world):
    errs = [None] * world

    def body(r):
        try:
            fn(r)
        except BaseException as e:  # noqa: BLE001
            errs[r] = e

    th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank hung"
    for e in errs:
        if e is not None:
            raise e


@pytest.mark.parametrize("dtype,shape,world", [(np.float64, (1024, 640, 1536), 2), (np.float64, (1280, 512, 1100), 3),
                                               (np.float32, (2048, 256, 2048), 2), (np.float64, (512, 2304, 768), 2)])
def test_mg_ranks_on_one_device_host_shards(request, oracle, dtype, shape, world):
    """la_gemm_*_mg_rank_host with every rank on device 0: three products in a row (the second and third overwrite the
    column blocks in the replicas, so the ack protocol is on the path), all rows compared with the oracle."""
    if _ran_in_child(request):
        return
    m, k, n = shape
    ctxs = [sharding.MgContext(r, world, 0, dtype, k, n) for r in range(world)]
    handles = [c.handle() for c in ctxs]
    for r, c in enumerate(ctxs):
        c.connect(handles)
        r0, r1, _, _ = sharding.shard(world, r, m, n, np.dtype(dtype).itemsize)
        c.reserve(r1 - r0)  # ranks share one device here: no allocation may happen while a peer's pull is spinning
    tol = 1e-12 * k if dtype == np.float64 else 4e-6 + 1.2e-7 * k
    try:
        for rep in range(3):
            a = oracle.fill((m, k), 1 + 10 * rep, dtype)
            b = oracle.fill((k, n), 2 + 10 * rep, dtype)
            outs = [None] * world

            held = []

            def rank_body(r):
                r0, r1, c0, c1 = sharding.shard(world, r, m, n, np.dtype(dtype).itemsize)
                (a_sh, pa), (b_bl, pb), (c_sh, pc) = (_pinned((r1 - r0, k), dtype), _pinned((k, c1 - c0), dtype),
                                                      _pinned((r1 - r0, n), dtype))
                held.extend([pa, pb, pc])
                a_sh[:] = a[r0:r1]
                b_bl[:] = b[:, c0:c1]
                c_sh[:] = np.nan
                ctxs[r].gemm_host(a_sh, b_bl, c_sh)
                outs[r] = c_sh.copy()

            _run_ranks(rank_body, world)
            for p in held:
                lib().la_host_free(p)
            got = np.concatenate(outs, axis=0)
            assert np.all(np.isfinite(got))
            assert max_rel_err(got, oracle.gemm(a, b)) <= tol
    finally:
        for c in ctxs:
            c.destroy()


def test_mg_ranks_on_one_device_resident_shards(request, oracle):
    """la_gemm_f64_mg_rank: shards resident in HBM, each rank's column block written into its replica by the device fill;
    two products back to back on each rank's stream without host synchronisation in between."""
    if _ran_in_child(request):
        return
    m, k, n, world = 1024, 768, 2048, 2
    ctxs = [sharding.MgContext(r, world, 0, np.float64, k, n) for r in range(world)]
    handles = [c.handle() for c in ctxs]
    for c in ctxs:
        c.connect(handles)
    a, b = oracle.fill((m, k), 1), oracle.fill((k, n), 2)
    outs = [None] * world
    try:
        def rank_body(r):
            r0, r1, c0, c1 = sharding.shard(world, r, m, n)
            da = DevBuf.from_array(a[r0:r1])
            dc = DevBuf((r1 - r0) * n * 8)
            ptr, ldb, bc0, bc1 = ctxs[r].b_block()
            assert (bc0, bc1, ldb) == (c0, c1, n)
            # the rank's column block, generated in place row by row: element (i, j) of B is hash(seed 2, i * n + j)
            for i in range(k):
                check(lib().la_fill_hash_f64_dev(ctypes.c_void_p(ptr + i * ldb * 8), c1 - c0, 2, i * n + c0, None))
            for _ in range(2):
                ctxs[r].gemm(da.ptr(), k, dc.ptr(), n, r1 - r0)
            sync()
            outs[r] = dc.to_array((r1 - r0, n), np.float64)

        _run_ranks(rank_body, world)
        got = np.concatenate(outs, axis=0)
        assert max_rel_err(got, oracle.gemm(a, b)) <= 1e-12 * k
    finally:
        for c in ctxs:
            c.destroy()


def _need_two_gpus():
    if _cabi.device_count() < 2:
        pytest.skip("needs 2 GPUs")


@pytest.mark.parametrize("dtype,shape", [(np.float64, (4096, 2304, 4608)), (np.float32, (8192, 1024, 4096)),
                                         (np.float64, (300, 100, 50))])
def test_mg_single_process_two_devices(oracle, dtype, shape):
    """la_gemm_*_mg on devices [0, 1]: host operands, sampled rows against the oracle plus the column-sum checksum."""
    _need_two_gpus()
    m, k, n = shape
    a, b = oracle.fill((m, k), 1, dtype), oracle.fill((k, n), 2, dtype)
    tol = 1e-12 * k if dtype == np.float64 else 4e-6 + 1.2e-7 * k
    for _ in range(2):  # the second call reuses the cached contexts
        c = sharding.gemm_mg(a, b, [0, 1])
        assert np.all(np.isfinite(c))
        rows = np.unique(np.concatenate([[0, m // 2 - 1, m // 2, m - 1], np.random.default_rng(1).integers(0, m, 60)]))
        for r in rows:
            assert max_rel_err(c[r:r + 1], oracle.gemm_rows(a, b, int(r), int(r) + 1)) <= tol
        lhs = a.astype(np.float64).sum(axis=0) @ b.astype(np.float64)
        assert np.max(np.abs(lhs - c.astype(np.float64).sum(axis=0)) / np.abs(lhs)) <= (1e-11 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype", ["f64", "f32"])
def test_mg_one_process_per_device_ipc(tmp_path, dtype):
    """Two processes, one device each, handles exchanged over gloo, device-resident and host-shard products checked
    against the oracle inside the workers (tests/mg_worker.py)."""
    _need_two_gpus()
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mg_worker.py"), dtype], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for r, p in enumerate(procs):
        assert p.returncode == 0, f"rank {r} failed:\n{outs[r][-3000:]}"
        assert "MG_WORKER_OK" in outs[r]
