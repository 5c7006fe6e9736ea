"""Host-side plan and binding of the multi-GPU Mul (`la_gemm_*_mg*`, include/la_cabi.h).

Every row of C = A*B depends on the same row of A and on all of B (reference loop nest, src/matrix/mod.rs:965-973), so
GPU g owns a row block of A and C.  B is not broadcast from one owner: rank q owns a COLUMN block of B (the part it
uploads or produces) and every rank pulls the other blocks over NVLink while it already multiplies by the blocks it has;
C[:, block] = A * B[:, block] needs only that block, so every multiply runs at full depth K with a plain store.

`shard` is the partition all ranks must agree on (la_mg_shard: pure arithmetic inside the C library, no device needed);
`column_ranges` is the order in which a rank multiplies (own block, the blocks right of it, the blocks left of it);
`gather_blocks_and_multiply` runs that plan with any transport (NVLink pulls inside the library on the GPU, gloo in
tests/test_multi_rank_cpu.py); `MgContext` wraps the one-process-per-GPU C API."""
import ctypes

import numpy as np

from . import _cabi
from ._cabi import check, lib


def shard(nranks, rank, m, n, elem_bytes=8):
    """(row0, row1, col0, col1): rows of A / C and columns of B owned by `rank` (la_mg_shard)."""
    out = [ctypes.c_size_t() for _ in range(4)]
    check(lib().la_mg_shard(nranks, rank, m, n, elem_bytes, *[ctypes.byref(o) for o in out]))
    return tuple(int(o.value) for o in out)


def row_shard(m, world, rank):
    """Rows [r0, r1) of A and C owned by `rank` (128-row tile bands, earlier ranks take the remainder)."""
    r0, r1, _, _ = shard(world, rank, m, 256 * world)
    return r0, r1


def column_ranges(nranks, rank, n, elem_bytes=8):
    """[(col0, col1, owners)]: the rank's own block first, then everything right of it, then everything left of it --
    the order in which the library multiplies while the pulls of the later ranges are still in flight."""
    cols = [shard(nranks, q, 128, n, elem_bytes)[2:] for q in range(nranks)]
    own = cols[rank]
    out = [(own[0], own[1], [rank])]
    if own[1] < n:
        out.append((own[1], n, list(range(rank + 1, nranks))))
    if own[0] > 0:
        out.append((0, own[0], list(range(0, rank))))
    return [r for r in out if r[1] > r[0]]


def gather_blocks_and_multiply(a_shard, b_own, c_shard, nranks, rank, n, fetch_block, multiply):
    """One rank's side of the plan with an arbitrary transport.

    fetch_block(q) -> the column block of B owned by rank q (b_own for q == rank)
    multiply(col0, col1, b_cols): C_shard[:, col0:col1] = A_shard * b_cols
    """
    elem = a_shard.dtype.itemsize
    for col0, col1, owners in column_ranges(nranks, rank, n, elem):
        blocks = [b_own if q == rank else fetch_block(q) for q in owners]
        multiply(col0, col1, np.concatenate(blocks, axis=1) if len(blocks) > 1 else blocks[0])
    return c_shard


class MgContext:
    """One rank of the one-process-per-GPU multi-GPU Mul (la_mg_create / la_mg_handle / la_mg_connect)."""

    def __init__(self, rank, nranks, device, dtype, k, n):
        self.dtype = np.dtype(dtype)
        self.suf = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[self.dtype]
        self.rank, self.nranks, self.device, self.k, self.n = rank, nranks, device, k, n
        self.h = ctypes.c_void_p()
        check(lib().la_mg_create(rank, nranks, device, self.dtype.itemsize, k, n, ctypes.byref(self.h)))

    def handle(self):
        buf = (ctypes.c_ubyte * _cabi.LA_MG_HANDLE_BYTES)()
        check(lib().la_mg_handle(self.h, buf))
        return bytes(buf)

    def connect(self, handles):
        """handles: the nranks handles in rank order (bytes objects or one concatenated bytes)."""
        blob = handles if isinstance(handles, (bytes, bytearray)) else b"".join(handles)
        assert len(blob) == self.nranks * _cabi.LA_MG_HANDLE_BYTES
        check(lib().la_mg_connect(self.h, blob))

    def b_block(self):
        """(device pointer of replica[0][col0], ldb, col0, col1)"""
        p, ld, c0, c1 = ctypes.c_void_p(), ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_size_t()
        check(lib().la_mg_b_block(self.h, ctypes.byref(p), ctypes.byref(ld), ctypes.byref(c0), ctypes.byref(c1)))
        return int(p.value), int(ld.value), int(c0.value), int(c1.value)

    def gemm(self, a_ptr, lda, c_ptr, ldc, m_local, stream=None):
        check(getattr(lib(), f"la_gemm_{self.suf}_mg_rank")(self.h, a_ptr, lda, c_ptr, ldc, m_local, stream))

    def gemm_host(self, a_shard, b_block, c_shard):
        """numpy (or pinned) host shards; b_block is this rank's k x (col1-col0) column block, C-contiguous."""
        assert a_shard.flags.c_contiguous and b_block.flags.c_contiguous and c_shard.flags.c_contiguous
        check(getattr(lib(), f"la_gemm_{self.suf}_mg_rank_host")(self.h, a_shard.ctypes.data, b_block.ctypes.data,
                                                               b_block.shape[1], c_shard.ctypes.data, a_shard.shape[0]))

    def reserve(self, m_local):
        """Size the device copies of a host shard while no rank is inside a product (la_mg_reserve)."""
        check(lib().la_mg_reserve(self.h, m_local))

    def quiesce(self, stream=None):
        check(lib().la_mg_quiesce(self.h, stream))

    def destroy(self):
        if self.h:
            lib().la_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def gemm_mg(a, b, devices):
    """Single-process multi-GPU product of host matrices (la_gemm_*_mg): what `&a * &b` binds when several devices are given."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape[1] == b.shape[0]
    suf = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}[a.dtype]
    c = np.empty((a.shape[0], b.shape[1]), dtype=a.dtype)
    devs = (ctypes.c_int * len(devices))(*devices)
    check(getattr(lib(), f"la_gemm_{suf}_mg")(len(devices), devs, a.ctypes.data, b.ctypes.data, c.ctypes.data, a.shape[0],
                                              a.shape[1], b.shape[1]))
    return c


def lu_mg_plan(n, ngpus, sm_count=148):
    """(block width, number of block columns, devices in use, [local columns per device]) of the multi-device LU layout
    (la_lu_mg_plan: pure arithmetic inside the C library, no device needed)."""
    nb, nblk, ndev = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    ncols = (ctypes.c_size_t * ngpus)()
    check(lib().la_lu_mg_plan(n, ngpus, sm_count, ctypes.byref(nb), ctypes.byref(nblk), ctypes.byref(ndev), ncols))
    return nb.value, nblk.value, ndev.value, [int(x) for x in ncols]


class LuMgContext:
    """LU of one n x n fp64 matrix across several devices driven by this thread (la_lu_mg_*, include/la_cabi.h): 128-column
    blocks dealt round-robin, the owner's factored block column copied to every device, no collective."""

    def __init__(self, devices, n):
        self.n = int(n)
        self.h = ctypes.c_void_p()
        devs = (ctypes.c_int * len(devices))(*devices)
        check(lib().la_lu_mg_create(len(devices), devs, self.n, ctypes.byref(self.h)))

    def devices_in_use(self):
        out = ctypes.c_int()
        check(lib().la_lu_mg_devices(self.h, ctypes.byref(out)))
        return out.value

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == (self.n, self.n)
        check(lib().la_lu_mg_upload_f64(self.h, a.ctypes.data))

    def fill_hash(self, seed):
        check(lib().la_lu_mg_fill_hash_f64(self.h, seed))

    def factor(self):
        check(lib().la_lu_mg_factor_f64(self.h))

    def sync(self):
        check(lib().la_lu_mg_sync(self.h))

    def last_ms(self):
        out = ctypes.c_float()
        check(lib().la_lu_mg_last_ms(self.h, ctypes.byref(out)))
        return out.value

    def download(self, want_lu=True):
        lu = np.empty((self.n, self.n), dtype=np.float64) if want_lu else None
        piv = np.empty(self.n, dtype=np.uint64)
        sign = ctypes.c_int(-7)
        check(lib().la_lu_mg_download_f64(self.h, lu.ctypes.data if want_lu else None, piv.ctypes.data, ctypes.byref(sign)))
        return lu, piv, bool(sign.value)

    def destroy(self):
        if self.h:
            lib().la_lu_mg_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def lu_factor_mg(a, devices):
    """(packed LU, piv, pospivsign) of a host matrix across `devices` (la_lu_factor_f64_mg)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    n = a.shape[0]
    assert a.shape == (n, n)
    lu = np.empty_like(a)
    piv = np.empty(n, dtype=np.uint64)
    sign = ctypes.c_int(-7)
    devs = (ctypes.c_int * len(devices))(*devices)
    check(lib().la_lu_factor_f64_mg(len(devices), devs, a.ctypes.data, lu.ctypes.data, n, piv.ctypes.data, ctypes.byref(sign)))
    return lu, piv, bool(sign.value)
