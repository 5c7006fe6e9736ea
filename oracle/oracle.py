"""ctypes binding of oracle/liblaoracle.so (TEST INFRASTRUCTURE ONLY).

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (rust-la_b200/) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblaoracle.so")


def build(force=False):
    """Compile the C restatement (gcc, -ffp-contract=off).  Building the checker is not using it."""
    src = os.path.join(_HERE, "la_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _declare(_lib)
    return _lib


_sz = ctypes.c_size_t
_u64 = ctypes.c_uint64
_p = ctypes.c_void_p
_int = ctypes.c_int


def _declare(L):
    L.oracle_num_threads.restype = _int
    for suf in ("f64", "f32"):
        getattr(L, f"oracle_fill_{suf}").argtypes = [_p, _sz, _u64, _u64]
        for form in ("canon", "fast"):
            getattr(L, f"oracle_gemm_{form}_{suf}").argtypes = [_p, _p, _p, _sz, _sz, _sz]
            getattr(L, f"oracle_lu_{form}_{suf}").argtypes = [_p, _sz, _sz, _p, _p]
        getattr(L, f"oracle_gemm_canon_rows_{suf}").argtypes = [_p, _p, _p, _sz, _sz, _sz, _sz, _sz, _int]
        getattr(L, f"oracle_gemm_fast_rows_{suf}").argtypes = [_p, _p, _p, _sz, _sz, _sz, _sz, _sz]
        f = getattr(L, f"oracle_lu_is_non_singular_{suf}")
        f.argtypes, f.restype = [_p, _sz], _int
        f = getattr(L, f"oracle_lu_det_{suf}")
        f.argtypes, f.restype = [_p, _sz, _int], (ctypes.c_double if suf == "f64" else ctypes.c_float)
        for form in ("", "_fast"):
            f = getattr(L, f"oracle_lu_solve{form}_{suf}")
            f.argtypes, f.restype = [_p, _sz, _sz, _p, _p, _sz, _p], _int
        getattr(L, f"oracle_lu_get_l_{suf}").argtypes = [_p, _sz, _sz, _p]
        getattr(L, f"oracle_lu_get_u_{suf}").argtypes = [_p, _sz, _sz, _p]
        getattr(L, f"oracle_identity_{suf}").argtypes = [_p, _sz]
        for form in ("canon", "fast"):
            f = getattr(L, f"oracle_chol_{form}_{suf}")
            f.argtypes, f.restype = [_p, _sz, _p], _int
        getattr(L, f"oracle_chol_solve_{suf}").argtypes = [_p, _sz, _p, _sz, _p]
        for form in ("canon", "fast"):
            getattr(L, f"oracle_qr_{form}_{suf}").argtypes = [_p, _sz, _sz, _p, _p]
        getattr(L, f"oracle_qr_get_q_{suf}").argtypes = [_p, _p, _sz, _sz, _p]
        getattr(L, f"oracle_qr_solve_{suf}").argtypes = [_p, _p, _sz, _sz, _p, _sz, _p]
        f = getattr(L, f"oracle_lu_backward_error_{suf}")
        f.argtypes, f.restype = [_p, _p, _sz, _sz, _p], ctypes.c_double
    L.oracle_gemm_canon_i64.argtypes = [_p, _p, _p, _sz, _sz, _sz]


def _suf(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"oracle supports f32/f64, got {dtype}")


def _ptr(a):
    return a.ctypes.data_as(_p)


def num_threads():
    return lib().oracle_num_threads()


def fill(shape, seed, dtype=np.float64, first_idx=0):
    """Synthetic uniform [0,1) matrix; element idx is a pure function of (seed, idx)."""
    out = np.empty(shape, dtype=dtype)
    getattr(lib(), f"oracle_fill_{_suf(dtype)}")(_ptr(out), out.size, seed, first_idx)
    return out


def fill_block(row0, rows, col0, cols, ld, seed, dtype=np.float64):
    """Rows [row0, row0+rows) x columns [col0, col0+cols) of the synthetic matrix with row length `ld` (element (i, j) is
    hash(seed, i * ld + j)): lets a checker rebuild a window of a 32768 x 32768 operand without the 8 GiB whole."""
    out = np.empty((rows, cols), dtype=dtype)
    f = getattr(lib(), f"oracle_fill_{_suf(dtype)}")
    isz = out.itemsize
    base = out.ctypes.data
    for i in range(rows):
        f(_p(base + i * cols * isz), cols, seed, (row0 + i) * ld + col0)
    return out


def gemm(a, b, form="fast"):
    """C = A*B in the reference's per-element order (src/matrix/mod.rs:957-980)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.ndim == 2 and b.ndim == 2 and a.shape[1] == b.shape[0] and a.dtype == b.dtype
    if a.dtype == np.int64:
        c = np.empty((a.shape[0], b.shape[1]), dtype=np.int64)
        lib().oracle_gemm_canon_i64(_ptr(a), _ptr(b), _ptr(c), a.shape[0], a.shape[1], b.shape[1])
        return c
    c = np.empty((a.shape[0], b.shape[1]), dtype=a.dtype)
    getattr(lib(), f"oracle_gemm_{form}_{_suf(a.dtype)}")(_ptr(a), _ptr(b), _ptr(c), a.shape[0], a.shape[1], b.shape[1])
    return c


def gemm_rows(a, b, row0, row1, form="fast", threads=1):
    """Rows [row0,row1) of A*B (rows are independent, so a sampled check is exact)."""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    c = np.empty((row1 - row0, b.shape[1]), dtype=a.dtype)
    suf = _suf(a.dtype)
    if form == "canon":
        getattr(lib(), f"oracle_gemm_canon_rows_{suf}")(_ptr(a), _ptr(b), _ptr(c), a.shape[0], a.shape[1], b.shape[1],
                                                        row0, row1, threads)
    else:
        getattr(lib(), f"oracle_gemm_fast_rows_{suf}")(_ptr(a), _ptr(b), _ptr(c), a.shape[0], a.shape[1], b.shape[1],
                                                       row0, row1)
    return c


def lu(a, form="fast"):
    """LUDecomposition::new (src/decomp/lu.rs:104-168). Returns (packed lu, piv[u64], pospivsign)."""
    a = np.ascontiguousarray(a)
    m, n = a.shape
    packed = a.copy()
    piv = np.empty(m, dtype=np.uint64)
    pos = _int(1)
    getattr(lib(), f"oracle_lu_{form}_{_suf(a.dtype)}")(_ptr(packed), m, n, _ptr(piv), ctypes.byref(pos))
    return packed, piv, bool(pos.value)


def lu_is_non_singular(packed):
    n = packed.shape[1]
    assert packed.shape[0] >= n, "reference would index out of bounds for m < n (lu.rs:174-182)"
    return bool(getattr(lib(), f"oracle_lu_is_non_singular_{_suf(packed.dtype)}")(_ptr(packed), n))


def lu_det(packed, pospivsign):
    assert packed.shape[0] == packed.shape[1]
    return packed.dtype.type(getattr(lib(), f"oracle_lu_det_{_suf(packed.dtype)}")(_ptr(packed), packed.shape[0],
                                                                                   int(pospivsign)))


def lu_solve(packed, piv, b, form="fast"):
    """LUDecomposition::solve (src/decomp/lu.rs:237-278). None when singular."""
    b = np.ascontiguousarray(b)
    m, n = packed.shape
    assert b.shape[0] == m
    x = np.empty((m, b.shape[1]), dtype=packed.dtype)
    name = f"oracle_lu_solve{'_fast' if form == 'fast' else ''}_{_suf(packed.dtype)}"
    ok = getattr(lib(), name)(_ptr(packed), m, n, _ptr(np.ascontiguousarray(piv, dtype=np.uint64)), _ptr(b),
                              b.shape[1], _ptr(x))
    return x if ok else None


def lu_get_l(packed):
    m, n = packed.shape
    l = np.empty((m, min(m, n)), dtype=packed.dtype)
    getattr(lib(), f"oracle_lu_get_l_{_suf(packed.dtype)}")(_ptr(packed), m, n, _ptr(l))
    return l


def lu_get_u(packed):
    m, n = packed.shape
    u = np.empty((min(m, n), n), dtype=packed.dtype)
    getattr(lib(), f"oracle_lu_get_u_{_suf(packed.dtype)}")(_ptr(packed), m, n, _ptr(u))
    return u


def identity(n, dtype=np.float64):
    d = np.empty((n, n), dtype=dtype)
    getattr(lib(), f"oracle_identity_{_suf(dtype)}")(_ptr(d), n)
    return d


def lu_backward_error(a, packed, piv):
    """||A(piv,:) - L*U||_F / ||A||_F, accumulated in extended precision."""
    a = np.ascontiguousarray(a)
    packed = np.ascontiguousarray(packed)
    m, n = a.shape
    return getattr(lib(), f"oracle_lu_backward_error_{_suf(a.dtype)}")(
        _ptr(a), _ptr(packed), m, n, _ptr(np.ascontiguousarray(piv, dtype=np.uint64)))


def chol(a, form="fast"):
    """CholeskyDecomposition::new (src/decomp/cholesky.rs:56-110): L (lower, zeros above) or None."""
    a = np.ascontiguousarray(a)
    if a.shape[0] != a.shape[1]:
        return None  # cholesky.rs:57-59
    n = a.shape[0]
    l = np.zeros((n, n), dtype=a.dtype)
    ok = getattr(lib(), f"oracle_chol_{'canon' if form == 'canon' else 'fast'}_{_suf(a.dtype)}")(_ptr(a), n, _ptr(l))
    return l if ok else None


def chol_solve(l, b):
    """CholeskyDecomposition::solve (cholesky.rs:116-144)."""
    l = np.ascontiguousarray(l)
    b = np.ascontiguousarray(b)
    assert l.shape[0] == b.shape[0]
    x = np.empty_like(b)
    getattr(lib(), f"oracle_chol_solve_{_suf(l.dtype)}")(_ptr(l), l.shape[0], _ptr(b), b.shape[1], _ptr(x))
    return x


def qr(a, form="fast"):
    """QRDecomposition::new (src/decomp/qr.rs:26-106): (packed qr, rdiag)."""
    a = np.ascontiguousarray(a)
    m, n = a.shape
    packed = np.empty_like(a)
    rdiag = np.empty(min(m, n), dtype=a.dtype)
    getattr(lib(), f"oracle_qr_{'canon' if form == 'canon' else 'fast'}_{_suf(a.dtype)}")(_ptr(a), m, n, _ptr(packed),
                                                                                        _ptr(rdiag))
    return packed, rdiag


def qr_is_full_rank(packed, rdiag):
    """is_full_rank (qr.rs:110-117): indexes rdiag[0..cols) -- out of bounds (a panic) when m < n."""
    n = packed.shape[1]
    if n > len(rdiag):
        raise IndexError("index out of bounds: rdiag has min(m, n) entries (qr.rs:112)")
    return not bool(np.any(rdiag[:n] == 0))


def qr_get_h(packed):
    """get_h (qr.rs:121-135): lower trapezoid holding the Householder vectors."""
    return np.tril(packed)


def qr_get_r(packed, rdiag):
    """get_r (qr.rs:138-152): m x n, strict upper part of qr, rdiag on the diagonal."""
    r = np.triu(packed, 1)
    k = len(rdiag)
    r[np.arange(k), np.arange(k)] = rdiag
    return r


def qr_get_q(packed, rdiag):
    """get_q (qr.rs:155-194): m x m."""
    packed = np.ascontiguousarray(packed)
    m, n = packed.shape
    q = np.empty((m, m), dtype=packed.dtype)
    getattr(lib(), f"oracle_qr_get_q_{_suf(packed.dtype)}")(_ptr(packed), _ptr(np.ascontiguousarray(rdiag)), m, n, _ptr(q))
    return q


def qr_solve(packed, rdiag, b):
    """solve (qr.rs:199-238): None when not full rank; the result is Matrix::new(cols, nx, <m*nx values>), which panics
    (AssertionError here) unless m == n -- the reference's least-squares solve only ever returns for square systems."""
    packed = np.ascontiguousarray(packed)
    b = np.ascontiguousarray(b)
    m, n = packed.shape
    assert b.shape[0] == m  # qr.rs:200
    if not qr_is_full_rank(packed, rdiag):
        return None
    x = np.empty_like(b)
    getattr(lib(), f"oracle_qr_solve_{_suf(packed.dtype)}")(_ptr(packed), _ptr(np.ascontiguousarray(rdiag)), m, n, _ptr(b),
                                                            b.shape[1], _ptr(x))
    assert n * b.shape[1] == x.size, "Matrix::new(cols, nx, data): rows * cols != data.len() (mod.rs:208)"
    return x.reshape(n, b.shape[1])
