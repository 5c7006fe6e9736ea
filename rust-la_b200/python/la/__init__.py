"""`la` -- host-side mirror of the rust-la crate API for the dense hot path, backed by libla_b200.so (sm_100a CUDA).

Mirrors (reference, relative to /root/reference):
  Matrix<T>            src/matrix/mod.rs:26-30     Matrix.new :207-211, rows :256, cols :260, get_data :264, get :557-560,
                                                   id :416-426, operator * :957-998, det/solve/inverse/is_singular :1025-1047
  Matrix::mmul         src/matrix/mmatrix.rs:82-98
  LUDecomposition<T>   src/decomp/lu.rs:95-278     new, is_singular, is_non_singular, get_l, get_u, get_p, get_piv, det, solve
  CholeskyDecomposition<T>  src/decomp/cholesky.rs:52-144   new (None unless square, symmetric, positive definite), get_l, solve
  m!                   src/macros.rs:39-42         -> m("1, 2; 3, 4") / m([[1, 2], [3, 4]])
  ApproxEq             src/approxeq.rs:34-47       absolute 1e-6

Error conventions are the reference's: contract violations "panic" (raise `Panic`, an AssertionError), numerical
singularity is `None`.  There is no CPU fallback: every multiplication / factorisation / solve runs in CUDA through the
C ABI in include/la_cabi.h, and raises `LaError` when the library or a B200 is not available.
"""
import ctypes

import numpy as np

from . import _cabi
from ._cabi import LaError, lib, check  # noqa: F401

__all__ = ["Matrix", "DeviceMatrix", "LUDecomposition", "CholeskyDecomposition", "QRDecomposition", "m", "Panic", "LaError", "APPROX_EPS"]

APPROX_EPS = 1e-6  # src/approxeq.rs:20,36


class Panic(AssertionError):
    """The reference `assert!`s on shape/contract violations; the mirror raises this instead of unwinding."""


def _assert(cond, what):
    if not cond:
        raise Panic(f"assertion failed: {what}")


_SUF = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32"}


def _suffix(dtype):
    try:
        return _SUF[np.dtype(dtype)]
    except KeyError:
        raise TypeError(f"the CUDA path supports f32/f64 (and i64 for Mul); got {np.dtype(dtype)}") from None


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class Matrix:
    """Row-major dense matrix: `{ no_rows, data }` with element (r, c) at data[r * cols + c]."""

    __slots__ = ("no_rows", "data")

    def __init__(self, no_rows, data):
        self.no_rows = int(no_rows)
        self.data = data  # 1-D contiguous numpy array (the `Vec<T>`)

    # ---- constructors ---------------------------------------------------------------------------
    @staticmethod
    def new(no_rows, no_cols, data, dtype=None):
        """Matrix::new, src/matrix/mod.rs:207-211."""
        arr = np.ascontiguousarray(np.asarray(data, dtype=dtype)).reshape(-1)
        if arr.dtype not in (np.float64, np.float32, np.int64):
            arr = arr.astype(np.int64 if np.issubdtype(arr.dtype, np.integer) else np.float64)
        _assert(no_rows * no_cols == arr.size, "no_rows * no_cols == data.len()")
        _assert(no_rows > 0 and no_cols > 0, "no_rows > 0 && no_cols > 0")
        return Matrix(no_rows, arr)

    @staticmethod
    def id(m, n, dtype=np.float64):
        """Matrix::id, src/matrix/mod.rs:416-426."""
        d = np.zeros(m * n, dtype=dtype)
        d[: min(m, n) * n : n + 1] = 1
        return Matrix(m, d)

    @staticmethod
    def from_numpy(a):
        a = np.asarray(a)
        _assert(a.ndim == 2, "2-D array")
        return Matrix.new(a.shape[0], a.shape[1], a)

    # ---- accessors ------------------------------------------------------------------------------
    def rows(self):
        return self.no_rows

    def cols(self):
        return self.data.size // self.no_rows

    def get_data(self):
        return self.data

    def get(self, row, col):
        _assert(row < self.no_rows and col < self.cols(), "row < rows && col < cols")
        return self.data[row * self.cols() + col]

    def to_numpy(self):
        return self.data.reshape(self.no_rows, self.cols())

    def t(self):
        return Matrix(self.cols(), np.ascontiguousarray(self.to_numpy().T).reshape(-1))

    def __eq__(self, other):  # #[derive(PartialEq)]
        return isinstance(other, Matrix) and self.no_rows == other.no_rows and self.data.size == other.data.size and \
            bool(np.array_equal(self.data, other.data))

    def approx_eq(self, other):
        """src/matrix/mod.rs:1141-1147 with ApproxEq = absolute 1e-6."""
        if self.no_rows != other.no_rows or self.data.size != other.data.size:
            return False
        return bool(np.all(np.abs(self.data - other.data) < APPROX_EPS))

    def __repr__(self):
        return f"Matrix({self.no_rows}x{self.cols()}, {self.data.dtype})"

    # ---- Mul: src/matrix/mod.rs:957-998 -----------------------------------------------------------
    def __mul__(self, other):
        _assert(isinstance(other, Matrix), "rhs is a Matrix")
        _assert(self.cols() == other.no_rows, "self.cols() == m.no_rows")
        _assert(self.data.dtype == other.data.dtype, "same element type")
        out = np.empty(self.no_rows * other.cols(), dtype=self.data.dtype)  # alloc_dirty_vec
        self._gemm_into(other, out)
        return Matrix(self.no_rows, out)

    def mmul(self, other, dst):
        """Matrix::mmul, src/matrix/mmatrix.rs:82-98: product into the caller's `dst`."""
        _assert(self.cols() == other.no_rows, "self.cols() == m.no_rows")
        _assert(dst.rows() == self.no_rows, "dst.rows() == self.no_rows")
        _assert(dst.cols() == other.cols(), "dst.cols() == m.cols()")
        _assert(self.data.dtype == other.data.dtype == dst.data.dtype, "same element type")
        self._gemm_into(other, dst.data)
        return dst

    def _gemm_into(self, other, out):
        L = lib()
        if self.data.dtype == np.int64:
            fn = L.la_gemm_i64_host
        else:
            fn = getattr(L, f"la_gemm_{_suffix(self.data.dtype)}_host")
        check(fn(_ptr(self.data), _ptr(other.data), _ptr(out), self.no_rows, self.cols(), other.cols()))

    # ---- LU callers: src/matrix/mod.rs:1025-1047 (each call re-factorises, like the reference) -------------------
    def det(self):
        _assert(self.cols() == self.no_rows, "self.cols() == self.no_rows")
        return LUDecomposition.new(self).det()

    def solve(self, b):
        return LUDecomposition.new(self).solve(b)

    def inverse(self):
        _assert(self.no_rows == self.cols(), "self.no_rows == self.cols()")
        return LUDecomposition.new(self).solve(Matrix.id(self.no_rows, self.no_rows, self.data.dtype))

    def is_singular(self):
        return not self.is_non_singular()

    def is_non_singular(self):
        _assert(self.no_rows == self.cols(), "self.no_rows == self.cols()")
        return LUDecomposition.new(self).is_non_singular()

    def pinverse(self):
        """src/matrix/mod.rs:1049-1057: `(r.t() * &r).inverse().unwrap() * &self.t()` with r = QRDecomposition::get_r;
        the whole chain runs on device-resident intermediates."""
        return DeviceMatrix.from_matrix(self).pinverse().to_matrix()


class _DeviceBuf:
    """RAII wrapper over la_buf (the device backing of a Matrix / LUDecomposition)."""

    def __init__(self, nbytes, device=0):
        self.handle = ctypes.c_void_p()
        check(lib().la_buf_alloc(nbytes, device, ctypes.byref(self.handle)))
        self.nbytes = nbytes
        self.device = device

    def upload(self, arr):
        check(lib().la_buf_upload(self.handle, 0, _ptr(arr), arr.nbytes))

    def download(self, arr):
        check(lib().la_buf_download(self.handle, 0, _ptr(arr), arr.nbytes))

    def __del__(self):
        try:
            if self.handle:
                lib().la_buf_free(self.handle)
                self.handle = None
        except Exception:
            pass


class LUDecomposition:
    """LUDecomposition<T>, src/decomp/lu.rs:95-101: `{ lu: Matrix<T>, pospivsign: bool, piv: Vec<usize> }`.

    The packed factors stay resident in HBM (la_buf); the host copy is materialised lazily for get_l/get_u/get_lu."""

    def __init__(self, m, n, dtype, buf, piv, pospivsign):
        self._m, self._n, self._dtype = m, n, np.dtype(dtype)
        self._buf = buf
        self.piv = piv
        self.pospivsign = bool(pospivsign)
        self._lu_host = None

    @staticmethod
    def new(a, device=0):
        """LUDecomposition::new, src/decomp/lu.rs:104-168 (factorises a copy; `a` is untouched)."""
        m, n = a.rows(), a.cols()
        suf = _suffix(a.data.dtype)
        buf = _DeviceBuf(a.data.nbytes, device)
        buf.upload(a.data)  # ludata = a.get_data().clone()
        piv = np.empty(m, dtype=np.uint64)
        sign = ctypes.c_int(1)
        check(getattr(lib(), f"la_lu_factor_{suf}")(buf.handle, m, n, _ptr(piv), ctypes.byref(sign)))
        return LUDecomposition(m, n, a.data.dtype, buf, piv, sign.value)

    @staticmethod
    def new_on_devices(a, devices):
        """The same factorisation spread over several GPUs (la_lu_factor_f64_mg): 128-column blocks dealt
        round-robin, the panel owner's block column copied to every device.  fp64, square.  The packed factors land on
        devices[0], so solve / det / get_l work as after `new`."""
        _assert(a.rows() == a.cols(), "new_on_devices: the matrix must be square")
        _assert(a.data.dtype == np.float64, "new_on_devices: fp64 only")
        _assert(len(devices) > 0, "new_on_devices: empty device list")
        from . import sharding
        n = a.rows()
        lu, piv, sign = sharding.lu_factor_mg(a.data.reshape(n, n), list(devices))
        buf = _DeviceBuf(lu.nbytes, devices[0])
        buf.upload(lu.reshape(-1))
        dec = LUDecomposition(n, n, np.float64, buf, piv, sign)
        dec._lu_host = Matrix(n, lu.reshape(-1))
        return dec

    def get_lu(self):
        if self._lu_host is None:
            h = np.empty(self._m * self._n, dtype=self._dtype)
            self._buf.download(h)
            self._lu_host = Matrix(self._m, h)
        return self._lu_host

    def is_singular(self):
        return not self.is_non_singular()

    def is_non_singular(self):
        """src/decomp/lu.rs:174-182 (indexes lu[j*n+j] for j < n: the reference panics out of bounds when m < n)."""
        _assert(self._m >= self._n, "index out of bounds: lu[j*n+j] for m < n")
        out = ctypes.c_int(0)
        check(getattr(lib(), f"la_lu_is_nonsingular_{_suffix(self._dtype)}")(self._buf.handle, self._n, ctypes.byref(out)))
        return bool(out.value)

    def get_l(self):
        """src/decomp/lu.rs:184-202 (host-side unpack)."""
        lu = self.get_lu().to_numpy()
        nn = min(self._m, self._n)
        l = np.tril(lu[:, :nn], -1)
        l[np.arange(nn), np.arange(nn)] = 1
        return Matrix(self._m, np.ascontiguousarray(l).reshape(-1))

    def get_u(self):
        """src/decomp/lu.rs:204-215."""
        lu = self.get_lu().to_numpy()
        mm = min(self._m, self._n)
        return Matrix(mm, np.ascontiguousarray(np.triu(lu[:mm, :])).reshape(-1))

    def get_p(self):
        """src/decomp/lu.rs:217-220: id(len, len).permute_rows(piv)."""
        ln = self.piv.size
        p = np.zeros((ln, ln), dtype=self._dtype)
        p[np.arange(ln), self.piv.astype(np.int64)] = 1
        return Matrix(ln, p.reshape(-1))

    def get_piv(self):
        return self.piv

    def det(self):
        """src/decomp/lu.rs:224-232."""
        _assert(self._m == self._n, "self.lu.rows() == self.lu.cols()")
        suf = _suffix(self._dtype)
        out = ctypes.c_double(0) if suf == "f64" else ctypes.c_float(0)
        check(getattr(lib(), f"la_lu_det_{suf}")(self._buf.handle, self._n, int(self.pospivsign), ctypes.byref(out)))
        return self._dtype.type(out.value)

    def solve(self, b):
        """src/decomp/lu.rs:237-278: Some(X) or None when singular."""
        _assert(b.rows() == self._m, "b.rows() == m")
        _assert(b.data.dtype == self._dtype, "same element type")
        if not self.is_non_singular():
            return None
        _assert(self._m == self._n, "solve needs a square factorisation (lu.rs:257-275 index with n)")
        nx = b.cols()
        suf = _suffix(self._dtype)
        bbuf = _DeviceBuf(b.data.nbytes, self._buf.device)
        xbuf = _DeviceBuf(b.data.nbytes, self._buf.device)
        bbuf.upload(b.data)
        check(getattr(lib(), f"la_lu_solve_{suf}")(self._buf.handle, self._m, self._n, _ptr(self.piv), bbuf.handle, nx,
                                                   xbuf.handle))
        x = np.empty(self._m * nx, dtype=self._dtype)
        xbuf.download(x)
        return Matrix(self._m, x)


class CholeskyDecomposition:
    """CholeskyDecomposition<T>, src/decomp/cholesky.rs:52-54: `{ l: Matrix<T> }` with A = L L'.  L stays in HBM."""

    def __init__(self, n, dtype, buf):
        self._n, self._dtype, self._buf = n, np.dtype(dtype), buf
        self._l_host = None

    @staticmethod
    def new(a, device=0):
        """cholesky.rs:56-110: None unless `a` is square, symmetric (exact) and positive definite."""
        if a.rows() != a.cols():
            return None  # :57-59
        n, suf = a.rows(), _suffix(a.data.dtype)
        buf = _DeviceBuf(a.data.nbytes, device)
        buf.upload(a.data)
        ok = ctypes.c_int(0)
        check(getattr(lib(), f"la_chol_factor_{suf}")(buf.handle, n, ctypes.byref(ok)))
        return CholeskyDecomposition(n, a.data.dtype, buf) if ok.value else None

    def get_l(self):
        if self._l_host is None:
            h = np.empty(self._n * self._n, dtype=self._dtype)
            self._buf.download(h)
            self._l_host = Matrix(self._n, h)
        return self._l_host

    def solve(self, b):
        """cholesky.rs:116-144."""
        _assert(b.rows() == self._n, "l.rows() == b.rows()")  # :118
        _assert(b.data.dtype == self._dtype, "same element type")
        nx = b.cols()
        bbuf = _DeviceBuf(b.data.nbytes, self._buf.device)
        xbuf = _DeviceBuf(b.data.nbytes, self._buf.device)
        bbuf.upload(b.data)
        check(getattr(lib(), f"la_chol_solve_{_suffix(self._dtype)}")(self._buf.handle, self._n, bbuf.handle, nx, xbuf.handle))
        x = np.empty(self._n * nx, dtype=self._dtype)
        xbuf.download(x)
        return Matrix(self._n, x)


class QRDecomposition:
    """QRDecomposition<T>, src/decomp/qr.rs:20-23: `{ qr: Matrix<T>, rdiag: Vec<T> }`.  The packed factors stay in HBM
    together with the block factors T' of the compact-WY form (needed by get_q); host copies are made lazily."""

    def __init__(self, m, n, dtype, buf, rdiag_buf, tmat_buf):
        self._m, self._n, self._dtype = m, n, np.dtype(dtype)
        self._buf, self._rdiag_buf, self._tmat_buf = buf, rdiag_buf, tmat_buf
        self._qr_host = None
        self.rdiag = np.empty(min(m, n), dtype=self._dtype)
        rdiag_buf.download(self.rdiag)

    @staticmethod
    def new(a, device=0):
        """qr.rs:26-43."""
        m, n, suf = a.rows(), a.cols(), _suffix(a.data.dtype)
        isz = a.data.dtype.itemsize
        buf = _DeviceBuf(a.data.nbytes, device)
        buf.upload(a.data)  # qrdata = m.get_data().clone()
        te = ctypes.c_size_t(0)
        check(lib().la_qr_tmat_elems(m, n, device, isz, ctypes.byref(te)))
        rd = _DeviceBuf(max(min(m, n) * isz, 8), device)
        tm = _DeviceBuf(te.value * isz, device)
        check(getattr(lib(), f"la_qr_factor_{suf}")(buf.handle, m, n, rd.handle, tm.handle))
        return QRDecomposition(m, n, a.data.dtype, buf, rd, tm)

    def get_qr(self):
        if self._qr_host is None:
            h = np.empty(self._m * self._n, dtype=self._dtype)
            self._buf.download(h)
            self._qr_host = Matrix(self._m, h)
        return self._qr_host

    def is_full_rank(self):
        """qr.rs:110-117: loops j over 0..cols and indexes rdiag[j] -- out of bounds (a panic) when m < n."""
        for j in range(self._n):
            _assert(j < self.rdiag.size, "index out of bounds: rdiag[j] (qr.rs:112)")
            if self.rdiag[j] == 0:
                return False
        return True

    def get_h(self):
        """qr.rs:121-135 (host-side unpack): the lower trapezoid holding the Householder vectors."""
        return Matrix(self._m, np.ascontiguousarray(np.tril(self.get_qr().to_numpy())).reshape(-1))

    def get_r(self):
        """qr.rs:138-152."""
        return self.get_r_device().to_matrix()

    def get_r_device(self):
        out = DeviceMatrix(self._m, self._n, self._dtype, _DeviceBuf(self._m * self._n * self._dtype.itemsize, self._buf.device))
        check(getattr(lib(), f"la_qr_get_r_{_suffix(self._dtype)}")(self._buf.handle, self._m, self._n, self._rdiag_buf.handle,
                                                                    out.buf.handle))
        return out

    def get_q(self):
        """qr.rs:155-194: m x m, the block reflectors applied in reverse order on the device."""
        out = DeviceMatrix(self._m, self._m, self._dtype, _DeviceBuf(self._m * self._m * self._dtype.itemsize, self._buf.device))
        check(getattr(lib(), f"la_qr_get_q_{_suffix(self._dtype)}")(self._buf.handle, self._m, self._n, self._tmat_buf.handle,
                                                                    out.buf.handle))
        return out.to_matrix()

    def solve(self, b):
        """qr.rs:199-238, quirks included: None unless full rank; the result is `Matrix::new(cols, nx, <m * nx values>)`,
        which panics unless m == n (:237)."""
        _assert(b.rows() == self._m, "b.rows() == self.qr.rows()")  # :200
        _assert(b.data.dtype == self._dtype, "same element type")
        if not self.is_full_rank():
            return None
        nx = b.cols()
        _assert(self._n * nx == self._m * nx, "no_rows * no_cols == data.len()")  # Matrix::new, mod.rs:208
        bbuf = _DeviceBuf(b.data.nbytes, self._buf.device)
        xbuf = _DeviceBuf(b.data.nbytes, self._buf.device)
        bbuf.upload(b.data)
        check(getattr(lib(), f"la_qr_solve_{_suffix(self._dtype)}")(self._buf.handle, self._m, self._n, self._rdiag_buf.handle,
                                                                    bbuf.handle, nx, xbuf.handle))
        x = np.empty(self._m * nx, dtype=self._dtype)
        xbuf.download(x)
        return Matrix(self._n, x)


class DeviceMatrix:
    """A `Matrix<T>` whose `data` lives in HBM (la_buf): SURVEY.md section 8(f) rank 1 -- chains such as pinverse's
    `(r.t() * &r).inverse() * &a.t()` (src/matrix/mod.rs:1049-1057) stay on the device between operations.

    Mirrors the subset of the Matrix API that has a device kernel: `t` (mod.rs:653-669), operator `*` (:957-998), `id`
    (:416-426), `permute_rows` (:757-759), `inverse` (:1034-1037, via LUDecomposition + solve with the identity); same
    panics, same `None` on singularity."""

    __slots__ = ("no_rows", "no_cols", "dtype", "buf")

    def __init__(self, no_rows, no_cols, dtype, buf):
        self.no_rows, self.no_cols, self.dtype, self.buf = int(no_rows), int(no_cols), np.dtype(dtype), buf

    @staticmethod
    def from_matrix(a, device=0):
        buf = _DeviceBuf(a.data.nbytes, device)
        buf.upload(a.data)
        return DeviceMatrix(a.rows(), a.cols(), a.data.dtype, buf)

    @staticmethod
    def id(n, dtype=np.float64, device=0):
        dtype = np.dtype(dtype)
        buf = _DeviceBuf(n * n * dtype.itemsize, device)
        check(getattr(lib(), f"la_identity_{_suffix(dtype)}")(buf.handle, n))
        return DeviceMatrix(n, n, dtype, buf)

    def to_matrix(self):
        h = np.empty(self.no_rows * self.no_cols, dtype=self.dtype)
        self.buf.download(h)
        return Matrix(self.no_rows, h)

    def rows(self):
        return self.no_rows

    def cols(self):
        return self.no_cols

    def _like(self, rows, cols):
        return DeviceMatrix(rows, cols, self.dtype, _DeviceBuf(rows * cols * self.dtype.itemsize, self.buf.device))

    def t(self):
        out = self._like(self.no_cols, self.no_rows)
        check(getattr(lib(), f"la_transpose_{_suffix(self.dtype)}")(self.buf.handle, out.buf.handle, self.no_rows,
                                                                    self.no_cols))
        return out

    def permute_rows(self, rows):
        idx = np.ascontiguousarray(rows, dtype=np.uint64)
        _assert(idx.size > 0 and int(idx.max()) < self.no_rows, "row index out of bounds")  # sub_matrix panics
        out = self._like(idx.size, self.no_cols)
        check(getattr(lib(), f"la_permute_rows_{_suffix(self.dtype)}")(self.buf.handle, self.no_rows, self.no_cols,
                                                                       _ptr(idx), idx.size, out.buf.handle))
        return out

    def __mul__(self, other):
        _assert(isinstance(other, DeviceMatrix), "DeviceMatrix * DeviceMatrix")
        _assert(self.no_cols == other.no_rows, "self.cols() == m.no_rows")  # mod.rs:961
        _assert(self.dtype == other.dtype, "same element type")
        out = self._like(self.no_rows, other.no_cols)
        check(getattr(lib(), f"la_gemm_{_suffix(self.dtype)}")(self.buf.handle, other.buf.handle, out.buf.handle,
                                                               self.no_rows, self.no_cols, other.no_cols))
        return out

    def inverse(self):
        """mod.rs:1034-1037: `LUDecomposition::new(self).solve(&Matrix::id(...))`; None when singular."""
        _assert(self.no_rows == self.no_cols, "self.no_rows == self.cols()")
        n, suf = self.no_rows, _suffix(self.dtype)
        lu = self._like(n, n)
        check(lib().la_buf_copy(lu.buf.handle, self.buf.handle, n * n * self.dtype.itemsize))  # the factorisation's copy
        piv = np.empty(n, dtype=np.uint64)
        sign = ctypes.c_int(1)
        check(getattr(lib(), f"la_lu_factor_{suf}")(lu.buf.handle, n, n, _ptr(piv), ctypes.byref(sign)))
        ok = ctypes.c_int(0)
        check(getattr(lib(), f"la_lu_is_nonsingular_{suf}")(lu.buf.handle, n, ctypes.byref(ok)))
        if not ok.value:
            return None
        eye = DeviceMatrix.id(n, self.dtype, self.buf.device)
        out = self._like(n, n)
        check(getattr(lib(), f"la_lu_solve_{suf}")(lu.buf.handle, n, n, _ptr(piv), eye.buf.handle, n, out.buf.handle))
        return out

    # ---- elementwise operators and norms (mod.rs:487-527, :853-929, :1059-1115) on device-resident data ----
    def _same_shape(self, other):
        _assert(isinstance(other, DeviceMatrix), "rhs is a DeviceMatrix")
        _assert(self.no_rows == other.no_rows, "self.no_rows == m.no_rows")
        _assert(self.no_cols == other.no_cols, "self.cols() == m.cols()")
        _assert(self.dtype == other.dtype, "same element type")

    def _elementwise(self, op, other=None, scalar=0.0):
        out = self._like(self.no_rows, self.no_cols)
        check(getattr(lib(), f"la_elementwise_{_suffix(self.dtype)}")(op, self.buf.handle, other.buf.handle if other else None,
                                                                      scalar, out.buf.handle, self.no_rows * self.no_cols))
        return out

    def __add__(self, other):
        self._same_shape(other)
        return self._elementwise(_cabi.LA_EW_ADD, other)

    def __sub__(self, other):
        self._same_shape(other)
        return self._elementwise(_cabi.LA_EW_SUB, other)

    def __neg__(self):
        return self._elementwise(_cabi.LA_EW_NEG)

    def scale(self, factor):
        return self._elementwise(_cabi.LA_EW_SCALE, None, float(factor))

    def elem_mul(self, other):
        self._same_shape(other)
        return self._elementwise(_cabi.LA_EW_MUL, other)

    def elem_div(self, other):
        self._same_shape(other)
        return self._elementwise(_cabi.LA_EW_DIV, other)

    def _reduce(self, kind, other=None):
        out = ctypes.c_double(0) if self.dtype == np.float64 else ctypes.c_float(0)
        check(getattr(lib(), f"la_reduce_{_suffix(self.dtype)}")(kind, self.buf.handle, other.buf.handle if other else None,
                                                                 self.no_rows * self.no_cols, ctypes.byref(out)))
        return self.dtype.type(out.value)

    def frobenius_norm(self):
        return self._reduce(_cabi.LA_RED_SUMSQ)

    def vector_euclidean_norm(self):
        _assert(self.no_cols == 1, "self.cols() == 1")
        return self._reduce(_cabi.LA_RED_SUMSQ)

    def vector_1_norm(self):
        _assert(self.no_cols == 1, "self.cols() == 1")
        return self._reduce(_cabi.LA_RED_ABS_SUM)

    def vector_inf_norm(self):
        _assert(self.no_cols == 1, "self.cols() == 1")
        return self._reduce(_cabi.LA_RED_ABS_MAX)

    def dot(self, other):
        """mod.rs:529-: both are vectors of the same length."""
        _assert(self.no_rows == other.no_rows and self.no_cols == 1 and other.no_cols == 1, "two column vectors of one length")
        return self._reduce(_cabi.LA_RED_DOT, other)

    def pinverse(self):
        """mod.rs:1049-1057: A+ = (R'R)^-1 A' with R from the QR factorisation; panics (unwrap) when R'R is singular."""
        m, n, suf = self.no_rows, self.no_cols, _suffix(self.dtype)
        isz, dev = self.dtype.itemsize, self.buf.device
        qr = self._like(m, n)
        check(lib().la_buf_copy(qr.buf.handle, self.buf.handle, m * n * isz))
        te = ctypes.c_size_t(0)
        check(lib().la_qr_tmat_elems(m, n, dev, isz, ctypes.byref(te)))
        rd, tm = _DeviceBuf(max(min(m, n) * isz, 8), dev), _DeviceBuf(te.value * isz, dev)
        check(getattr(lib(), f"la_qr_factor_{suf}")(qr.buf.handle, m, n, rd.handle, tm.handle))
        r = self._like(m, n)
        check(getattr(lib(), f"la_qr_get_r_{suf}")(qr.buf.handle, m, n, rd.handle, r.buf.handle))
        inv = (r.t() * r).inverse()
        _assert(inv is not None, "called `Option::unwrap()` on a `None` value")
        return inv * self.t()


def m(spec, dtype=None):
    """The `m!` macro (src/macros.rs:39-42): m("1.0, 2.0; 3.0, 4.0") or m([[1.0, 2.0], [3.0, 4.0]]).

    Like the Rust literal, integer literals give an integer matrix and float literals a float matrix."""
    if isinstance(spec, str):
        rows = [r.strip() for r in spec.strip().split(";") if r.strip()]
        parsed = []
        is_float = False
        for r in rows:
            items = [t.strip() for t in r.split(",") if t.strip()]
            is_float |= any(("." in t) or ("e" in t.lower()) or ("nan" in t.lower()) or ("inf" in t.lower()) for t in items)
            parsed.append(items)
        conv = float if is_float else int
        spec = [[conv(t) for t in items] for items in parsed]
    rows = len(spec)
    cols = len(spec[0])
    _assert(all(len(r) == cols for r in spec), "all rows have the same number of columns")
    flat = [v for r in spec for v in r]
    if dtype is None:
        dtype = np.int64 if all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in flat) else np.float64
    return Matrix.new(rows, cols, np.array(flat, dtype=dtype))
