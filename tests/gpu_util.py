"""Helpers shared by the -m gpu tests: thin device-buffer wrappers over the C ABI (no torch needed)."""
import ctypes

import numpy as np

from la import _cabi
from la._cabi import check, lib


def have_gpu():
    return _cabi.device_count() > 0


class DevBuf:
    def __init__(self, nbytes, device=0):
        self.h = ctypes.c_void_p()
        check(lib().la_buf_alloc(int(nbytes), device, ctypes.byref(self.h)))
        self.nbytes = int(nbytes)

    @classmethod
    def from_array(cls, arr):
        arr = np.ascontiguousarray(arr)
        b = cls(arr.nbytes)
        check(lib().la_buf_upload(b.h, 0, arr.ctypes.data, arr.nbytes))
        return b

    def ptr(self, byte_offset=0):
        return ctypes.c_void_p(lib().la_buf_device_ptr(self.h) + byte_offset)

    def to_array(self, shape, dtype, byte_offset=0):
        out = np.empty(shape, dtype=dtype)
        check(lib().la_buf_download(self.h, byte_offset, out.ctypes.data, out.nbytes))
        return out

    def free(self):
        if self.h:
            lib().la_buf_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def fill_hash(buf, count, seed, dtype, first_idx=0, elem_offset=0):
    suf = "f64" if np.dtype(dtype) == np.float64 else "f32"
    check(getattr(lib(), f"la_fill_hash_{suf}_dev")(buf.ptr(elem_offset * np.dtype(dtype).itemsize), count, seed, first_idx,
                                                    None))


def sync():
    check(lib().la_sync(0))


def gemm_dev(a_buf, lda, b_buf, ldb, c_buf, ldc, m, k, n, mode, dtype, a_off=0, b_off=0, c_off=0):
    suf = "f64" if np.dtype(dtype) == np.float64 else "f32"
    isz = np.dtype(dtype).itemsize
    check(getattr(lib(), f"la_gemm_{suf}_dev")(a_buf.ptr(a_off * isz), lda, b_buf.ptr(b_off * isz), ldb,
                                               c_buf.ptr(c_off * isz), ldc, m, k, n, mode, None))


def max_rel_err(got, ref):
    """max |got-ref| / |ref| over elements with ref != 0 (uniform [0,1) inputs never give 0 sums)."""
    ref = np.asarray(ref, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    den = np.where(ref == 0, 1.0, np.abs(ref))
    return float(np.max(np.abs(got - ref) / den))
