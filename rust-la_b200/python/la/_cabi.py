"""ctypes binding of include/la_cabi.h (libla_b200.so).  There is NO CPU fallback: if the CUDA library is missing or
no device is usable, every compute call raises."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LA_B200_LIB") or os.path.normpath(os.path.join(_PKG, "..", "..", "libla_b200.so"))

LA_OK, LA_ERR_INVALID, LA_ERR_CUDA, LA_ERR_NOMEM, LA_ERR_NO_DEVICE, LA_ERR_UNSUPPORTED = range(6)
LA_GEMM_ASSIGN, LA_GEMM_SUB, LA_GEMM_ADD = 0, 1, 2
LA_F32_3XTF32, LA_F32_TF32 = 0, 1
LA_EW_ADD, LA_EW_SUB, LA_EW_MUL, LA_EW_DIV, LA_EW_SCALE, LA_EW_NEG = range(6)
LA_RED_SUMSQ, LA_RED_ABS_SUM, LA_RED_ABS_MAX, LA_RED_DOT = range(4)


class LaError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"la_b200 status {status}: {text}")
        self.status = status


_sz = ctypes.c_size_t
_p = ctypes.c_void_p
_i = ctypes.c_int
_u64 = ctypes.c_uint64
_pp = ctypes.POINTER(ctypes.c_void_p)
_pi = ctypes.POINTER(ctypes.c_int)
_psz = ctypes.POINTER(ctypes.c_size_t)
LA_MG_HANDLE_BYTES = 256

# name -> (argtypes, restype); the C header is the single source of truth, tests/test_cabi_symbols.py cross-checks it
SIGNATURES = {
    "la_version": ([], _i),
    "la_last_error": ([], ctypes.c_char_p),
    "la_device_count": ([_pi], _i),
    "la_device_sm_count": ([_i, _pi], _i),
    "la_sync": ([_i], _i),
    "la_buf_alloc": ([_sz, _i, _pp], _i),
    "la_buf_free": ([_p], _i),
    "la_buf_upload": ([_p, _sz, _p, _sz], _i),
    "la_buf_download": ([_p, _sz, _p, _sz], _i),
    "la_buf_copy": ([_p, _p, _sz], _i),
    "la_buf_device_ptr": ([_p], _p),
    "la_buf_bytes": ([_p], _sz),
    "la_buf_device": ([_p], _i),
    "la_host_alloc": ([_sz, _pp], _i),
    "la_host_free": ([_p], _i),
    "la_gemm_f64": ([_p, _p, _p, _sz, _sz, _sz], _i),
    "la_gemm_f32": ([_p, _p, _p, _sz, _sz, _sz], _i),
    "la_gemm_f64_host": ([_p, _p, _p, _sz, _sz, _sz], _i),
    "la_gemm_f32_host": ([_p, _p, _p, _sz, _sz, _sz], _i),
    "la_gemm_i64_host": ([_p, _p, _p, _sz, _sz, _sz], _i),
    "la_gemm_f64_dev": ([_p, _sz, _p, _sz, _p, _sz, _sz, _sz, _sz, _i, _p], _i),
    "la_gemm_f32_dev": ([_p, _sz, _p, _sz, _p, _sz, _sz, _sz, _sz, _i, _p], _i),
    "la_gemm_f64_mg": ([_i, _pi, _p, _p, _p, _sz, _sz, _sz], _i),
    "la_gemm_f32_mg": ([_i, _pi, _p, _p, _p, _sz, _sz, _sz], _i),
    "la_mg_shard": ([_i, _i, _sz, _sz, _sz, _psz, _psz, _psz, _psz], _i),
    "la_mg_create": ([_i, _i, _i, _sz, _sz, _sz, _pp], _i),
    "la_mg_handle": ([_p, _p], _i),
    "la_mg_connect": ([_p, _p], _i),
    "la_mg_destroy": ([_p], _i),
    "la_mg_b_block": ([_p, _pp, _psz, _psz, _psz], _i),
    "la_gemm_f64_mg_rank": ([_p, _p, _sz, _p, _sz, _sz, _p], _i),
    "la_gemm_f32_mg_rank": ([_p, _p, _sz, _p, _sz, _sz, _p], _i),
    "la_gemm_f64_mg_rank_host": ([_p, _p, _p, _sz, _p, _sz], _i),
    "la_gemm_f32_mg_rank_host": ([_p, _p, _p, _sz, _p, _sz], _i),
    "la_mg_reserve": ([_p, _sz], _i),
    "la_mg_quiesce": ([_p, _p], _i),
    "la_lu_factor_f64": ([_p, _sz, _sz, _p, _pi], _i),
    "la_lu_factor_f32": ([_p, _sz, _sz, _p, _pi], _i),
    "la_lu_factor_f64_host": ([_p, _p, _sz, _sz, _p, _pi], _i),
    "la_lu_factor_f32_host": ([_p, _p, _sz, _sz, _p, _pi], _i),
    "la_lu_factor_f64_dev": ([_p, _sz, _sz, _p, _p, _p], _i),
    "la_lu_factor_f32_dev": ([_p, _sz, _sz, _p, _p, _p], _i),
    "la_lu_mg_create": ([_i, _p, _sz, _p], _i),
    "la_lu_mg_destroy": ([_p], _i),
    "la_lu_mg_devices": ([_p, _pi], _i),
    "la_lu_mg_plan": ([_sz, _i, _i, _pi, _pi, _pi, _psz], _i),
    "la_lu_mg_upload_f64": ([_p, _p], _i),
    "la_lu_mg_fill_hash_f64": ([_p, ctypes.c_uint64], _i),
    "la_lu_mg_factor_f64": ([_p], _i),
    "la_lu_mg_sync": ([_p], _i),
    "la_lu_mg_download_f64": ([_p, _p, _p, _pi], _i),
    "la_lu_mg_last_ms": ([_p, _p], _i),
    "la_lu_factor_f64_mg": ([_i, _p, _p, _p, _sz, _p, _pi], _i),
    "la_lu_is_nonsingular_f64": ([_p, _sz, _pi], _i),
    "la_lu_is_nonsingular_f32": ([_p, _sz, _pi], _i),
    "la_lu_det_f64": ([_p, _sz, _i, ctypes.POINTER(ctypes.c_double)], _i),
    "la_lu_det_f32": ([_p, _sz, _i, ctypes.POINTER(ctypes.c_float)], _i),
    "la_lu_solve_f64": ([_p, _sz, _sz, _p, _p, _sz, _p], _i),
    "la_lu_solve_f32": ([_p, _sz, _sz, _p, _p, _sz, _p], _i),
    "la_lu_solve_f64_host": ([_p, _sz, _sz, _p, _p, _sz, _p], _i),
    "la_lu_solve_f32_host": ([_p, _sz, _sz, _p, _p, _sz, _p], _i),
    "la_lu_solve_f64_dev": ([_p, _sz, _p, _p, _sz, _p, _p], _i),
    "la_lu_solve_f32_dev": ([_p, _sz, _p, _p, _sz, _p, _p], _i),
    "la_chol_factor_f64": ([_p, _sz, _pi], _i),
    "la_chol_factor_f32": ([_p, _sz, _pi], _i),
    "la_chol_factor_f64_host": ([_p, _p, _sz, _pi], _i),
    "la_chol_factor_f32_host": ([_p, _p, _sz, _pi], _i),
    "la_chol_factor_f64_dev": ([_p, _sz, _p, _p], _i),
    "la_chol_factor_f32_dev": ([_p, _sz, _p, _p], _i),
    "la_chol_solve_f64_dev": ([_p, _sz, _p, _sz, _p, _p], _i),
    "la_chol_solve_f32_dev": ([_p, _sz, _p, _sz, _p, _p], _i),
    "la_chol_solve_f64": ([_p, _sz, _p, _sz, _p], _i),
    "la_chol_solve_f32": ([_p, _sz, _p, _sz, _p], _i),
    "la_chol_solve_f64_host": ([_p, _sz, _p, _sz, _p], _i),
    "la_chol_solve_f32_host": ([_p, _sz, _p, _sz, _p], _i),
    "la_qr_tmat_elems": ([_sz, _sz, _i, _sz, _psz], _i),
    "la_qr_factor_f64": ([_p, _sz, _sz, _p, _p], _i),
    "la_qr_factor_f32": ([_p, _sz, _sz, _p, _p], _i),
    "la_qr_factor_f64_host": ([_p, _p, _p, _sz, _sz], _i),
    "la_qr_factor_f32_host": ([_p, _p, _p, _sz, _sz], _i),
    "la_qr_factor_f64_dev": ([_p, _sz, _sz, _p, _p, _p], _i),
    "la_qr_factor_f32_dev": ([_p, _sz, _sz, _p, _p, _p], _i),
    "la_qr_get_r_f64": ([_p, _sz, _sz, _p, _p], _i),
    "la_qr_get_r_f32": ([_p, _sz, _sz, _p, _p], _i),
    "la_qr_get_q_f64": ([_p, _sz, _sz, _p, _p], _i),
    "la_qr_get_q_f32": ([_p, _sz, _sz, _p, _p], _i),
    "la_qr_solve_f64": ([_p, _sz, _sz, _p, _p, _sz, _p], _i),
    "la_qr_solve_f32": ([_p, _sz, _sz, _p, _p, _sz, _p], _i),
    "la_elementwise_f64": ([_i, _p, _p, ctypes.c_double, _p, _sz], _i),
    "la_elementwise_f32": ([_i, _p, _p, ctypes.c_float, _p, _sz], _i),
    "la_elementwise_f64_dev": ([_i, _p, _p, ctypes.c_double, _p, _sz, _p], _i),
    "la_elementwise_f32_dev": ([_i, _p, _p, ctypes.c_float, _p, _sz, _p], _i),
    "la_reduce_f64": ([_i, _p, _p, _sz, ctypes.POINTER(ctypes.c_double)], _i),
    "la_reduce_f32": ([_i, _p, _p, _sz, ctypes.POINTER(ctypes.c_float)], _i),
    "la_reduce_f64_dev": ([_i, _p, _p, _sz, ctypes.POINTER(ctypes.c_double), _p], _i),
    "la_reduce_f32_dev": ([_i, _p, _p, _sz, ctypes.POINTER(ctypes.c_float), _p], _i),
    "la_identity_f64": ([_p, _sz], _i),
    "la_identity_f32": ([_p, _sz], _i),
    "la_transpose_f64": ([_p, _p, _sz, _sz], _i),
    "la_transpose_f32": ([_p, _p, _sz, _sz], _i),
    "la_permute_rows_f64": ([_p, _sz, _sz, _p, _sz, _p], _i),
    "la_permute_rows_f32": ([_p, _sz, _sz, _p, _sz, _p], _i),
    "la_fill_hash_f64_dev": ([_p, _sz, _u64, _u64, _p], _i),
    "la_fill_hash_f32_dev": ([_p, _sz, _u64, _u64, _p], _i),
    "la_set_gemm_f32_mode": ([_i], _i),
    "la_get_gemm_f32_mode": ([_pi], _i),
    "la_debug_set_gemm_path": ([_i], _i),
    "la_debug_set_gemm_f32_path": ([_i], _i),
}

_lib = None


def lib():
    """Loads libla_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LaError(LA_ERR_NO_DEVICE,
                          f"{LIB_PATH} is missing: build it with `make -C rust-la_b200` (there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (args, res) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def check(status):
    if status != LA_OK:
        raise LaError(status, lib().la_last_error().decode("utf-8", "replace"))


def device_count():
    n = _i(0)
    st = lib().la_device_count(ctypes.byref(n))
    return n.value if st == LA_OK else 0
