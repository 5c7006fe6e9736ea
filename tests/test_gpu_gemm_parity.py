"""GEMM parity: CUDA path (through the C ABI) vs the oracle on the same seeded inputs.
Bar (BASELINE.json): max relative element error <= 1e-12 * k for f64, <= 1e-4 * k for f32 (k = inner dimension)."""
import numpy as np
import pytest

from gpu_util import DevBuf, fill_hash, gemm_dev, max_rel_err, sync
from la import _cabi
from la._cabi import check, lib

pytestmark = pytest.mark.gpu

F64_TOL = 1e-12
F32_TOL = 1e-4

SHAPES = [(1, 1, 1), (2, 2, 2), (3, 5, 7), (64, 64, 64), (128, 128, 128), (129, 130, 131), (127, 17, 255),
          (256, 512, 384), (200, 1000, 136), (512, 512, 512), (130, 4, 260), (1000, 2, 1000), (640, 1024, 768)]


@pytest.mark.parametrize("path", ["auto", "simt", "tma"])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_f64_host_api(oracle, shape, path):
    m, k, n = shape
    if path == "tma" and (k % 2 or n % 2):
        pytest.skip("TMA needs 16-byte aligned rows")
    a = oracle.fill((m, k), 1)
    b = oracle.fill((k, n), 2)
    ref = oracle.gemm(a, b)
    c = np.full((m, n), np.nan)
    check(lib().la_debug_set_gemm_path({"auto": 0, "simt": 1, "tma": 2}[path]))
    try:
        check(lib().la_gemm_f64_host(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
    finally:
        lib().la_debug_set_gemm_path(0)
    assert np.all(np.isfinite(c))
    assert max_rel_err(c, ref) <= F64_TOL * k


@pytest.mark.parametrize("shape", [(3, 5, 7), (128, 128, 128), (257, 300, 129), (512, 1024, 256)])
def test_gemm_f32_host_api(oracle, shape):
    m, k, n = shape
    a = oracle.fill((m, k), 1, np.float32)
    b = oracle.fill((k, n), 2, np.float32)
    ref = oracle.gemm(a, b)
    c = np.full((m, n), np.nan, dtype=np.float32)
    check(lib().la_gemm_f32_host(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
    assert max_rel_err(c, ref) <= F32_TOL * k


@pytest.mark.parametrize("shape", [(128, 32, 256), (512, 1024, 256), (1000, 520, 768), (640, 100, 328), (130, 64, 260),
                                   (2048, 1024, 2048), (256, 36, 512), (4096, 64, 4096), (3000, 40, 5000)])
@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("acc", ["3xtf32", "tf32"])
def test_gemm_f32_tcgen05_tf32(oracle, shape, mode, acc):
    """The tcgen05 kind::tf32 kernel (TMEM accumulators, TMA-fed), forced, vs the fp32 oracle: <= 1e-4 * k relative in the
    opt-in TF32 mode; the default split-compensated mode (three passes over big/small operand parts) is fp32-grade:
    <= 4e-6 + 1.2e-7 * k relative on these positive inputs."""
    m, k, n = shape
    a = oracle.fill((m, k), 1, np.float32)
    b = oracle.fill((k, n), 2, np.float32)
    c0 = oracle.fill((m, n), 4, np.float32)
    ref = oracle.gemm(a, b)
    want = ref if mode == 0 else (c0.astype(np.float64) + ref.astype(np.float64))
    da, db, dc = DevBuf.from_array(a), DevBuf.from_array(b), DevBuf.from_array(c0)
    check(lib().la_debug_set_gemm_f32_path(2))
    check(lib().la_set_gemm_f32_mode(_cabi.LA_F32_TF32 if acc == "tf32" else _cabi.LA_F32_3XTF32))
    try:
        gemm_dev(da, k, db, n, dc, n, m, k, n, mode, np.float32)
        sync()
    finally:
        lib().la_debug_set_gemm_f32_path(0)
        lib().la_set_gemm_f32_mode(_cabi.LA_F32_3XTF32)
    got = dc.to_array((m, n), np.float32)
    assert np.all(np.isfinite(got))
    assert max_rel_err(got, want) <= (F32_TOL * k if acc == "tf32" else 4e-6 + 1.2e-7 * k)


@pytest.mark.parametrize("mode", [0, 2])
def test_gemm_f32_tcgen05_submatrix_views(oracle, mode):
    """ld > width on all three operands: the TMA stores of the epilogue must clip at the view's edge (m, n ragged
    against the 128 x 256 tile and the 32 x 32 store box) and leave the rest of the parent matrix untouched."""
    ld = 1200
    m, k, n = 300, 96, 1000
    big_a = oracle.fill((400, ld), 21, np.float32)
    big_b = oracle.fill((400, ld), 22, np.float32)
    big_c = oracle.fill((400, ld), 23, np.float32)
    a, b = big_a[:m, :k], big_b[:k, :n]
    ref = oracle.gemm(np.ascontiguousarray(a), np.ascontiguousarray(b))
    want = big_c.copy()
    want[:m, :n] = ref if mode == 0 else big_c[:m, :n] + ref
    da, db, dc = DevBuf.from_array(big_a), DevBuf.from_array(big_b), DevBuf.from_array(big_c)
    check(lib().la_debug_set_gemm_f32_path(2))
    try:
        gemm_dev(da, ld, db, ld, dc, ld, m, k, n, mode, np.float32)
        sync()
    finally:
        lib().la_debug_set_gemm_f32_path(0)
    got = dc.to_array((400, ld), np.float32)
    assert np.array_equal(got[m:, :], big_c[m:, :]) and np.array_equal(got[:, n:], big_c[:, n:])
    assert max_rel_err(got[:m, :n], want[:m, :n]) <= F32_TOL * k


@pytest.mark.parametrize("acc", ["3xtf32", "tf32"])
def test_gemm_f32_signed_inputs_error_vs_abs_product(oracle, acc):
    """Signed fp32 inputs (sums cancel, so an error relative to the RESULT is meaningless): the error is measured against
    |A|.|B| (round-1 advisor finding).  Default mode: fp32-grade, within 4e-6 + 1.2e-7*k of |A||B| -- the same order as
    the reference's own sequential fp32 loop (k * 2^-24); opt-in TF32: 2 * 2^-11 per product."""
    m, k, n = 384, 1024, 512
    a = (oracle.fill((m, k), 11, np.float32) - np.float32(0.5)).astype(np.float32)
    b = (oracle.fill((k, n), 12, np.float32) - np.float32(0.5)).astype(np.float32)
    exact = a.astype(np.float64) @ b.astype(np.float64)
    absprod = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    ref = oracle.gemm(a, b)  # the reference's fp32 loop order
    ref_err = float(np.max(np.abs(ref.astype(np.float64) - exact) / absprod))
    c = np.empty((m, n), dtype=np.float32)
    check(lib().la_set_gemm_f32_mode(_cabi.LA_F32_TF32 if acc == "tf32" else _cabi.LA_F32_3XTF32))
    try:
        check(lib().la_gemm_f32_host(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
    finally:
        lib().la_set_gemm_f32_mode(_cabi.LA_F32_3XTF32)
    err = float(np.max(np.abs(c.astype(np.float64) - exact) / absprod))
    if acc == "tf32":
        assert err <= 1.1e-3
    else:
        assert err <= 4e-6 + 1.2e-7 * k
        assert err <= max(10 * ref_err, 4e-6), (err, ref_err)  # comparable with the reference's own rounding error


def test_gemm_f32_default_mode_is_fp32_grade():
    mode = __import__("ctypes").c_int(-1)
    check(lib().la_get_gemm_f32_mode(__import__("ctypes").byref(mode)))
    assert mode.value == _cabi.LA_F32_3XTF32


@pytest.mark.parametrize("acc", ["tf32", "3xtf32"])
def test_gemm_f32_full_size_config4_sampled_rows(oracle, acc):
    """BASELINE config 4 at full size: f32 65536 x 1024 times 1024 x 16384, inputs generated on the device, 64 sampled
    full rows against the fp32 oracle (rows of a product are independent, so the sample is exact): <= 1e-4 * k in the
    TF32 mode the config names, fp32-grade in the default mode; plus the column-sum checksum over a 4096-row band."""
    m, k, n = 65536, 1024, 16384
    da, db, dc = DevBuf(m * k * 4), DevBuf(k * n * 4), DevBuf(m * n * 4)
    fill_hash(da, m * k, 1, np.float32)
    fill_hash(db, k * n, 2, np.float32)
    check(lib().la_set_gemm_f32_mode(_cabi.LA_F32_TF32 if acc == "tf32" else _cabi.LA_F32_3XTF32))
    try:
        gemm_dev(da, k, db, n, dc, n, m, k, n, 0, np.float32)
        sync()
    finally:
        lib().la_set_gemm_f32_mode(_cabi.LA_F32_3XTF32)
    b = oracle.fill((k, n), 2, np.float32)
    rows = np.unique(np.concatenate([[0, 1, 127, 128, 8191, 8192, m - 1], np.random.default_rng(4).integers(0, m, 57)]))
    worst = 0.0
    for r in rows:
        a_row = oracle.fill((1, k), 1, np.float32, first_idx=int(r) * k)
        ref = oracle.gemm_rows(a_row, b, 0, 1)
        got = dc.to_array((1, n), np.float32, byte_offset=int(r) * n * 4)
        worst = max(worst, max_rel_err(got, ref))
    assert worst <= (F32_TOL * k if acc == "tf32" else 4e-6 + 1.2e-7 * k), worst
    band = dc.to_array((4096, n), np.float32, byte_offset=20480 * n * 4).astype(np.float64)
    a_band = oracle.fill((4096, k), 1, np.float32, first_idx=20480 * k).astype(np.float64)
    lhs = a_band.sum(axis=0) @ b.astype(np.float64)
    rhs = band.sum(axis=0)
    assert np.max(np.abs(lhs - rhs) / np.abs(lhs)) <= (1e-3 if acc == "tf32" else 1e-5)


def test_gemm_signed_inputs_absolute_error(oracle):
    """Inputs in [-0.5, 0.5): sums cancel, so compare against the norm-wise bound k * eps * |a|.|b| instead."""
    m, k, n = 192, 777 * 2, 320
    a = oracle.fill((m, k), 11) - 0.5
    b = oracle.fill((k, n), 12) - 0.5
    ref = oracle.gemm(a, b)
    c = np.empty((m, n))
    check(lib().la_gemm_f64_host(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
    bound = (np.abs(a) @ np.abs(b)) * (k * 2.3e-16)
    assert np.all(np.abs(c - ref) <= bound)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("path", ["simt", "tma"])
def test_gemm_dev_modes_and_leading_dims(oracle, mode, path):
    """Device-pointer entry with sub-matrix views (ld > width), as the LU trailing update uses it."""
    ld = 400
    m, k, n = 150, 96, 170
    big_a = oracle.fill((300, ld), 21)
    big_b = oracle.fill((300, ld), 22)
    big_c = oracle.fill((300, ld), 23)
    a_off, b_off, c_off = 10 * ld + 4, 20 * ld + 8, 30 * ld + 2
    a = big_a.ravel()[a_off:].reshape(-1)  # views via offsets
    A = np.lib.stride_tricks.as_strided(big_a.ravel()[a_off:], (m, k), (ld * 8, 8))
    B = np.lib.stride_tricks.as_strided(big_b.ravel()[b_off:], (k, n), (ld * 8, 8))
    C0 = np.lib.stride_tricks.as_strided(big_c.ravel()[c_off:], (m, n), (ld * 8, 8)).copy()
    prod = oracle.gemm(np.ascontiguousarray(A), np.ascontiguousarray(B))
    want = {0: prod, 1: C0 - prod, 2: C0 + prod}[mode]
    da, db, dc = DevBuf.from_array(big_a), DevBuf.from_array(big_b), DevBuf.from_array(big_c)
    check(lib().la_debug_set_gemm_path({"simt": 1, "tma": 2}[path]))
    try:
        gemm_dev(da, ld, db, ld, dc, ld, m, k, n, mode, np.float64, a_off, b_off, c_off)
        sync()
    finally:
        lib().la_debug_set_gemm_path(0)
    out = dc.to_array((300, ld), np.float64)
    got = np.lib.stride_tricks.as_strided(out.ravel()[c_off:], (m, n), (ld * 8, 8))
    scale = np.maximum(np.abs(want), np.abs(prod))
    assert np.max(np.abs(got - want) / scale) <= F64_TOL * k
    # nothing outside the C view was touched
    mask = np.ones((300, ld), dtype=bool)
    r0, c0 = divmod(c_off, ld)
    mask[r0:r0 + m, c0:c0 + n] = False
    assert np.array_equal(out[mask], big_c[mask])


def test_gemm_buf_api(oracle):
    m, k, n = 300, 200, 100
    a, b = oracle.fill((m, k), 1), oracle.fill((k, n), 2)
    da, db, dc = DevBuf.from_array(a), DevBuf.from_array(b), DevBuf(m * n * 8)
    check(lib().la_gemm_f64(da.h, db.h, dc.h, m, k, n))
    sync()
    assert max_rel_err(dc.to_array((m, n), np.float64), oracle.gemm(a, b)) <= F64_TOL * k
    # contract violations come back as LA_ERR_INVALID, not as garbage
    assert lib().la_gemm_f64(da.h, db.h, dc.h, m, k, n * 2) == _cabi.LA_ERR_INVALID
    assert lib().la_gemm_f64(da.h, db.h, da.h, m, k, n) == _cabi.LA_ERR_INVALID
    assert lib().la_gemm_f64(da.h, db.h, dc.h, 0, k, n) == _cabi.LA_ERR_INVALID


def test_gemm_linearity_property(oracle):
    """(A1 + A2) * B == A1*B + A2*B up to rounding: a size-independent check that needs no oracle."""
    m, k, n = 512, 640, 384
    a1, a2, b = oracle.fill((m, k), 31), oracle.fill((m, k), 32), oracle.fill((k, n), 33)
    out = []
    for a in (a1, a2, a1 + a2):
        c = np.empty((m, n))
        check(lib().la_gemm_f64_host(np.ascontiguousarray(a).ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
        out.append(c)
    assert max_rel_err(out[0] + out[1], out[2]) <= F64_TOL * k


def test_device_fill_matches_oracle(oracle):
    for dt in (np.float64, np.float32):
        n = 100003
        buf = DevBuf(n * np.dtype(dt).itemsize)
        fill_hash(buf, n, 7, dt, first_idx=12345)
        sync()
        assert np.array_equal(buf.to_array((n,), dt), oracle.fill((n,), 7, dt, first_idx=12345))


@pytest.mark.parametrize("n", [8192])
def test_gemm_f64_full_size_sampled_rows(oracle, n):
    """BASELINE config 1 (8192^3): inputs generated on the device, 64 full rows compared with the oracle (rows are
    independent, so the sample is exact), plus a checksum identity: (1^T A) B == 1^T C."""
    da, db, dc = DevBuf(n * n * 8), DevBuf(n * n * 8), DevBuf(n * n * 8)
    fill_hash(da, n * n, 1, np.float64)
    fill_hash(db, n * n, 2, np.float64)
    gemm_dev(da, n, db, n, dc, n, n, n, n, 0, np.float64)
    sync()
    b = oracle.fill((n, n), 2)
    rows = np.unique(np.concatenate([[0, 1, 127, 128, n - 1], np.random.default_rng(0).integers(0, n, 59)]))
    worst = 0.0
    for r in rows:
        a_row = oracle.fill((1, n), 1, first_idx=int(r) * n)
        ref = oracle.gemm_rows(a_row, b, 0, 1)
        got = dc.to_array((1, n), np.float64, byte_offset=int(r) * n * 8)
        worst = max(worst, max_rel_err(got, ref))
    assert worst <= F64_TOL * n
    # column-sum checksum over the whole product
    c = dc.to_array((n, n), np.float64)
    a_colsum = np.zeros(n)
    for r0 in range(0, n, 1024):
        a_colsum += oracle.fill((1024, n), 1, first_idx=r0 * n).sum(axis=0)
    lhs = a_colsum @ b
    rhs = c.sum(axis=0)
    assert np.max(np.abs(lhs - rhs) / np.abs(lhs)) <= 1e-11


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, F32_TOL)])
def test_gemm_host_two_phase_pipeline(oracle, dtype, tol):
    """Large products through the host-pointer entry run the two-phase pipeline (K-panels while B uploads, then row
    blocks while C downloads).  Ragged against every block size: k is not a multiple of the 1024-wide panels, m not of the
    row blocks.  Checked on sampled rows against the oracle (rows of a product are independent) and as a whole against the
    single-launch device-resident product."""
    m, k, n = 4224 + 40, 2304 + 8, 4100
    a = oracle.fill((m, k), 1, dtype)
    b = oracle.fill((k, n), 2, dtype)
    c = np.full((m, n), np.nan, dtype=dtype)
    fn = lib().la_gemm_f64_host if dtype == np.float64 else lib().la_gemm_f32_host
    check(fn(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
    assert np.all(np.isfinite(c))
    for r0 in (0, 1023, 2111, m - 3):
        ref = oracle.gemm_rows(a, b, r0, r0 + 3)
        assert max_rel_err(c[r0:r0 + 3], ref) <= tol * k
    da, db, dc = DevBuf.from_array(a), DevBuf.from_array(b), DevBuf(m * n * a.itemsize)
    gemm_dev(da, k, db, n, dc, n, m, k, n, 0, dtype)
    sync()
    whole = dc.to_array((m, n), dtype)
    assert max_rel_err(c, whole) <= (1e-13 if dtype == np.float64 else 2e-3)
