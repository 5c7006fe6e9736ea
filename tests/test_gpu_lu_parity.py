"""LU parity: blocked CUDA LU (C ABI) vs the oracle's restatement of src/decomp/lu.rs on the same seeded inputs.
Bars (BASELINE.json): identical pivot permutation; element error <= 1e-12*n (f64) / 1e-4*n (f32) relative to
max(|ref|, max|A|); backward error ||PA-LU||/||A|| within 10x of the reference's."""
import ctypes

import numpy as np
import pytest

from gpu_util import DevBuf, fill_hash, sync
from la import LUDecomposition, Matrix, _cabi
from la._cabi import check, lib

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (2, 2), (5, 5), (16, 16), (33, 33), (127, 127), (128, 128), (129, 129), (130, 130), (200, 200),
          (256, 256), (300, 300), (512, 512), (1000, 1000), (1024, 1024),
          (300, 100), (100, 300), (129, 1), (1, 129), (1000, 260), (260, 1000), (513, 257), (2, 400), (400, 2)]


def factor_host(a):
    m, n = a.shape
    suf = "f64" if a.dtype == np.float64 else "f32"
    lu = np.empty_like(a)
    piv = np.empty(m, dtype=np.uint64)
    sign = ctypes.c_int(-7)
    check(getattr(lib(), f"la_lu_factor_{suf}_host")(a.ctypes.data, lu.ctypes.data, m, n, piv.ctypes.data,
                                                     ctypes.byref(sign)))
    return lu, piv, bool(sign.value)


def check_lu(oracle, a, tol_unit):
    m, n = a.shape
    ref_lu, ref_piv, ref_sign = oracle.lu(a)
    lu, piv, sign = factor_host(a)
    assert np.array_equal(piv, ref_piv), f"pivot mismatch at {np.nonzero(piv != ref_piv)[0][:5]}"
    assert sign == ref_sign
    amax = float(np.max(np.abs(a)))
    den = np.maximum(np.abs(ref_lu.astype(np.float64)), amax)
    err = float(np.max(np.abs(lu.astype(np.float64) - ref_lu.astype(np.float64)) / den))
    assert err <= tol_unit * max(m, n), f"element error {err}"
    be_ref = oracle.lu_backward_error(a, ref_lu, ref_piv)
    be = oracle.lu_backward_error(a, lu, piv)
    eps = np.finfo(a.dtype).eps
    assert be <= 10 * max(be_ref, eps), f"backward error {be} vs reference {be_ref}"
    return lu, piv, sign


@pytest.mark.parametrize("shape", SHAPES, ids=[f"{m}x{n}" for m, n in SHAPES])
def test_lu_f64_parity(oracle, shape):
    a = oracle.fill(shape, 1)
    check_lu(oracle, a, 1e-12)


@pytest.mark.parametrize("shape", [(5, 5), (130, 130), (300, 300), (200, 77), (77, 200), (640, 640)])
def test_lu_f64_signed_entries(oracle, shape):
    a = oracle.fill(shape, 41) - 0.5
    check_lu(oracle, a, 1e-12)


@pytest.mark.parametrize("shape", [(3, 3), (64, 64), (129, 129), (300, 200), (200, 300), (512, 512)])
def test_lu_f32_parity(oracle, shape):
    a = oracle.fill(shape, 1, np.float32)
    check_lu(oracle, a, 1e-4)


@pytest.mark.parametrize("shape,dtype", [((8000, 128), np.float64), ((16384, 64), np.float64), ((20000, 100), np.float64),
                                         ((8000, 128), np.float32), ((20000, 100), np.float32), ((7000, 96), np.float64)])
def test_lu_tall_skinny_single_panel(oracle, shape, dtype):
    """Tall single-panel matrices: the bit-exact panel keeps two shared-memory arrays per row, so the rows per CTA must be
    clamped to the shared-memory budget (round-1 advisor finding: 8000 x 128 asked for 258 KB and failed to launch)."""
    a = oracle.fill(shape, 9, dtype)
    check_lu(oracle, a, 1e-12 if dtype == np.float64 else 1e-4)


def test_lu_input_untouched_and_odd_sizes(oracle):
    a = oracle.fill((333, 333), 2)
    keep = a.copy()
    check_lu(oracle, a, 1e-12)
    assert np.array_equal(a, keep)


def test_lu_ties_pick_lowest_row(oracle):
    """lu.rs:132-137 strict '>': among equal |x| the first row wins.  All-ones columns and sign-mixed ties."""
    n = 150
    a = np.ones((n, n))
    a += np.triu(np.arange(n * n, dtype=np.float64).reshape(n, n) % 7, 1)
    check_lu(oracle, a, 1e-12)
    b = oracle.fill((200, 200), 3)
    b[:, 0] = np.where(np.arange(200) % 2 == 0, 0.75, -0.75)
    check_lu(oracle, b, 1e-12)


def test_lu_zero_pivot_continues(oracle):
    """lu.rs:156-160: exact zero pivot => no division, factorisation continues, solve -> None."""
    a = oracle.fill((140, 140), 5)
    a[:, 3] = 0.0          # column 3 is identically zero => zero pivot at step 3
    a[:, 131] = 0.0        # and one inside the second panel
    lu, piv, sign = check_lu(oracle, a, 1e-12)
    assert np.all(np.isfinite(lu))
    A = Matrix.from_numpy(a)
    dec = LUDecomposition.new(A)
    assert dec.is_singular()
    assert dec.solve(Matrix.from_numpy(np.ones((140, 1)))) is None
    assert dec.det() == 0.0 or dec.det() == -0.0


def test_lu_nan_semantics(oracle):
    """NaN never displaces the incumbent, a NaN incumbent is never displaced (lu.rs:132-137)."""
    a = oracle.fill((40, 40), 6)
    a[17, 0] = np.nan       # candidate NaN in column 0: must not be chosen
    ref_lu, ref_piv, ref_sign = oracle.lu(a)
    lu, piv, sign = factor_host(a)
    assert np.array_equal(piv, ref_piv) and sign == ref_sign
    assert np.array_equal(np.isnan(lu), np.isnan(ref_lu))
    b = oracle.fill((40, 40), 7)
    b[0, 0] = np.nan        # incumbent NaN: stays
    ref_lu, ref_piv, ref_sign = oracle.lu(b)
    lu, piv, sign = factor_host(b)
    assert np.array_equal(piv, ref_piv) and sign == ref_sign


@pytest.mark.parametrize("n,nx", [(1, 1), (7, 3), (64, 1), (65, 16), (200, 5), (512, 1), (512, 16), (1000, 33)])
def test_solve_parity(oracle, n, nx):
    a = oracle.fill((n, n), 1)
    b = oracle.fill((n, nx), 3)
    ref_lu, ref_piv, _ = oracle.lu(a)
    ref_x = oracle.lu_solve(ref_lu, ref_piv, b)
    dec = LUDecomposition.new(Matrix.from_numpy(a))
    x = dec.solve(Matrix.from_numpy(b)).to_numpy()
    # forward-error comparison is conditioning-limited: compare residuals (||Ax-b||) within 10x of the reference's,
    # and the solutions to a condition-scaled tolerance
    r_ref = np.linalg.norm(a @ ref_x - b) / (np.linalg.norm(a) * np.linalg.norm(ref_x))
    r = np.linalg.norm(a @ x - b) / (np.linalg.norm(a) * np.linalg.norm(x))
    assert r <= 10 * max(r_ref, np.finfo(np.float64).eps)
    cond = np.linalg.cond(a)
    assert np.max(np.abs(x - ref_x)) / np.max(np.abs(ref_x)) <= 1e-12 * n * max(1.0, cond / n)


@pytest.mark.parametrize("n,nx", [(64, 3), (300, 16), (511, 16), (510, 16)])
def test_solve_given_reference_factors_matches_oracle(oracle, n, nx):
    """Feed the ORACLE's packed LU to the CUDA solve: isolates the triangular sweeps (lu.rs:257-275).  Below 512 rows
    the launch-per-block path runs: same order, separately rounded ops -> bit-exact."""
    a = oracle.fill((n, n), 1)
    b = oracle.fill((n, nx), 3)
    ref_lu, ref_piv, _ = oracle.lu(a)
    ref_x = oracle.lu_solve(ref_lu, ref_piv, b)
    x = np.empty((n, nx))
    check(lib().la_lu_solve_f64_host(ref_lu.ctypes.data, n, n, ref_piv.ctypes.data, b.ctypes.data, nx, x.ctypes.data))
    assert np.array_equal(x.view(np.uint64), ref_x.view(np.uint64))


@pytest.mark.parametrize("n,nx", [(512, 16), (640, 7), (1154, 16), (2048, 1), (2306, 9), (4096, 16),
                                  (640, 32), (1154, 100), (2048, 2048), (2306, 18)])
def test_solve_sweep_given_reference_factors(oracle, n, nx):
    """n >= 512, even, nx <= 16: the persistent sweep kernels (ragged last block included).  Same elimination order
    with fused multiply-adds and a reciprocal diagonal.  nx > 16 (even): inverted 128 x 128 diagonal blocks + DMMA GEMMs
    (the path `inverse()` takes).  Either way the solution must satisfy the SAME triangular systems as the
    oracle's to rounding -- component-wise backward error of L*U*x = b(piv) within 4x the oracle's own, and the
    difference to the oracle's x within the 1e-12*n bar scaled by the conditioning."""
    a = oracle.fill((n, n), 1)
    b = oracle.fill((n, nx), 3)
    ref_lu, ref_piv, _ = oracle.lu(a)
    ref_x = oracle.lu_solve(ref_lu, ref_piv, b)
    x = np.empty((n, nx))
    check(lib().la_lu_solve_f64_host(ref_lu.ctypes.data, n, n, ref_piv.ctypes.data, b.ctypes.data, nx, x.ctypes.data))
    l = np.tril(ref_lu, -1) + np.eye(n)
    u = np.triu(ref_lu)
    bp = b[ref_piv.astype(np.int64)]

    def backward_error(sol):
        return np.max(np.abs(l @ (u @ sol) - bp) / (np.abs(l) @ (np.abs(u) @ np.abs(sol)) + np.abs(bp)))

    assert backward_error(x) <= 4 * max(backward_error(ref_x), np.finfo(np.float64).eps)
    cond = np.linalg.cond(a)
    assert np.max(np.abs(x - ref_x)) / np.max(np.abs(ref_x)) <= 1e-12 * n * max(1.0, cond / n)


def test_det_parity_and_overflow_order(oracle):
    a = oracle.fill((64, 64), 1)
    ref_lu, ref_piv, ref_sign = oracle.lu(a)
    d_ref = oracle.lu_det(ref_lu, ref_sign)
    d = Matrix.from_numpy(a).det()
    assert np.sign(d) == np.sign(d_ref) and abs(d - d_ref) <= 1e-12 * 64 * abs(d_ref)
    # sequential product semantics (lu.rs:226-231): overflow to inf first, then * 0 -> NaN, exactly as in order
    n = 400
    diag = np.full(n, 1e300)
    diag[-1] = 0.0
    lu = np.diag(diag)
    buf = DevBuf.from_array(lu)
    out = ctypes.c_double(0)
    check(lib().la_lu_det_f64(buf.h, n, 1, ctypes.byref(out)))
    assert np.isnan(out.value) and np.isnan(oracle.lu_det(lu, True))
    diag2 = np.full(n, 0.5)
    lu2 = np.diag(diag2)
    buf2 = DevBuf.from_array(lu2)
    check(lib().la_lu_det_f64(buf2.h, n, 0, ctypes.byref(out)))
    assert out.value == oracle.lu_det(lu2, False)


def test_inverse_roundtrip(oracle):
    """A * A^-1 ~ I, judged against the reference's own residual (seed 9 at n = 257 is ill-conditioned: cond ~ 1.7e8)."""
    for n in (130, 256, 257):
        a = oracle.fill((n, n), 9)
        A = Matrix.from_numpy(a)
        inv = A.inverse()
        assert inv is not None
        ref_lu, ref_piv, _ = oracle.lu(a)
        ref_inv = oracle.lu_solve(ref_lu, ref_piv, oracle.identity(n))
        r_ref = np.max(np.abs(a @ ref_inv - np.eye(n)))
        r = np.max(np.abs((A * inv).to_numpy() - np.eye(n)))
        assert r <= 10 * max(r_ref, 1e-13), (n, r, r_ref)


def test_lu_reconstruction_property_2048(oracle):
    """P*A == L*U through the CUDA GEMM itself, at a size where the oracle is still cheap (pivot identity checked too)."""
    n = 2048
    a = oracle.fill((n, n), 1)
    ref_lu, ref_piv, _ = oracle.lu(a)
    lu, piv, sign = factor_host(a)
    assert np.array_equal(piv, ref_piv)
    l = np.tril(lu, -1) + np.eye(n)
    u = np.triu(lu)
    rec = (Matrix.from_numpy(l) * Matrix.from_numpy(u)).to_numpy()
    assert np.linalg.norm(rec - a[piv.astype(np.int64)]) / np.linalg.norm(a) <= 1e-13
    amax = 1.0
    err = np.max(np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), amax))
    assert err <= 1e-12 * n


@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (17, 17), (64, 64), (100, 100), (128, 128), (90, 40), (40, 90), (128, 300)])
def test_single_panel_lu_is_bit_exact(oracle, shape):
    """min(m,n) <= 128: the deferred-subtraction panel + TRSM reproduce the reference bit for bit (incl. exact zeros)."""
    for seed, shift in ((1, 0.0), (8, 0.5)):
        a = oracle.fill(shape, seed) - shift
        ref_lu, ref_piv, ref_sign = oracle.lu(a, form="canon")
        lu, piv, sign = factor_host(a)
        assert np.array_equal(piv, ref_piv) and sign == ref_sign
        assert np.array_equal(lu.view(np.uint64), ref_lu.view(np.uint64))
    # an exactly singular integer-valued matrix keeps its exact zero pivot (is_singular parity with the reference)
    m, n = shape
    if m == n and m >= 3:
        s = np.round(oracle.fill(shape, 5) * 8)
        s[2] = s[0] + s[1]
        ref_lu, ref_piv, _ = oracle.lu(s, form="canon")
        lu, piv, _ = factor_host(s)
        assert np.array_equal(lu.view(np.uint64), ref_lu.view(np.uint64))


@pytest.mark.parametrize("shape", [(5, 7, 3), (64, 64, 64), (100, 130, 90), (128, 128, 128)])
def test_small_gemm_is_bit_exact(oracle, shape):
    """Small products run the reference-order CUDA-core kernel: bit-identical to src/matrix/mod.rs:965-973."""
    m, k, n = shape
    for dt in (np.float64, np.float32):
        a = oracle.fill((m, k), 1, dt) - dt(0.5)
        b = oracle.fill((k, n), 2, dt) - dt(0.5)
        c = np.empty((m, n), dtype=dt)
        suf = "f64" if dt == np.float64 else "f32"
        check(getattr(lib(), f"la_gemm_{suf}_host")(a.ctypes.data, b.ctypes.data, c.ctypes.data, m, k, n))
        assert np.array_equal(c.view(np.uint8), oracle.gemm(a, b, form="canon").view(np.uint8))


@pytest.mark.parametrize("cond", [1e4, 1e8, 1e12])
def test_solve_residual_on_ill_conditioned_systems(oracle, cond):
    """The fast solves multiply by explicitly inverted 128 x 128 diagonal blocks instead of substituting (lu_solve.cu).
    On ill-conditioned systems (prescribed singular values 1 .. 1/cond, and a row-graded matrix) the scaled residual must
    stay within 10x of the reference's substitution (lu.rs:255-275) for both the sweep kernels (nx <= 16) and the GEMM sweeps
    (nx > 16); the solutions themselves may differ by cond * eps."""
    rng = np.random.default_rng(7)
    n = 1024
    q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
    q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    mats = [np.ascontiguousarray(q1 @ np.diag(np.logspace(0, -np.log10(cond), n)) @ q2)]
    if cond == 1e8:
        mats.append(np.ascontiguousarray(rng.standard_normal((n, n)) * np.logspace(0, -8, n)[:, None]))

    def res(a, x, b):
        return np.linalg.norm(a @ x - b) / (np.linalg.norm(a) * np.linalg.norm(x))

    for a in mats:
        lu, piv, _ = oracle.lu(a)
        dec = LUDecomposition.new(Matrix.from_numpy(a))
        for nx in (3, 40):
            b = np.ascontiguousarray(rng.standard_normal((n, nx)))
            xr = oracle.lu_solve(lu, piv, b)
            xg = dec.solve(Matrix.from_numpy(b)).to_numpy()
            assert np.all(np.isfinite(xg))
            assert res(a, xg, b) <= 10 * max(res(a, xr, b), 1e-16)
            assert np.linalg.norm(xg - xr) <= 1e3 * cond * np.finfo(np.float64).eps * np.linalg.norm(xr)
