"""CholeskyDecomposition on the CUDA path against the oracle and the reference's own unit
tests (src/decomp/cholesky.rs:146-183)."""
import numpy as np
import pytest

from la import CholeskyDecomposition, Matrix, Panic, m
from la._cabi import check, lib

pytestmark = pytest.mark.gpu


def spd(oracle, n, seed, dtype=np.float64):
    g = oracle.fill((n, n), seed, dtype)
    a = oracle.gemm(g, np.ascontiguousarray(g.T))  # exactly symmetric (same products, same order, for [i][j] and [j][i])
    return a + dtype(n) * np.eye(n, dtype=dtype)


def test_reference_unit_tests():
    a = m("4.0, 12.0, -16.0; 12.0, 37.0, -43.0; -16.0, -43.0, 98.0")
    c = CholeskyDecomposition.new(a)
    assert c.get_l() * c.get_l().t() == a                                                        # cholesky.rs:149
    assert list(c.get_l().get_data()) == [2.0, 0.0, 0.0, 6.0, 1.0, 0.0, -8.0, 5.0, 3.0]          # :150
    assert CholeskyDecomposition.new(m("4.0, 12.0, -16.0; 12.0, 37.0, 43.0; -16.0, 43.0, 98.0")) is None   # :154-157
    assert CholeskyDecomposition.new(m("4.0, 12.0, -16.0; 12.0, 37.0, 43.0")) is None            # :160-163
    a = m("2.0, 1.0, 0.0; 1.0, 1.0, 0.0; 0.0, 0.0, 1.0")
    c = CholeskyDecomposition.new(a)
    assert c.solve(m("1.0; 2.0; 3.0")).approx_eq(m("-1.0; 3.0; 3.0"))                            # :166-171
    with pytest.raises(Panic):
        c.solve(m("1.0; 2.0; 3.0; 4.0"))                                                         # :173-180


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 17, 100, 128])
def test_single_block_is_bit_identical(oracle, dtype, n):
    a = spd(oracle, n, 3, dtype)
    ref = oracle.chol(a)
    got = CholeskyDecomposition.new(Matrix.from_numpy(a)).get_l().to_numpy()
    assert np.array_equal(got.view(np.uint8), ref.view(np.uint8))


@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-4)])
@pytest.mark.parametrize("n", [129, 300, 640, 1154, 2048])
def test_blocked_factor_parity(oracle, dtype, tol, n):
    a = spd(oracle, n, 5, dtype)
    ref = oracle.chol(a)
    c = CholeskyDecomposition.new(Matrix.from_numpy(a))
    assert c is not None
    l = c.get_l().to_numpy()
    assert np.all(np.triu(l, 1) == 0)  # explicit zeros above the diagonal (cholesky.rs:107-109)
    assert np.max(np.abs(l - ref)) / np.max(np.abs(ref)) <= tol * n
    l64, a64, r64 = l.astype(np.float64), a.astype(np.float64), ref.astype(np.float64)
    err = np.linalg.norm(l64 @ l64.T - a64) / np.linalg.norm(a64)
    err_ref = np.linalg.norm(r64 @ r64.T - a64) / np.linalg.norm(a64)
    assert err <= 10 * max(err_ref, np.finfo(dtype).eps)  # backward error within 10x of the reference's


def test_none_cases_in_large_matrices(oracle):
    n = 700
    a = spd(oracle, n, 6)
    b = a.copy()
    b[650, 3] = np.nextafter(b[650, 3], np.inf)  # one asymmetric pair, far from the first block
    assert oracle.chol(b) is None and CholeskyDecomposition.new(Matrix.from_numpy(b)) is None
    c = a.copy()
    c[690, 690] = -c[690, 690]  # not positive definite, detected in the last block
    assert oracle.chol(c) is None and CholeskyDecomposition.new(Matrix.from_numpy(c)) is None
    d = a.copy()
    d[5, 9] = d[9, 5] = np.nan  # NaN != NaN: "not symmetric"
    assert oracle.chol(d) is None and CholeskyDecomposition.new(Matrix.from_numpy(d)) is None
    assert CholeskyDecomposition.new(Matrix.from_numpy(a)) is not None


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,nx", [(3, 1), (64, 5), (300, 16), (511, 2)])
def test_solve_given_reference_factor_is_bit_identical(oracle, dtype, n, nx):
    """Feed the ORACLE's L to the CUDA solve: the reference-order kernel (cholesky.rs:116-144) is exact."""
    a = spd(oracle, n, 7, dtype)
    ref_l = oracle.chol(a)
    b = oracle.fill((n, nx), 8, dtype)
    want = oracle.chol_solve(ref_l, b)
    x = np.empty_like(b)
    fn = lib().la_chol_solve_f64_host if dtype == np.float64 else lib().la_chol_solve_f32_host
    check(fn(ref_l.ctypes.data, n, b.ctypes.data, nx, x.ctypes.data))
    assert np.array_equal(x.view(np.uint8), want.view(np.uint8))


@pytest.mark.parametrize("n,nx", [(512, 16), (1154, 2), (2048, 64)])
def test_solve_gemm_sweeps(oracle, n, nx):
    a = spd(oracle, n, 9)
    b = oracle.fill((n, nx), 10)
    ref_l = oracle.chol(a)
    want = oracle.chol_solve(ref_l, b)
    c = CholeskyDecomposition.new(Matrix.from_numpy(a))
    x = c.solve(Matrix.from_numpy(b)).to_numpy()
    r = np.linalg.norm(a @ x - b) / (np.linalg.norm(a) * np.linalg.norm(x))
    r_ref = np.linalg.norm(a @ want - b) / (np.linalg.norm(a) * np.linalg.norm(want))
    assert r <= 10 * max(r_ref, np.finfo(np.float64).eps)
    assert np.max(np.abs(x - want)) / np.max(np.abs(want)) <= 1e-12 * n * max(1.0, np.linalg.cond(a) / n)
