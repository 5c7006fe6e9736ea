"""The Cholesky oracle (oracle/la_oracle.c, DEFINE_CHOL) against the reference's own unit tests
(src/decomp/cholesky.rs:146-183) and its two forms against each other."""
import numpy as np
import pytest


def spd(oracle, n, seed, dtype=np.float64):
    m = oracle.fill((n, n), seed, dtype)
    a = oracle.gemm(m, np.ascontiguousarray(m.T))  # exactly symmetric: (m m')[i][j] and [j][i] sum the same products in order
    return a + dtype(n) * np.eye(n, dtype=dtype)


def test_reference_square_pos_def(oracle):
    a = np.array([[4.0, 12.0, -16.0], [12.0, 37.0, -43.0], [-16.0, -43.0, 98.0]])
    for form in ("canon", "fast"):
        l = oracle.chol(a, form)
        assert np.array_equal(l.reshape(-1), [2.0, 0.0, 0.0, 6.0, 1.0, 0.0, -8.0, 5.0, 3.0])  # cholesky.rs:150
        assert np.array_equal(oracle.gemm(l, np.ascontiguousarray(l.T)), a)                   # cholesky.rs:149


def test_reference_none_cases(oracle):
    not_pd = np.array([[4.0, 12.0, -16.0], [12.0, 37.0, 43.0], [-16.0, 43.0, 98.0]])          # cholesky.rs:154-157
    not_square = np.array([[4.0, 12.0, -16.0], [12.0, 37.0, 43.0]])                           # cholesky.rs:160-163
    not_sym = np.array([[4.0, 1.0], [2.0, 5.0]])
    nan_pair = np.array([[4.0, np.nan], [np.nan, 5.0]])  # NaN != NaN: "not symmetric"
    for form in ("canon", "fast"):
        assert oracle.chol(not_pd, form) is None
        assert oracle.chol(not_square, form) is None
        assert oracle.chol(not_sym, form) is None
        assert oracle.chol(nan_pair, form) is None


def test_reference_solve(oracle):
    a = np.array([[2.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    x = oracle.chol_solve(oracle.chol(a), np.array([[1.0], [2.0], [3.0]]))
    assert np.max(np.abs(x - np.array([[-1.0], [3.0], [3.0]]))) < 1e-6                        # cholesky.rs:166-171 (approx_eq)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 2, 17, 130, 300])
def test_forms_bit_identical(oracle, dtype, n):
    a = spd(oracle, n, 3, dtype)
    lc, lf = oracle.chol(a, "canon"), oracle.chol(a, "fast")
    assert lc is not None and np.array_equal(lc.view(np.uint8), lf.view(np.uint8))
    rel = np.max(np.abs(lc.astype(np.float64) @ lc.astype(np.float64).T - a)) / np.max(np.abs(a))
    assert rel <= (1e-13 if dtype == np.float64 else 1e-5) * n
    b = oracle.fill((n, 3), 4, dtype)
    x = oracle.chol_solve(lc, b)
    assert np.max(np.abs(a.astype(np.float64) @ x - b)) <= (1e-11 if dtype == np.float64 else 1e-3) * n
