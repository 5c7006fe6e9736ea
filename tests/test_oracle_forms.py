"""The order-preserving fast forms of the oracle must be bit-identical to the canonical loop nests."""
import numpy as np
import pytest


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(64, 64, 64), (129, 67, 250), (1, 300, 17), (257, 1, 31), (384, 384, 384)])
def test_gemm_fast_equals_canon(oracle, dtype, shape):
    m, k, n = shape
    a = oracle.fill((m, k), 1, dtype)
    b = oracle.fill((k, n), 2, dtype)
    c0 = oracle.gemm(a, b, form="canon")
    c1 = oracle.gemm(a, b, form="fast")
    assert np.array_equal(c0.view(np.uint8), c1.view(np.uint8))
    rows = oracle.gemm_rows(a, b, m // 3, m // 3 + max(1, m // 4), form="canon", threads=2)
    assert np.array_equal(rows, c0[m // 3: m // 3 + max(1, m // 4)])
    rows = oracle.gemm_rows(a, b, 0, 1, form="fast")
    assert np.array_equal(rows, c0[:1])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (37, 37), (200, 200), (130, 75), (75, 130), (513, 513)])
def test_lu_fast_equals_canon(oracle, dtype, shape):
    m, n = shape
    a = oracle.fill((m, n), 5, dtype) - dtype(0.5)
    p0, piv0, s0 = oracle.lu(a, form="canon")
    p1, piv1, s1 = oracle.lu(a, form="fast")
    assert np.array_equal(p0.view(np.uint8), p1.view(np.uint8))
    assert np.array_equal(piv0, piv1) and s0 == s1
    if m == n:
        b = oracle.fill((m, 5), 3, dtype)
        x0 = oracle.lu_solve(p0, piv0, b, form="canon")
        x1 = oracle.lu_solve(p0, piv0, b, form="fast")
        assert np.array_equal(x0.view(np.uint8), x1.view(np.uint8))


def test_lu_zero_pivot_continues(oracle):
    """lu.rs:156-160: zero pivot => division skipped, factorisation continues, solve -> None."""
    a = np.array([[0.0, 0.0, 1.0], [0.0, 0.0, 2.0], [0.0, 3.0, 4.0]])
    p, piv, pos = oracle.lu(a, form="canon")
    assert np.all(np.isfinite(p))
    assert not oracle.lu_is_non_singular(p)
    assert oracle.lu_solve(p, piv, np.ones((3, 1))) is None
    p2, piv2, pos2 = oracle.lu(a, form="fast")
    assert np.array_equal(p, p2) and np.array_equal(piv, piv2) and pos == pos2


def test_lu_nan_never_wins_pivot(oracle):
    """lu.rs:132-137: strict '>' so a NaN candidate is never selected over the current pivot."""
    a = np.array([[1.0, 2.0], [np.nan, 3.0]])
    _, piv, _ = oracle.lu(a, form="canon")
    assert piv.tolist() == [0, 1]


def test_fill_is_counter_based(oracle):
    a = oracle.fill((10, 10), 1)
    b = oracle.fill((5, 10), 1, first_idx=50)
    assert np.array_equal(a[5:], b)
    assert a.min() >= 0.0 and a.max() < 1.0
    f = oracle.fill((1000,), 9, np.float32)
    assert f.min() >= 0.0 and f.max() < 1.0
    # known-answer: splitmix64 finaliser of (seed*golden + idx)
    z = (1 * 0x9E3779B97F4A7C15 + 0) & (2**64 - 1)
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
    z ^= z >> 31
    assert a[0, 0] == (z >> 11) * 2.0**-53
