"""The QR oracle (oracle/la_oracle.c, DEFINE_QR) against the reference's own unit tests (src/decomp/qr.rs:241-262: three
inputs, `(q * r).approx_eq(a)`), against the independent pure-Python restatement (oracle/pyref.py, bit for bit) and its two
forms against each other; pinverse's test (src/matrix/mod.rs:1549-1552) through the oracle chain."""
import numpy as np
import pytest

from oracle import pyref

REF_INPUTS = [
    ("qr_test", np.array([[12.0, -51.0, 4.0], [6.0, 167.0, -68.0], [-4.0, 24.0, -41.0]])),   # qr.rs:243
    ("qr_test_m_over_n", np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])),                    # qr.rs:250
    ("qr_test_n_over_m", np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])),                      # qr.rs:257
]


@pytest.mark.parametrize("name,a", REF_INPUTS)
def test_reference_q_times_r(oracle, name, a):
    m, n = a.shape
    for form in ("canon", "fast"):
        packed, rdiag = oracle.qr(a, form)
        q, r = oracle.qr_get_q(packed, rdiag), oracle.qr_get_r(packed, rdiag)
        assert q.shape == (m, m) and r.shape == (m, n)
        assert np.all(np.abs(oracle.gemm(q, r) - a) < 1e-6)  # approx_eq, src/approxeq.rs:36
        # independent restatement, bit for bit
        pq, prd = pyref.qr_new(a.reshape(-1).tolist(), m, n)
        assert np.array_equal(packed.reshape(-1), np.array(pq)) and np.array_equal(rdiag, np.array(prd))
        assert np.array_equal(q.reshape(-1), np.array(pyref.qr_get_q(pq, prd, m, n)))
        assert np.array_equal(r.reshape(-1), np.array(pyref.qr_get_r(pq, prd, m, n)))


def test_known_values_of_the_first_reference_input(oracle):
    """Hand-checkable: the first column of qr_test's input has norm 14, so rdiag[0] = -14 and u_0 = (26, 6, -4)."""
    packed, rdiag = oracle.qr(REF_INPUTS[0][1])
    assert np.array_equal(rdiag, [-14.0, -175.0, 35.0])
    assert np.array_equal(packed[:, 0], [26.0, 6.0, -4.0])


def test_solve_quirks(oracle):
    a = REF_INPUTS[0][1]
    packed, rdiag = oracle.qr(a)
    b = np.array([[1.0], [2.0], [3.0]])
    x = oracle.qr_solve(packed, rdiag, b)
    ref = pyref.qr_solve(packed.reshape(-1).tolist(), rdiag.tolist(), 3, 3, b.reshape(-1).tolist(), 1)
    assert np.array_equal(x.reshape(-1), np.array(ref))
    # the first phase applies I - u u'/u_k, not the reflection: the result is NOT the solution of a x = b (parity, no fix)
    assert np.max(np.abs(a @ x - b)) > 1.0
    # m > n: Matrix::new(cols, nx, <m * nx values>) panics (qr.rs:237, mod.rs:208)
    p2, r2 = oracle.qr(REF_INPUTS[1][1])
    with pytest.raises(AssertionError):
        oracle.qr_solve(p2, r2, np.ones((3, 1)))
    # m < n: is_full_rank indexes rdiag out of bounds (qr.rs:112)
    p3, r3 = oracle.qr(REF_INPUTS[2][1])
    with pytest.raises(IndexError):
        oracle.qr_solve(p3, r3, np.ones((2, 1)))
    # rank deficient: None
    pz, rz = oracle.qr(np.zeros((2, 2)))
    assert np.array_equal(rz, [0.0, 0.0]) and oracle.qr_solve(pz, rz, np.ones((2, 1))) is None


def test_pinverse_reference_test(oracle):
    """src/matrix/mod.rs:1549-1552: (a.pinverse() * a).approx_eq(id(2, 2)) for a 3 x 2 input, via r = get_r."""
    a = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    packed, rdiag = oracle.qr(a)
    r = oracle.qr_get_r(packed, rdiag)
    rtr = oracle.gemm(np.ascontiguousarray(r.T), r)
    lu, piv, _ = oracle.lu(rtr)
    inv = oracle.lu_solve(lu, piv, np.eye(2))
    pinv = oracle.gemm(inv, np.ascontiguousarray(a.T))
    assert np.all(np.abs(oracle.gemm(pinv, a) - np.eye(2)) < 1e-6)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (5, 5), (40, 17), (17, 40), (130, 130), (200, 70)])
def test_forms_bit_identical(oracle, dtype, shape):
    a = oracle.fill(shape, 7, dtype) - dtype(0.5)
    pc, rc = oracle.qr(a, "canon")
    pf, rf = oracle.qr(a, "fast")
    assert np.array_equal(pc.view(np.uint8), pf.view(np.uint8)) and np.array_equal(rc.view(np.uint8), rf.view(np.uint8))
    q, r = oracle.qr_get_q(pc, rc), oracle.qr_get_r(pc, rc)
    err = np.max(np.abs(q.astype(np.float64) @ r.astype(np.float64) - a))
    assert err <= (1e-13 if dtype == np.float64 else 1e-5) * max(shape)
    k = min(shape)
    qq = q.astype(np.float64)[:, :k]
    assert np.max(np.abs(qq.T @ qq - np.eye(k))) <= (1e-13 if dtype == np.float64 else 1e-5) * max(shape)
