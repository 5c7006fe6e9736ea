"""QRDecomposition on the CUDA path (csrc/qr.cu through la_qr_*) against the oracle: the reference's
own three tests and pinverse test replayed through the mirror, packed qr / rdiag / Q / R parity on random shapes, the
reference's solve (quirks included) and the device-resident pinverse chain."""
import numpy as np
import pytest

import la
from la import DeviceMatrix, Matrix, QRDecomposition

pytestmark = pytest.mark.gpu

REF_INPUTS = [
    np.array([[12.0, -51.0, 4.0], [6.0, 167.0, -68.0], [-4.0, 24.0, -41.0]]),   # qr.rs:243
    np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]]),                             # qr.rs:250
    np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]),                               # qr.rs:257
]


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_reference_tests(oracle, idx):
    a = Matrix.from_numpy(REF_INPUTS[idx])
    qr = QRDecomposition.new(a)
    assert (qr.get_q() * qr.get_r()).approx_eq(a)  # qr.rs:245, :252, :259
    packed, rdiag = oracle.qr(REF_INPUTS[idx])
    assert np.max(np.abs(qr.get_qr().to_numpy() - packed)) <= 1e-12 * np.max(np.abs(packed))
    assert np.max(np.abs(qr.rdiag - rdiag)) <= 1e-12 * np.max(np.abs(rdiag))


def test_reference_pinverse():
    a = la.m("1.0, 2.0; 3.0, 4.0; 5.0, 6.0")
    assert (a.pinverse() * a).approx_eq(Matrix.id(2, 2))  # src/matrix/mod.rs:1549-1552


def _parity(oracle, a, tol):
    m, n = a.shape
    packed, rdiag = oracle.qr(a)
    qr = QRDecomposition.new(Matrix.from_numpy(a))
    got = qr.get_qr().to_numpy().astype(np.float64)
    scale = max(np.max(np.abs(packed)), 1.0)
    # The sign of a reflection is decided by `x_kk > 0` (qr.rs:58) on the partially reduced diagonal entry x_kk = u_kk + a.
    # Where |x_kk| is at rounding level relative to the column norm |a| either sign is a legitimate outcome of a different
    # summation order; everything downstream then differs by sign, so element-wise parity only applies when all decisions agree.
    k = min(m, n)
    xkk = packed[np.arange(k), np.arange(k)].astype(np.float64) + rdiag.astype(np.float64)
    flipped = np.sign(qr.rdiag) != np.sign(rdiag)
    assert np.all(np.abs(xkk[flipped]) <= 50 * tol * np.abs(rdiag[flipped].astype(np.float64))), "a reflection took the other sign"
    if np.any(flipped):
        q, r = qr.get_q().to_numpy().astype(np.float64), qr.get_r().to_numpy().astype(np.float64)
        assert np.max(np.abs(q @ r - a)) <= 10 * tol * max(m, n) * scale  # still a valid factorisation of a
        assert np.max(np.abs(np.abs(qr.rdiag.astype(np.float64)) - np.abs(rdiag))) <= tol * max(m, n) * np.max(np.abs(rdiag))
        pytest.skip("a rounding-level diagonal entry took the other sign: element-wise parity does not apply")
    assert np.max(np.abs(qr.rdiag.astype(np.float64) - rdiag)) <= tol * max(m, n) * np.max(np.abs(rdiag))
    assert np.max(np.abs(got - packed)) <= tol * max(m, n) * scale
    return qr, packed, rdiag


@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (17, 17), (64, 64), (130, 130), (300, 300), (500, 200), (200, 500),
                                   (129, 1000), (1000, 129), (1024, 1024), (2000, 1300)])
def test_packed_qr_parity_f64(oracle, shape):
    a = oracle.fill(shape, 1)
    qr, packed, rdiag = _parity(oracle, a, 1e-12)
    m, n = shape
    if m <= 1100:
        q = qr.get_q().to_numpy()
        r = qr.get_r().to_numpy()
        ref_q = oracle.qr_get_q(packed, rdiag)
        assert np.max(np.abs(q - ref_q)) <= 1e-12 * max(m, n)
        assert np.array_equal(r, np.triu(r)) and np.max(np.abs(r - oracle.qr_get_r(packed, rdiag))) <= 1e-12 * max(m, n) * np.max(np.abs(r))
        assert np.max(np.abs(q @ r - a)) <= 1e-13 * max(m, n)
        k = min(m, n)
        assert np.max(np.abs(q[:, :k].T @ q[:, :k] - np.eye(k))) <= 1e-13 * max(m, n)
        if m > n:
            assert not np.any(q[:, n:])  # qr.rs:158-161: the identity only covers min(m, n) columns


def test_signed_inputs_and_backward_error_4096(oracle):
    """m = n = 4096, entries in [-0.5, 0.5): the sign rule (:58) sees both branches; ||QR - A|| through R'R = A'A."""
    n = 4096
    a = oracle.fill((n, n), 5) - 0.5
    qr = QRDecomposition.new(Matrix.from_numpy(a))
    r = np.triu(qr.get_r().to_numpy())
    assert np.any(qr.rdiag > 0) and np.any(qr.rdiag < 0)
    lhs, rhs = r.T @ r, a.T @ a
    assert np.max(np.abs(lhs - rhs)) <= 1e-12 * n * np.max(np.abs(rhs))
    # sampled columns against the oracle's packed factor of the leading 4096 x 256 columns (reflections are left-looking:
    # column j of the packed matrix depends only on columns <= j)
    packed, rdiag = oracle.qr(np.ascontiguousarray(a[:, :256]))
    got = qr.get_qr().to_numpy()[:, :256]
    assert np.max(np.abs(got - packed)) <= 1e-12 * n * np.max(np.abs(packed))
    assert np.max(np.abs(qr.rdiag[:256] - rdiag)) <= 1e-12 * n * np.max(np.abs(rdiag))


@pytest.mark.parametrize("shape", [(3, 3), (100, 100), (300, 200), (200, 300), (700, 700)])
def test_packed_qr_parity_f32(oracle, shape):
    a = oracle.fill(shape, 2, np.float32)
    qr, packed, rdiag = _parity(oracle, a, 1e-4)
    q, r = qr.get_q().to_numpy().astype(np.float64), qr.get_r().to_numpy().astype(np.float64)
    assert np.max(np.abs(q @ r - a)) <= 1e-5 * max(shape)


def test_zero_column_is_skipped_and_rank_deficient_solve_is_none(oracle):
    a = oracle.fill((40, 40), 3)
    a[:, 7] = 0.0
    a[:7, 7] = 0.0
    z = np.zeros((6, 6))
    qz = QRDecomposition.new(Matrix.from_numpy(z))
    assert np.array_equal(qz.rdiag, np.zeros(6)) and not qz.is_full_rank()          # qr.rs:62, :110-117
    assert qz.solve(Matrix.from_numpy(np.ones((6, 1)))) is None                       # qr.rs:201-203
    assert np.array_equal(qz.get_q().to_numpy(), np.eye(6))                           # every reflection skipped (:166)
    _parity(oracle, a, 1e-12)


@pytest.mark.parametrize("n,nx", [(3, 1), (8, 2), (16, 16), (32, 5), (64, 3)])
def test_solve_matches_the_reference_arithmetic(oracle, n, nx):
    a, b = oracle.fill((n, n), 1), oracle.fill((n, nx), 3)
    packed, rdiag = oracle.qr(a)
    ref = oracle.qr_solve(packed, rdiag, b)
    x = QRDecomposition.new(Matrix.from_numpy(a)).solve(Matrix.from_numpy(b))
    assert x.rows() == n and x.cols() == nx
    assert np.max(np.abs(x.to_numpy() - ref)) <= 1e-9 * np.max(np.abs(ref))


def test_solve_panics_like_the_reference():
    tall = QRDecomposition.new(Matrix.from_numpy(REF_INPUTS[1]))
    with pytest.raises(la.Panic):
        tall.solve(Matrix.from_numpy(np.ones((3, 1))))      # Matrix::new(cols, nx, <m * nx values>), qr.rs:237
    with pytest.raises(la.Panic):
        tall.solve(Matrix.from_numpy(np.ones((2, 1))))      # b.rows() != m, qr.rs:200
    wide = QRDecomposition.new(Matrix.from_numpy(REF_INPUTS[2]))
    with pytest.raises(la.Panic):
        wide.is_full_rank()                                  # rdiag[j] out of bounds, qr.rs:112


def test_pinverse_chain_on_the_device(oracle):
    """A+ = (R'R)^-1 A' (src/matrix/mod.rs:1049-1057) with every intermediate in HBM, against the oracle's chain."""
    m, n = 600, 200
    a = oracle.fill((m, n), 9)
    got = DeviceMatrix.from_matrix(Matrix.from_numpy(a)).pinverse().to_matrix().to_numpy()
    packed, rdiag = oracle.qr(a)
    r = oracle.qr_get_r(packed, rdiag)
    rtr = oracle.gemm(np.ascontiguousarray(r.T), r)
    lu, piv, _ = oracle.lu(rtr)
    ref = oracle.gemm(oracle.lu_solve(lu, piv, np.eye(n)), np.ascontiguousarray(a.T))
    assert got.shape == (n, m)
    assert np.max(np.abs(got - ref)) <= 1e-9 * np.max(np.abs(ref))
    assert np.max(np.abs(got @ a - np.eye(n))) <= 1e-9
