"""Device-resident chains: `Matrix::t` (mod.rs:653-669), `permute_rows` (:757-759), `id`
(:416-426), operator `*` (:957-998) and `inverse` (:1034-1037) on matrices that stay in HBM between operations, checked
against the oracle / numpy on the same inputs."""
import numpy as np
import pytest

from la import DeviceMatrix, LaError, Matrix, Panic

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (3, 2), (1, 77), (33, 65), (257, 130), (1000, 31)])
def test_transpose_matches_reference_walk(oracle, dtype, shape):
    a = oracle.fill(shape, 5, dtype)
    got = DeviceMatrix.from_matrix(Matrix.from_numpy(a)).t().to_matrix()
    assert got.rows() == shape[1] and got.cols() == shape[0]
    assert np.array_equal(got.to_numpy(), a.T)  # pure data movement: bit-exact


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_permute_rows(oracle, dtype):
    a = oracle.fill((50, 37), 6, dtype)
    d = DeviceMatrix.from_matrix(Matrix.from_numpy(a))
    rows = [49, 0, 7, 7, 13]  # sub_matrix(rows, ..): repeats allowed, any length
    assert np.array_equal(d.permute_rows(rows).to_matrix().to_numpy(), a[rows])
    perm = np.random.default_rng(0).permutation(50)
    assert np.array_equal(d.permute_rows(perm).to_matrix().to_numpy(), a[perm])
    with pytest.raises((Panic, LaError)):
        d.permute_rows([0, 50])  # the reference panics on an out-of-range row


def test_identity_and_mul_stay_on_device(oracle):
    a = oracle.fill((96, 64), 7)
    d = DeviceMatrix.from_matrix(Matrix.from_numpy(a))
    assert np.array_equal((DeviceMatrix.id(96) * d).to_matrix().to_numpy(), a)  # exact: products with 0 and 1 only
    with pytest.raises(Panic):
        d * d  # 96x64 * 96x64: mod.rs:961


@pytest.mark.parametrize("m,n", [(40, 12), (700, 300), (2048, 640)])
def test_pinverse_chain_device_resident(oracle, m, n):
    """pinverse (mod.rs:1049-1057): (a' a)^-1 a' for a tall matrix, every intermediate in HBM; compared with the same chain
    evaluated by the oracle (gemm -> lu -> solve with the identity -> gemm)."""
    a = oracle.fill((m, n), 8)
    d = DeviceMatrix.from_matrix(Matrix.from_numpy(a))
    dt = d.t()
    pinv = ((dt * d).inverse() * dt).to_matrix().to_numpy()
    at = np.ascontiguousarray(a.T)
    g = oracle.gemm(at, a)
    lu, piv, _ = oracle.lu(g)
    ginv = oracle.lu_solve(lu, piv, oracle.identity(n))
    ref = oracle.gemm(ginv, at)
    cond = np.linalg.cond(g)
    assert np.max(np.abs(pinv - ref)) / np.max(np.abs(ref)) <= 1e-12 * n * max(1.0, cond / n)
    # size-independent property: pinv * a == I to the conditioning of a' a
    assert np.max(np.abs(pinv @ a - np.eye(n))) <= 1e-13 * n * cond


def test_inverse_singular_is_none_and_roundtrip(oracle):
    s = DeviceMatrix.from_matrix(Matrix.from_numpy(np.array([[1.0, 2.0], [2.0, 4.0]])))
    assert s.inverse() is None  # mod.rs:1542-1546
    n = 1536  # multi-panel factorisation + many-right-hand-side solve
    a = oracle.fill((n, n), 9) + n * np.eye(n)  # well conditioned
    d = DeviceMatrix.from_matrix(Matrix.from_numpy(a))
    prod = (d * d.inverse()).to_matrix().to_numpy()
    assert np.max(np.abs(prod - np.eye(n))) <= 1e-12 * n
