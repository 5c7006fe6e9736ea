"""Device-resident elementwise operators and norms (csrc/elementwise.cu through la_elementwise_* / la_reduce_*; reference
src/matrix/mod.rs:487-527, :853-929, :1059-1115).  One IEEE operation per element: bit-identical to numpy (and so to the
reference's scalar loops); the reductions are compared with the reference's SEQUENTIAL sums (restated here with
math.fsum-free Python loops on small inputs, numpy cumulative order on large ones) within count * eps."""
import numpy as np
import pytest

from la import DeviceMatrix, Matrix, Panic

pytestmark = pytest.mark.gpu


def dm(a):
    return DeviceMatrix.from_matrix(Matrix.from_numpy(a))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (257, 129), (1000, 1001), (2048, 4096)])
def test_elementwise_bit_exact(oracle, dtype, shape):
    a = oracle.fill(shape, 1, dtype) - dtype(0.5)
    b = oracle.fill(shape, 2, dtype) + dtype(0.25)
    da, db = dm(a), dm(b)
    for got, ref in (((da + db), a + b), ((da - db), a - b), (da.elem_mul(db), a * b), (da.elem_div(db), a / b),
                     (-da, -a), (da.scale(3.7), dtype(3.7) * a)):
        out = got.to_matrix().to_numpy()
        assert out.dtype == dtype and np.array_equal(out.view(np.uint8), np.ascontiguousarray(ref).view(np.uint8))


def test_shape_panics():
    a, b = dm(np.ones((2, 3))), dm(np.ones((3, 2)))
    for f in (lambda: a + b, lambda: a - b, lambda: a.elem_mul(b), lambda: a.elem_div(b), lambda: a.vector_1_norm()):
        with pytest.raises(Panic):
            f()


def seq_sum(v):
    s = v.dtype.type(0)
    for x in v:
        s = s + x
    return s


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n", [1, 31, 1000, 70001])
def test_norms_against_the_sequential_sums(oracle, dtype, n):
    v = oracle.fill((n, 1), 3, dtype) - dtype(0.5)
    w = oracle.fill((n, 1), 4, dtype)
    d, e = dm(v), dm(w)
    eps = np.finfo(dtype).eps
    flat, wf = v.reshape(-1), w.reshape(-1)
    ref2 = np.sqrt(seq_sum(flat * flat))                    # mod.rs:1062-1067
    ref1 = seq_sum(np.abs(flat))                            # mod.rs:1078-1083
    refd = seq_sum(flat * wf)
    assert abs(d.vector_euclidean_norm() - ref2) <= n * eps * ref2
    assert abs(d.frobenius_norm() - ref2) <= n * eps * ref2
    assert abs(d.vector_1_norm() - ref1) <= n * eps * ref1
    assert d.vector_inf_norm() == np.max(np.abs(flat))      # mod.rs:1103-1115: exact
    assert abs(d.dot(e) - refd) <= n * eps * float(np.sum(np.abs(flat * wf)))


def test_frobenius_of_a_large_matrix_and_residual_use(oracle):
    """||A B - C||_F / ||C||_F entirely on the device: the parity harness's own residuals no longer need the host."""
    n = 2048
    a, b = oracle.fill((n, n), 1), oracle.fill((n, n), 2)
    da, db = dm(a), dm(b)
    c = da * db
    r = (c - da * db).frobenius_norm()
    assert r == 0.0
    nrm = c.frobenius_norm()
    ref = np.linalg.norm(a @ b)
    assert abs(nrm - ref) <= 1e-12 * ref
