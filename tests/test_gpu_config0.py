"""BASELINE configs[0] -- "f64 512x512 random Matrix*Matrix plus LUDecomposition::new and solve" -- replayed exactly as
SURVEY.md 8(d) states it: A = hash(seed 1), B = hash(seed 2), RHS = hash(seed 3); Mul 512^3, LU(512), solve with 1 and 16
right-hand sides, det, is_non_singular -- every result compared with the oracle's restatement of the reference loops
(src/matrix/mod.rs:965-973, src/decomp/lu.rs:104-168, :224-232, :237-278) through the mirror of the crate API."""
import numpy as np
import pytest

from gpu_util import max_rel_err
from la import LUDecomposition, Matrix

pytestmark = pytest.mark.gpu
N = 512


def test_config0_mul_lu_solve_512(oracle):
    a, b = oracle.fill((N, N), 1), oracle.fill((N, N), 2)
    A, B = Matrix.from_numpy(a), Matrix.from_numpy(b)
    # Mul
    c = (A * B).to_numpy()
    assert max_rel_err(c, oracle.gemm(a, b)) <= 1e-12 * N
    # LU: identical pivots, element error, backward error within 10x
    ref_lu, ref_piv, ref_sign = oracle.lu(a)
    dec = LUDecomposition.new(A)
    assert np.array_equal(dec.get_piv(), ref_piv)
    assert dec.pospivsign == ref_sign
    lu = dec.get_lu().to_numpy()
    den = np.maximum(np.abs(ref_lu), float(np.max(np.abs(a))))
    assert float(np.max(np.abs(lu - ref_lu) / den)) <= 1e-12 * N
    be_ref = oracle.lu_backward_error(a, ref_lu, ref_piv)
    be = oracle.lu_backward_error(a, lu, dec.get_piv())
    assert be <= 10 * max(be_ref, np.finfo(np.float64).eps)
    # l * u == p * a within the same bar (get_l / get_u / get_p, lu.rs:184-220)
    l, u, p = dec.get_l().to_numpy(), dec.get_u().to_numpy(), dec.get_p().to_numpy()
    assert np.linalg.norm(l @ u - p @ a) / np.linalg.norm(a) <= 10 * max(be_ref, 1e-16)
    # det / is_non_singular
    assert dec.is_non_singular() == oracle.lu_is_non_singular(ref_lu)
    d, rd = dec.det(), oracle.lu_det(ref_lu, ref_sign)
    assert np.sign(d) == np.sign(rd) and np.isfinite(d) == np.isfinite(rd)
    if np.isfinite(rd) and rd != 0:
        assert abs(d - rd) / abs(rd) <= 1e-12 * N
    # solve with 1 and 16 right-hand sides
    for nx in (1, 16):
        rhs = oracle.fill((N, nx), 3)
        ref_x = oracle.lu_solve(ref_lu, ref_piv, rhs)
        x = dec.solve(Matrix.from_numpy(rhs)).to_numpy()
        scale = np.maximum(np.abs(ref_x), float(np.max(np.abs(ref_x))))
        assert float(np.max(np.abs(x - ref_x) / scale)) <= 1e-12 * N * 100  # cond(A) amplifies the factor difference
        res = np.linalg.norm(a @ x - rhs) / (np.linalg.norm(a) * np.linalg.norm(x))
        res_ref = np.linalg.norm(a @ ref_x - rhs) / (np.linalg.norm(a) * np.linalg.norm(ref_x))
        assert res <= 10 * max(res_ref, 1e-16)
