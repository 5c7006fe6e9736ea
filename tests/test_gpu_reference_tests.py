"""The reference's own unit tests for the hot path, replayed through the CUDA path (they read like the originals).
Sources: src/matrix/mod.rs:1479-1571, src/matrix/mmatrix.rs:234-259, src/decomp/lu.rs:281-375, and the golden vectors
of tests/golden/ref_tests.json (bit-exact packed LU / piv / det / solve for those inputs)."""
import json
import os

import numpy as np
import pytest

import la
from la import LUDecomposition, Matrix, Panic, m

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_tests.json")))


def test_mul():  # src/matrix/mod.rs:1479-1484 (integer matrices)
    m1 = m("1, 2; 3, 4")
    m2 = m("3, 4; 5, 6")
    assert (m1 * m2).get_data().tolist() == [13, 16, 29, 36]
    assert (m("1.0, 2.0; 3.0, 4.0") * m("3.0, 4.0; 5.0, 6.0")).get_data().tolist() == [13.0, 16.0, 29.0, 36.0]
    f = (m("1.0, 2.0; 3.0, 4.0", ) * m("3.0, 4.0; 5.0, 6.0"))
    assert f == m("13.0, 16.0; 29.0, 36.0")
    a32 = Matrix.new(2, 2, np.array([1, 2, 3, 4], dtype=np.float32))
    b32 = Matrix.new(2, 2, np.array([3, 4, 5, 6], dtype=np.float32))
    assert (a32 * b32).get_data().tolist() == [13.0, 16.0, 29.0, 36.0]


def test_mul_incompatible():  # mod.rs:1486-1492
    with pytest.raises(Panic):
        m("1, 2; 3, 4") * m("1, 2; 3, 4; 5, 6")


def test_mmul():  # src/matrix/mmatrix.rs:234-241
    a, b = m("1, 2; 3, 4"), m("3, 4; 5, 6")
    c = m("0, 0; 0, 0")
    a.mmul(b, c)
    assert c.get_data().tolist() == [13, 16, 29, 36]


@pytest.mark.parametrize("spec", ["1.0, 2.0, 0.0; 3.0, 6.0, -1.0; 1.0, 2.0, 1.0",   # lu.rs:281-289
                                  "1.0, 2.0; 3.0, 4.0; 5.0, 6.0",                     # lu.rs:291-299
                                  "1.0, 2.0, 3.0; 4.0, 5.0, 6.0"])                    # lu.rs:301-309
def test_lu_l_times_u_equals_p_times_a(spec):
    a = m(spec)
    lu = LUDecomposition.new(a)
    l, u, p = lu.get_l(), lu.get_u(), lu.get_p()
    assert l * u == p * a   # exact ==, as in the reference


def test_lu_solve():  # lu.rs:311-317
    a = m("2.0, 1.0, 0.0; 1.0, 1.0, 0.0; 0.0, 0.0, 1.0")
    lu = LUDecomposition.new(a)
    b = m("1.0; 2.0; 3.0")
    assert lu.solve(b).approx_eq(m("-1.0; 3.0; 3.0"))


def test_lu_solve_incompatible():  # lu.rs:319-326
    lu = LUDecomposition.new(m("2.0, 1.0, 0.0; 1.0, 1.0, 0.0; 0.0, 0.0, 1.0"))
    with pytest.raises(Panic):
        lu.solve(m("1.0; 2.0; 3.0; 4.0"))


def test_lu_solve_singular():  # lu.rs:328-334
    lu = LUDecomposition.new(m("2.0, 6.0; 1.0, 3.0"))
    assert lu.solve(m("1.0; 2.0")) is None


def test_lu_is_singular():  # lu.rs:336-356
    assert LUDecomposition.new(m("2.0, 6.0; 1.0, 3.0")).is_singular()
    assert not LUDecomposition.new(m("2.0, 6.0; 1.0, 4.0")).is_singular()
    assert LUDecomposition.new(m("4.0, 8.0; 3.0, 4.0")).is_non_singular()
    assert not LUDecomposition.new(m("4.0, 6.0; 2.0, 3.0")).is_non_singular()


def test_lu_det():  # lu.rs:358-367: exact
    assert LUDecomposition.new(m("4.0, 8.0; 3.0, 4.0")).det() == -8.0
    assert LUDecomposition.new(m("4.0, 8.0; 2.0, 4.0")).det() == 0.0


def test_lu_det_not_square():  # lu.rs:369-375
    lu = LUDecomposition.new(m("1.0, 2.0, 3.0; 4.0, 5.0, 6.0"))
    with pytest.raises(Panic):
        lu.det()


def test_matrix_det_solve_inverse():  # src/matrix/mod.rs:1506-1546
    a = m("6.0, -7.0, 10.0; 0.0, 3.0, -1.0; 0.0, 5.0, -7.0")
    assert (a.det() - -96.0) <= 1e-6
    assert a.det() == -96.0
    s = m("1.0, 1.0, 1.0; 1.0, -1.0, 4.0; 2.0, 3.0, -5.0")
    assert s.solve(m("3.0; 4.0; 0.0")) == m("1.0; 1.0; 1.0")           # mod.rs:1522 uses eq
    inv = a.inverse()
    expect = Matrix.new(3, 3, np.array([16.0, -1.0, 23.0, 0.0, 42.0, -6.0, 0.0, 30.0, -18.0]) / 96.0)
    assert inv.approx_eq(expect)
    assert (a * inv).approx_eq(Matrix.id(3, 3))
    assert m("2.0, 6.0; 1.0, 3.0").inverse() is None                    # mod.rs:1542-1546


def test_matrix_is_singular():  # mod.rs:1554-1571
    assert m("2.0, 6.0; 1.0, 3.0").is_singular()
    assert not m("2.0, 6.0; 6.0, 3.0").is_singular()
    assert m("2.0, 6.0; 6.0, 3.0").is_non_singular()


@pytest.mark.parametrize("case", GOLD["lu"], ids=[c["name"] for c in GOLD["lu"]])
def test_golden_vectors_bit_exact(case):
    """On these <= 3x3 inputs the blocked CUDA path must reproduce the reference bit-for-bit."""
    mm, nn = case["m"], case["n"]
    a = Matrix.new(mm, nn, np.array([float.fromhex(x) for x in case["a"]]))
    d = case["derived"]
    lu = LUDecomposition.new(a)
    want = np.array([float.fromhex(x) for x in d["lu"]])
    assert np.array_equal(lu.get_lu().get_data().view(np.uint64), want.view(np.uint64))
    assert lu.get_piv().tolist() == d["piv"]
    assert lu.pospivsign == d["pospivsign"]
    if mm == nn:
        assert lu.is_non_singular() == d["non_singular"]
        det = lu.det()
        assert np.float64(det).view(np.uint64) == np.float64(float.fromhex(d["det"])).view(np.uint64)  # incl. -0.0
        if "solve" in d:
            asserted = case["asserted"]
            sp = asserted.get("solve_approx") or asserted.get("solve_exact") or asserted.get("solve_none")
            x = lu.solve(Matrix.new(mm, sp["nx"], np.array(sp["b"], dtype=np.float64)))
            if d["solve"] is None:
                assert x is None
            else:
                wantx = np.array([float.fromhex(v) for v in d["solve"]["x"]])
                assert np.array_equal(x.get_data().view(np.uint64), wantx.view(np.uint64))
        if "inverse" in d:
            inv = a.inverse()
            if d["inverse"] is None:
                assert inv is None
            else:
                wanti = np.array([float.fromhex(v) for v in d["inverse"]])
                assert np.array_equal(inv.get_data().view(np.uint64), wanti.view(np.uint64))
