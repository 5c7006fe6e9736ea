"""Host-side logic of the `la` mirror that needs no GPU: constructors, the m! analogue, and the reference's panics,
which must fire BEFORE any FFI call (src/matrix/mod.rs:208-209, :961; src/decomp/lu.rs:225,240; mmatrix.rs:83-85)."""
import numpy as np
import pytest

import la
from la import Matrix, Panic, m


def test_new_asserts_like_reference():
    a = Matrix.new(2, 2, [1.0, 2.0, 3.0, 4.0])
    assert a.rows() == 2 and a.cols() == 2 and a.get(1, 0) == 3.0
    with pytest.raises(Panic):
        Matrix.new(2, 2, [1.0, 2.0, 3.0])        # mod.rs:208
    with pytest.raises(Panic):
        Matrix.new(0, 0, [])                     # mod.rs:209


def test_m_macro_analogue():
    a = m("1, 2; 3, 4")
    assert a.get_data().dtype == np.int64 and a.get_data().tolist() == [1, 2, 3, 4] and a.rows() == 2
    b = m("1.0, 2.0, 0.0; 3.0, 6.0, -1.0; 1.0, 2.0, 1.0")
    assert b.get_data().dtype == np.float64 and b.cols() == 3
    c = m([[1.0], [2.0], [3.0]])
    assert c.rows() == 3 and c.cols() == 1
    assert m([[1, 2], [3, 4]]) == a


def test_id_matches_reference():
    i = Matrix.id(2, 3)
    assert i.get_data().tolist() == [1, 0, 0, 0, 1, 0]
    assert Matrix.id(3, 2).get_data().tolist() == [1, 0, 0, 1, 0, 0]


def test_mul_incompatible_panics_before_ffi():
    """src/matrix/mod.rs:1486-1492 #[should_panic]."""
    with pytest.raises(Panic):
        m("1, 2; 3, 4") * m("1, 2; 3, 4; 5, 6")


def test_mmul_shape_panics():
    """src/matrix/mmatrix.rs:243-259 #[should_panic] cases."""
    a, b = m("1, 2; 3, 4"), m("3, 4; 5, 6")
    with pytest.raises(Panic):
        a.mmul(m("1, 2, 3; 4, 5, 6; 7, 8, 9"), m("0, 0; 0, 0"))
    with pytest.raises(Panic):
        a.mmul(b, m("0, 0, 0; 0, 0, 0"))
    with pytest.raises(Panic):
        a.mmul(b, m("0, 0; 0, 0; 0, 0"))


def test_det_non_square_panics():
    with pytest.raises(Panic):
        m("1.0, 2.0, 3.0; 4.0, 5.0, 6.0").det()  # mod.rs:1026


def test_approx_eq_is_absolute_1e6():
    a = m("1.0, 2.0")
    assert a.approx_eq(Matrix.new(1, 2, [1.0 + 5e-7, 2.0]))
    assert not a.approx_eq(Matrix.new(1, 2, [1.0 + 2e-6, 2.0]))
