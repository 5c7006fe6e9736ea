"""The 64-wide split panel of the LU look-ahead (lu.cu, LA_LU_SPLIT_ROWS) is only taken for tall trailing matrices by
default (>= 12288 rows: exercised by tests/test_gpu_full_size.py).  Force it for every panel in a child process -- the
knob is read once per process -- and check pivots, factors and backward error against the oracle on small systems."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = textwrap.dedent("""
    import sys
    sys.path.insert(0, %r); sys.path.insert(0, %r)
    import numpy as np
    from oracle import oracle
    from la import LUDecomposition, Matrix
    for n, m in ((700, 700), (1154, 1154), (900, 640), (640, 900), (1536, 1536), (2000, 1300)):
        a = oracle.fill((n, m), 1)
        ref_lu, ref_piv, ref_sign = oracle.lu(a)
        dec = LUDecomposition.new(Matrix.from_numpy(a))
        assert np.array_equal(dec.get_piv(), ref_piv) and dec.pospivsign == ref_sign, (n, m, "pivots differ")
        lu = dec.get_lu().to_numpy()
        err = np.max(np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), np.max(np.abs(a))))
        assert err <= 1e-12 * max(n, m), (n, m, err)
        be, be_ref = (oracle.lu_backward_error(a, x, ref_piv) for x in (lu, ref_lu))
        assert be <= 10 * max(be_ref, 1e-16), (n, m, be, be_ref)
    print("split ok")
""") % (ROOT, os.path.join(ROOT, "rust-la_b200", "python"))


@pytest.mark.gpu
def test_split_panel_forced_for_every_panel():
    env = dict(os.environ, LA_LU_SPLIT_ROWS="129")
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "split ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("group", [2, 3, 4])
def test_grouped_bulk_update_forced(group):
    """Grouped trailing updates (lu.cu, LA_LU_GROUP / LA_LU_GROUP_ROWS: K = group * 128 bulk GEMMs, the chain catching up
    the next panel's columns itself) are only taken for trailing matrices of >= 6144 rows by default; force them from the
    first panel on, with and without split panels."""
    for split in ("0", "129"):
        env = dict(os.environ, LA_LU_GROUP=str(group), LA_LU_GROUP_ROWS="1", LA_LU_SPLIT_ROWS=split)
        out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0 and "split ok" in out.stdout, out.stdout + out.stderr
