#!/usr/bin/env python3
"""Generates tests/golden/ref_tests.json.

Inputs are the literal matrices of the reference's OWN unit tests for the hot path (file:line given per case,
relative to /root/reference).  Two kinds of expectation are recorded:

  "asserted" : what the reference test itself asserts (exact values, L*U == P*A, None on singular, ...).
  "derived"  : packed LU / piv / pospivsign / det / solve computed by oracle/pyref.py (pure Python, IEEE
               binary64, no FMA -- the reference's arithmetic).  Stored as C99 hex floats so the comparison is
               bit-exact.  These reproduce SURVEY.md section 4.1.

The script refuses to write the file unless every "asserted" expectation holds for the derived values, i.e. the
restatement passes the reference's own tests.  Run:  python tests/golden/make_ref_goldens.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyref  # noqa: E402

APPROX_EPS = 1e-6  # src/approxeq.rs:34-47 (absolute)


def hexl(v):
    return [float(x).hex() for x in v]


LU_CASES = [
    # name, source, m, n, A, asserted
    ("lu_square", "src/decomp/lu.rs:281-289", 3, 3, [1.0, 2.0, 0.0, 3.0, 6.0, -1.0, 1.0, 2.0, 1.0], {"lu_eq_pa": True}),
    ("lu_m_over_n", "src/decomp/lu.rs:291-299", 3, 2, [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], {"lu_eq_pa": True}),
    ("lu_m_under_n", "src/decomp/lu.rs:301-309", 2, 3, [1.0, 2.0, 3.0, 4.0, 5.0, 6.0], {"lu_eq_pa": True}),
    ("lu_solve", "src/decomp/lu.rs:311-317", 3, 3, [2.0, 1.0, 0.0, 1.0, 1.0, 0.0, 0.0, 0.0, 1.0],
     {"solve_approx": {"b": [1.0, 2.0, 3.0], "nx": 1, "x": [-1.0, 3.0, 3.0]}}),
    ("lu_solve_singular", "src/decomp/lu.rs:328-334", 2, 2, [2.0, 6.0, 1.0, 3.0],
     {"solve_none": {"b": [1.0, 2.0], "nx": 1}, "singular": True}),
    ("lu_is_singular_b", "src/decomp/lu.rs:341-344", 2, 2, [2.0, 6.0, 1.0, 4.0], {"singular": False}),
    ("lu_non_singular_a", "src/decomp/lu.rs:348-351,358-361", 2, 2, [4.0, 8.0, 3.0, 4.0],
     {"singular": False, "det_exact": -8.0}),
    ("lu_non_singular_b", "src/decomp/lu.rs:353-355", 2, 2, [4.0, 6.0, 2.0, 3.0], {"singular": True}),
    ("lu_det_zero", "src/decomp/lu.rs:363-366", 2, 2, [4.0, 8.0, 2.0, 4.0], {"det_exact": 0.0}),
    ("matrix_det_inverse", "src/matrix/mod.rs:1506-1516,1527-1540", 3, 3,
     [6.0, -7.0, 10.0, 0.0, 3.0, -1.0, 0.0, 5.0, -7.0],
     {"det_exact": -96.0,
      "inverse_approx": [v / 96.0 for v in [16.0, -1.0, 23.0, 0.0, 42.0, -6.0, 0.0, 30.0, -18.0]]}),
    ("matrix_solve", "src/matrix/mod.rs:1518-1523", 3, 3, [1.0, 1.0, 1.0, 1.0, -1.0, 4.0, 2.0, 3.0, -5.0],
     {"solve_exact": {"b": [3.0, 4.0, 0.0], "nx": 1, "x": [1.0, 1.0, 1.0]}}),
    ("matrix_inverse_singular", "src/matrix/mod.rs:1542-1546", 2, 2, [2.0, 6.0, 1.0, 3.0],
     {"singular": True, "inverse_none": True}),
    ("matrix_is_singular", "src/matrix/mod.rs:1554-1571", 2, 2, [2.0, 6.0, 6.0, 3.0], {"singular": False}),
]

MUL_CASES = [
    ("mul_int", "src/matrix/mod.rs:1479-1484; src/matrix/mmatrix.rs:234-241", 2, 2, 2,
     [1, 2, 3, 4], [3, 4, 5, 6], [13, 16, 29, 36]),
]


def main():
    out = {"_generator": "tests/golden/make_ref_goldens.py", "_arith": "IEEE binary64, separate mul/add (pyref)",
           "lu": [], "mul": []}
    for name, src, m, n, a, asserted in LU_CASES:
        lu, piv, pos = pyref.lu_new(a, m, n)
        rec = {"name": name, "source": src, "m": m, "n": n, "a": hexl(a), "asserted": asserted,
               "derived": {"lu": hexl(lu), "piv": piv, "pospivsign": pos}}
        # --- check the reference's own assertions against the restatement ---
        if asserted.get("lu_eq_pa"):
            l, lm, ln = pyref.get_l(lu, m, n)
            u, um, un = pyref.get_u(lu, m, n)
            assert pyref.mul(l, u, lm, ln, un) == pyref.permute_rows(a, m, n, piv), name
        if m == n:
            ns = pyref.is_non_singular(lu, n)
            d = pyref.det(lu, n, pos)
            rec["derived"]["non_singular"] = ns
            rec["derived"]["det"] = float(d).hex()
            if "singular" in asserted:
                assert ns == (not asserted["singular"]), name
            if "det_exact" in asserted:
                assert d == asserted["det_exact"], (name, d)
            for key in ("solve_approx", "solve_exact", "solve_none"):
                if key in asserted:
                    sp = asserted[key]
                    x = pyref.solve(lu, m, n, piv, sp["b"], sp["nx"])
                    if key == "solve_none":
                        assert x is None, name
                        rec["derived"]["solve"] = None
                    else:
                        assert x is not None
                        if key == "solve_exact":
                            assert x == sp["x"], (name, x)
                        else:
                            assert all(abs(p - q) < APPROX_EPS for p, q in zip(x, sp["x"])), (name, x)
                        rec["derived"]["solve"] = {"b": hexl(sp["b"]), "nx": sp["nx"], "x": hexl(x)}
            if "inverse_approx" in asserted or asserted.get("inverse_none"):
                ident = [1.0 if i == j else 0.0 for i in range(n) for j in range(n)]  # Matrix::id, mod.rs:416-426
                inv = pyref.solve(lu, m, n, piv, ident, n)
                if asserted.get("inverse_none"):
                    assert inv is None
                    rec["derived"]["inverse"] = None
                else:
                    assert all(abs(p - q) < APPROX_EPS for p, q in zip(inv, asserted["inverse_approx"])), name
                    rec["derived"]["inverse"] = hexl(inv)
        out["lu"].append(rec)
    for name, src, m, k, n, a, b, c in MUL_CASES:
        assert pyref.mul(a, b, m, k, n) == c
        assert pyref.mul([float(v) for v in a], [float(v) for v in b], m, k, n) == [float(v) for v in c]
        out["mul"].append({"name": name, "source": src, "m": m, "k": k, "n": n, "a": a, "b": b, "c": c})
    path = os.path.join(HERE, "ref_tests.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, len(out["lu"]), "lu cases,", len(out["mul"]), "mul cases")


if __name__ == "__main__":
    main()
