"""Residuals of the inverse-block sweep solve vs the reference's substitution on ill-conditioned systems (ADVICE, round 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
import numpy as np
from oracle import oracle as o
from la import Matrix, LUDecomposition
o.build()
rng = np.random.default_rng(7)
n = 1024
for cond in (1e4, 1e8, 1e12):
    q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
    q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
    a = np.ascontiguousarray(q1 @ np.diag(np.logspace(0, -np.log10(cond), n)) @ q2)
    for nx in (3, 40):
        b = np.ascontiguousarray(rng.standard_normal((n, nx)))
        lu, piv, _ = o.lu(a)
        xr = o.lu_solve(lu, piv, b)
        xg = LUDecomposition.new(Matrix.from_numpy(a)).solve(Matrix.from_numpy(b)).to_numpy()
        res = lambda x: np.linalg.norm(a @ x - b) / (np.linalg.norm(a) * np.linalg.norm(x))
        print(f"cond {cond:.0e} nx {nx}: residual ref {res(xr):.2e} gpu {res(xg):.2e}  |x_gpu - x_ref|/|x_ref| {np.linalg.norm(xg - xr) / np.linalg.norm(xr):.2e}", flush=True)
# graded matrix: rows scaled by 10^(-8 i / n)
a = np.ascontiguousarray(rng.standard_normal((n, n)) * np.logspace(0, -8, n)[:, None])
b = np.ascontiguousarray(rng.standard_normal((n, 5)))
lu, piv, _ = o.lu(a); xr = o.lu_solve(lu, piv, b)
xg = LUDecomposition.new(Matrix.from_numpy(a)).solve(Matrix.from_numpy(b)).to_numpy()
print(f"graded rows: residual ref {np.linalg.norm(a @ xr - b) / (np.linalg.norm(a) * np.linalg.norm(xr)):.2e} gpu {np.linalg.norm(a @ xg - b) / (np.linalg.norm(a) * np.linalg.norm(xg)):.2e}")
