"""Pure-Python restatement of rust-la's hot-path loops (TEST INFRASTRUCTURE ONLY, small cases).

Python floats are IEEE binary64 and CPython never fuses multiply-add, so these loops reproduce the
reference's f64 arithmetic bit-for-bit.  They exist to (1) generate tests/golden/ref_tests.json from the
inputs of the reference's own unit tests and (2) cross-check the C oracle (oracle/la_oracle.c) independently.

Reference citations (relative to /root/reference):
  mul        src/matrix/mod.rs:957-980
  lu_new     src/decomp/lu.rs:104-168
  non_sing   src/decomp/lu.rs:174-182
  det        src/decomp/lu.rs:224-232
  solve      src/decomp/lu.rs:237-278
  get_l/u/p  src/decomp/lu.rs:184-220
"""


def mul(a, b, m, k, n):
    """C = A*B, row-major flat lists. src/matrix/mod.rs:965-973."""
    c = [None] * (m * n)
    for row in range(m):
        for col in range(n):
            res = 0 * a[0]  # num::zero() of the element type (int stays int, float stays float)
            for idx in range(k):
                res = res + a[row * k + idx] * b[idx * n + col]
            c[row * n + col] = res
    return c


def lu_new(a, m, n):
    """Returns (lu, piv, pospivsign). src/decomp/lu.rs:104-168."""
    lu = list(a)
    piv = list(range(m))
    pos = True
    for j in range(n):
        for i in range(m):
            s = 0.0
            for k in range(min(i, j)):
                s = s + lu[i * n + k] * lu[k * n + j]
            lu[i * n + j] = lu[i * n + j] - s
        p = j
        for i in range(j + 1, m):
            if abs(lu[i * n + j]) > abs(lu[p * n + j]):
                p = i
        if p != j:
            for k in range(n):
                lu[p * n + k], lu[j * n + k] = lu[j * n + k], lu[p * n + k]
            piv[p], piv[j] = piv[j], piv[p]
            pos = not pos
        if j < m and lu[j * n + j] != 0.0:
            for i in range(j + 1, m):
                lu[i * n + j] = lu[i * n + j] / lu[j * n + j]
    return lu, piv, pos


def is_non_singular(lu, n):
    """src/decomp/lu.rs:174-182."""
    for j in range(n):
        if lu[j * n + j] == 0.0:
            return False
    return True


def det(lu, n, pos):
    """src/decomp/lu.rs:224-232."""
    d = 1.0 if pos else -1.0
    for j in range(n):
        d = d * lu[j * n + j]
    return d


def solve(lu, m, n, piv, b, nx):
    """Returns x (flat m*nx) or None when singular. src/decomp/lu.rs:237-278."""
    if not is_non_singular(lu, n):
        return None
    x = [b[piv[i] * nx + j] for i in range(m) for j in range(nx)]
    for k in range(n):
        for i in range(k + 1, n):
            for j in range(nx):
                x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * lu[i * n + k]
    for k in range(n - 1, -1, -1):
        for j in range(nx):
            x[k * nx + j] = x[k * nx + j] / lu[k * n + k]
        for i in range(k):
            for j in range(nx):
                x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * lu[i * n + k]
    return x


def get_l(lu, m, n):
    """src/decomp/lu.rs:184-202."""
    nn = n if m >= n else m
    return [lu[i * n + j] if i > j else (1.0 if i == j else 0.0) for i in range(m) for j in range(nn)], m, nn


def get_u(lu, m, n):
    """src/decomp/lu.rs:204-215."""
    mm = n if m >= n else m
    return [lu[i * n + j] if i <= j else 0.0 for i in range(mm) for j in range(n)], mm, n


def permute_rows(a, m, n, piv):
    """P*A = A(piv,:). src/decomp/lu.rs:217-220 via Matrix::permute_rows."""
    return [a[piv[i] * n + j] for i in range(m) for j in range(n)]


def qr_new(a, m, n):
    """Returns (qr, rdiag). src/decomp/qr.rs:26-106 (pure-Python floats: binary64, never fused)."""
    import math
    qr = list(a)
    dc = min(m, n)
    rdiag = [None] * dc
    for minor in range(dc):
        x_norm_sqr = 0.0
        for i in range(minor, m):
            c = qr[i * n + minor]
            x_norm_sqr = x_norm_sqr + c * c
        aa = -math.sqrt(x_norm_sqr) if qr[minor * n + minor] > 0.0 else math.sqrt(x_norm_sqr)
        rdiag[minor] = aa
        if aa != 0.0:
            qr[minor * n + minor] = qr[minor * n + minor] - aa
            for column in range(minor + 1, n):
                x_dot_u = 0.0
                for row in range(minor, m):
                    x_dot_u = x_dot_u + qr[row * n + minor] * qr[row * n + column]
                factor = x_dot_u / (aa * qr[minor * n + minor])
                for row in range(minor, m):
                    qr[row * n + column] = qr[row * n + column] + factor * qr[row * n + minor]
    return qr, rdiag


def qr_get_r(qr, rdiag, m, n):
    """src/decomp/qr.rs:138-152."""
    return [qr[i * n + j] if i < j else (rdiag[i] if i == j else 0.0) for i in range(m) for j in range(n)]


def qr_get_q(qr, rdiag, m, n):
    """src/decomp/qr.rs:155-194."""
    q = [0.0] * (m * m)
    for minor in range(min(m, n)):
        q[minor * m + minor] = 1.0
    for minor in reversed(range(min(m, n))):
        if qr[minor * n + minor] != 0.0:
            for column in range(minor, m):
                x_dot_u = 0.0
                for row in range(minor, m):
                    x_dot_u = x_dot_u + qr[row * n + minor] * q[row * m + column]
                factor = x_dot_u / (rdiag[minor] * qr[minor * n + minor])
                for row in range(minor, m):
                    q[row * m + column] = q[row * m + column] + factor * qr[row * n + minor]
    return q


def qr_solve(qr, rdiag, m, n, b, nx):
    """src/decomp/qr.rs:199-238: the m*nx work array after both phases (the reference then calls Matrix::new(n, nx, ..))."""
    x = list(b)
    for k in range(n):
        for j in range(nx):
            s = 0.0
            for i in range(k, m):
                s = s + qr[i * n + k] * x[i * nx + j]
            s = -s / qr[k * n + k]
            for i in range(k, m):
                x[i * nx + j] = x[i * nx + j] + s * qr[i * n + k]
    for k in reversed(range(n)):
        for j in range(nx):
            x[k * nx + j] = x[k * nx + j] / rdiag[k]
        for i in range(k):
            for j in range(nx):
                x[i * nx + j] = x[i * nx + j] - x[k * nx + j] * qr[i * n + k]
    return x
