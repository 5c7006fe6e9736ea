"""Host logic of the multi-device LU (la_lu_mg_*) without a GPU: the layout arithmetic the C library exports
(la_lu_mg_plan) and a numpy walk through the protocol on that layout -- every "device" holds only its block columns, the
owner factors a panel and hands over (L11 over L21, pivots), every device interchanges rows and updates its own columns --
against the oracle's restatement of LUDecomposition::new (src/decomp/lu.rs:104-168)."""
import numpy as np
import pytest

from la import sharding
from la._cabi import LaError


@pytest.mark.parametrize("n,ngpus", [(16384, 8), (16384, 3), (300, 2), (129, 4), (100, 5), (28672, 8), (29304, 2)])
def test_plan_is_a_partition_with_128_wide_blocks(n, ngpus):
    nb, nblk, ndev, ncols = sharding.lu_mg_plan(n, ngpus)
    assert nb == 128 and nblk == -(-n // 128) and ndev == min(ngpus, nblk)
    assert sum(ncols) == n and all(c == 0 for c in ncols[ndev:])
    want = [0] * ngpus
    for b in range(nblk):
        want[b % ndev] += min(128, n - 128 * b)
    assert ncols == want


def test_plan_narrows_the_blocks_when_a_panel_exceeds_shared_memory():
    """198 rows x 129 doubles fill the 200 KiB panel budget of an SM: beyond 148 x 198 rows the block columns get narrower in
    steps of 16, exactly as the single-device panel width does."""
    assert sharding.lu_mg_plan(29304, 4)[0] == 128
    assert sharding.lu_mg_plan(29305, 4)[0] == 112
    nb, nblk, ndev, ncols = sharding.lu_mg_plan(32768, 8)
    assert (nb, nblk, ndev) == (112, 293, 8) and sum(ncols) == 32768
    assert sharding.lu_mg_plan(65536, 8)[0] == 48
    assert sharding.lu_mg_plan(1000, 2, sm_count=4)[0] == 96   # a small device: 250 rows per SM
    with pytest.raises(LaError):
        sharding.lu_mg_plan(400000, 8)                          # no panel width >= 16 fits


def test_plan_rejects_bad_arguments():
    for args in ((0, 2), (100, 0), (100, 17), (1 << 30, 2)):
        with pytest.raises(LaError):
            sharding.lu_mg_plan(*args)


def _walk_protocol(a, ngpus):
    """numpy model of lu_mg_factor: local block-column arrays, owner panel, hand-over, per-device updates."""
    n = a.shape[0]
    nb, nblk, ndev, ncols = sharding.lu_mg_plan(n, ngpus)
    width = lambda b: min(nb, n - nb * b)  # noqa: E731
    local = [np.zeros((n, ncols[q])) for q in range(ndev)]
    for b in range(nblk):
        local[b % ndev][:, (b // ndev) * nb:(b // ndev) * nb + width(b)] = a[:, b * nb:b * nb + width(b)]
    piv = [np.arange(n) for _ in range(ndev)]
    sign = [True] * ndev
    for k in range(nblk):
        j0, jb, o, lc0 = k * nb, width(k), k % ndev, (k // ndev) * nb
        c1 = j0 + jb
        panel = local[o][:, lc0:lc0 + jb]            # a view: the owner factors in place, rows j0.. only
        ipiv = []
        for c in range(jb):                           # lu.rs:132-160 on the panel's columns
            col = np.abs(panel[j0 + c:, c])
            p = j0 + c + int(np.argmax(col))          # first maximum == strict '>' scan
            ipiv.append(p)
            if p != j0 + c:
                panel[[j0 + c, p], :] = panel[[p, j0 + c], :]
            if panel[j0 + c, c] != 0.0:
                panel[j0 + c + 1:, c] /= panel[j0 + c, c]
            panel[j0 + c + 1:, c + 1:] -= np.outer(panel[j0 + c + 1:, c], panel[j0 + c, c + 1:])
        slot = panel[j0:, :].copy()                   # what travels: L11 over L21 (and the pivots)
        l11 = np.tril(slot[:jb, :jb], -1) + np.eye(jb)
        for q in range(ndev):
            first_right = min(((k - q) // ndev + 1) * nb if k >= q else 0, ncols[q])
            for c, p in enumerate(ipiv):              # every device: its own piv / sign, its own columns
                if p != j0 + c:
                    piv[q][[j0 + c, p]] = piv[q][[p, j0 + c]]
                    sign[q] = not sign[q]
                    cols = np.ones(ncols[q], dtype=bool)
                    if q == o:
                        cols[lc0:lc0 + jb] = False    # interchanged inside the panel already
                    idx = np.nonzero(cols)[0]
                    local[q][np.ix_([j0 + c, p], idx)] = local[q][np.ix_([p, j0 + c], idx)]
            if first_right < ncols[q]:
                u12 = np.linalg.solve(l11, local[q][j0:c1, first_right:])
                local[q][j0:c1, first_right:] = u12
                local[q][c1:, first_right:] -= slot[jb:, :] @ u12
    lu = np.empty_like(a)
    for b in range(nblk):
        lu[:, b * nb:b * nb + width(b)] = local[b % ndev][:, (b // ndev) * nb:(b // ndev) * nb + width(b)]
    assert all(np.array_equal(piv[0], p) for p in piv) and len(set(sign)) == 1
    return lu, piv[0], sign[0]


@pytest.mark.parametrize("n,ngpus", [(300, 2), (515, 3), (129, 2), (260, 5)])
def test_protocol_on_the_planned_layout_reproduces_the_reference_factorisation(oracle, n, ngpus):
    a = oracle.fill((n, n), 1) - 0.25
    ref_lu, ref_piv, ref_sign = oracle.lu(a)
    lu, piv, sign = _walk_protocol(a, ngpus)
    assert np.array_equal(piv, ref_piv.astype(np.int64)) and sign == ref_sign
    assert np.max(np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), 1.0)) <= 1e-12 * n
