#!/usr/bin/env python3
"""Small-shape calls of every kernel family, meant to run under compute-sanitizer (racecheck / synccheck / memcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Each result is still checked against the oracle, so a sanitizer run is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import la  # noqa: E402
from la import _cabi  # noqa: E402
from la._cabi import check, lib  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from gpu_util import DevBuf, gemm_dev, max_rel_err, sync  # noqa: E402

which = sys.argv[1:] or ["gemm64", "gemm32", "lu", "lu_mg", "solve", "chol", "qr"]
orc.build()
if "gemm64" in which:  # TMA + DMMA kernel, both tile configurations, ragged edges, all three epilogues
    for path, (m, k, n) in ((3, (256, 528, 320)), (4, (256, 528, 320)), (2, (130, 64, 260))):
        a, b, c0 = orc.fill((m, k), 1), orc.fill((k, n), 2), orc.fill((m, n), 4)
        ref = orc.gemm(a, b)
        for mode, want in ((0, ref), (1, c0 - ref), (2, c0 + ref)):
            da, db, dc = DevBuf.from_array(a), DevBuf.from_array(b), DevBuf.from_array(c0)
            check(lib().la_debug_set_gemm_path(path))
            gemm_dev(da, k, db, n, dc, n, m, k, n, mode, np.float64)
            sync()
            lib().la_debug_set_gemm_path(0)
            got = dc.to_array((m, n), np.float64)
            assert np.max(np.abs(got - want) / np.maximum(np.abs(want), np.abs(ref))) <= 1e-12 * k
    print("gemm64 ok")
if "gemm32" in which:  # tcgen05 kernel, plain TF32 and the 3-pass compensated mode
    m, k, n = 384, 160, 512
    a, b = orc.fill((m, k), 1, np.float32), orc.fill((k, n), 2, np.float32)
    ref = orc.gemm(a, b)
    for mode, tol in ((_cabi.LA_F32_TF32, 1e-4 * k), (_cabi.LA_F32_3XTF32, 4e-6 + 1.2e-7 * k)):
        da, db, dc = DevBuf.from_array(a), DevBuf.from_array(b), DevBuf(m * n * 4)
        check(lib().la_debug_set_gemm_f32_path(2))
        check(lib().la_set_gemm_f32_mode(mode))
        gemm_dev(da, k, db, n, dc, n, m, k, n, 0, np.float32)
        sync()
        lib().la_debug_set_gemm_f32_path(0)
        lib().la_set_gemm_f32_mode(_cabi.LA_F32_3XTF32)
        assert max_rel_err(dc.to_array((m, n), np.float32), ref) <= tol
    print("gemm32 ok")
if "lu" in which:  # single exact panel, multi-panel look-ahead pipeline (fp64 and fp32), wide and tall shapes
    for shape, dt in (((100, 100), np.float64), ((400, 400), np.float64), ((300, 520), np.float64), ((520, 300), np.float64),
                      ((400, 400), np.float32)):
        a = orc.fill(shape, 1, dt)
        ref_lu, ref_piv, ref_sign = orc.lu(a)
        dec = la.LUDecomposition.new(la.Matrix.from_numpy(a))
        assert np.array_equal(dec.get_piv(), ref_piv) and dec.pospivsign == ref_sign
        lu = dec.get_lu().to_numpy()
        tol = (1e-12 if dt == np.float64 else 1e-4) * max(shape)
        assert float(np.max(np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), 1.0))) <= tol
    print("lu ok")
if "lu_mg" in which:  # multi-device driver with ranks sharing device 0: shifted base pointers, ring slots, peer copies
    from la import sharding
    for n, world in ((300, 2), (515, 3), (129, 2)):
        a = orc.fill((n, n), 1)
        ref_lu, ref_piv, ref_sign = orc.lu(a)
        lu, piv, sign = sharding.lu_factor_mg(a, [0] * world)
        assert np.array_equal(piv, ref_piv) and sign == ref_sign
        assert float(np.max(np.abs(lu - ref_lu) / np.maximum(np.abs(ref_lu), 1.0))) <= 1e-12 * n
    print("lu_mg ok")
if "solve" in which:  # reference-order path (n < 512), persistent sweep kernels (nx <= 16), GEMM sweeps (many RHS)
    for n, nx in ((200, 3), (640, 16), (768, 5), (640, 64)):
        a, rhs = orc.fill((n, n), 1), orc.fill((n, nx), 3)
        dec = la.LUDecomposition.new(la.Matrix.from_numpy(a))
        x = dec.solve(la.Matrix.from_numpy(rhs)).to_numpy()
        assert np.linalg.norm(a @ x - rhs) / (np.linalg.norm(a) * np.linalg.norm(x)) <= 1e-13
    print("solve ok")
if "chol" in which:
    for n in (100, 400, 644):  # 644: several block columns -> the lower-triangle-only trailing GEMM; solve with 4 RHS
        g = orc.fill((n, n), 5)
        spd = orc.gemm(g, np.ascontiguousarray(g.T)) + n * np.eye(n)
        ch = la.CholeskyDecomposition.new(la.Matrix.from_numpy(spd))
        ref_l = orc.chol(spd)
        assert float(np.max(np.abs(ch.get_l().to_numpy() - ref_l)) / np.max(np.abs(ref_l))) <= 1e-12 * n
        if n > 512:  # triangular transpose + sweep kernels
            rhs = orc.fill((n, 4), 3)
            x = ch.solve(la.Matrix.from_numpy(rhs)).to_numpy()
            assert np.linalg.norm(spd @ x - rhs) / (np.linalg.norm(spd) * np.linalg.norm(x)) <= 1e-13
    print("chol ok")
if "qr" in which and hasattr(la, "QRDecomposition"):
    for shape in ((200, 120), (400, 400)):
        a = orc.fill(shape, 1)
        qr = la.QRDecomposition.new(la.Matrix.from_numpy(a))
        q, r = qr.get_q().to_numpy(), qr.get_r().to_numpy()
        assert np.linalg.norm(q @ r - a) / np.linalg.norm(a) <= 1e-13
    print("qr ok")
print("SANITIZE_SMALL_DONE")
