"""world_size-2 (and 3) gloo runs of the multi-GPU Mul plan on CPU: row-block shards of A and C, every rank owns a column
block of B and fetches the others' (gloo all_gather here; NVLink pulls inside libla_b200.so on the GPU), own block first.
The partition comes from the C library itself (la_mg_shard is pure arithmetic, no device needed); the per-range multiply
is numpy here."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest

from la import _cabi, sharding


def test_shard_partition_covers_and_aligns():
    for m, n, world, elem in [(32768, 32768, 8, 8), (65536, 16384, 8, 4), (1000, 520, 3, 8), (129, 36, 2, 4), (1, 2, 1, 8),
                              (4096, 4100, 4, 8), (300, 16, 8, 8)]:
        parts = [sharding.shard(world, r, m, n, elem) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == m and parts[0][2] == 0 and parts[-1][3] == n
        for a, b in zip(parts, parts[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        for r0, r1, c0, c1 in parts:
            assert r0 % 128 == 0 or r0 == m
            assert (c0 * elem) % 16 == 0  # column blocks start on 16-byte boundaries (TMA / 16-byte pulls)
            if n // world >= 256:
                assert c0 % 256 == 0
    # the BASELINE shapes split evenly
    assert [sharding.shard(8, r, 32768, 32768)[1] - sharding.shard(8, r, 32768, 32768)[0] for r in range(8)] == [4096] * 8
    assert [sharding.shard(8, r, 65536, 16384, 4)[3] - sharding.shard(8, r, 65536, 16384, 4)[2] for r in range(8)] == [2048] * 8


def test_column_ranges_order():
    for world in (1, 2, 3, 8):
        for rank in range(world):
            rg = sharding.column_ranges(world, rank, 4096 * world)
            assert rg[0][2] == [rank]  # own block first: no transfer needed before the first multiply
            covered = sorted((c0, c1) for c0, c1, _ in rg)
            assert covered[0][0] == 0 and covered[-1][1] == 4096 * world
            assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
            owners = sorted(q for _, _, qs in rg for q in qs)
            assert owners == list(range(world))


def test_mg_entry_points_fail_loudly_without_device():
    if _cabi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    L = _cabi.lib()
    h = ctypes.c_void_p()
    assert L.la_mg_create(0, 2, 0, 8, 64, 64, ctypes.byref(h)) == _cabi.LA_ERR_NO_DEVICE
    a = np.ones((256, 8))
    b = np.ones((8, 256))
    c = np.zeros((256, 256))
    devs = (ctypes.c_int * 2)(0, 1)
    st = L.la_gemm_f64_mg(2, devs, a.ctypes.data, b.ctypes.data, c.ctypes.data, 256, 8, 256)
    assert st == _cabi.LA_ERR_NO_DEVICE and np.all(c == 0)


def _worker(rank, world, port, m, k, n, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "rust-la_b200", "python"))
    from la import sharding as sh
    rng = np.random.default_rng(123)
    a = rng.random((m, k))
    b = rng.random((k, n))
    r0, r1, c0, c1 = sh.shard(world, rank, m, n)
    a_shard = np.ascontiguousarray(a[r0:r1])
    b_own = np.ascontiguousarray(b[:, c0:c1])  # a rank holds ONLY its own column block of B
    c_shard = np.full((r1 - r0, n), np.nan)
    # the handle exchange of the GPU path, with stand-in bytes: every rank ends up with all handles in rank order
    mine = torch.full((256,), rank, dtype=torch.uint8)
    allh = [torch.empty(256, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(allh, mine)
    assert [int(t[0]) for t in allh] == list(range(world))
    # the transport: blocks are padded to the widest block so that all_gather works on ragged partitions
    widths = [sh.shard(world, q, m, n)[3] - sh.shard(world, q, m, n)[2] for q in range(world)]
    wmax = max(widths)
    send = torch.zeros((k, wmax), dtype=torch.float64)
    send[:, : c1 - c0] = torch.from_numpy(b_own)
    recv = [torch.empty((k, wmax), dtype=torch.float64) for _ in range(world)]
    dist.all_gather(recv, send)

    def fetch_block(q):
        return recv[q].numpy()[:, : widths[q]]

    def multiply(col0, col1, b_cols):
        assert b_cols.shape == (k, col1 - col0)
        c_shard[:, col0:col1] = a_shard @ b_cols

    sh.gather_blocks_and_multiply(a_shard, b_own, c_shard, world, rank, n, fetch_block, multiply)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), c_shard)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (300, 96, 40)), (3, (520, 33, 18)), (2, (256, 64, 1024))])
def test_sharded_gemm_gloo(tmp_path, world, shape):
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    m, k, n = shape
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, m, k, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(123)
    a = rng.random((m, k))
    b = rng.random((k, n))
    got = np.concatenate([np.load(tmp_path / f"c_{r}.npy") for r in range(world)], axis=0)
    assert got.shape == (m, n)
    assert np.allclose(got, a @ b, rtol=1e-13, atol=0)
