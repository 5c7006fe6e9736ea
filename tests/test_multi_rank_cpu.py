"""world_size-2 (and 3) gloo runs of the multi-GPU Mul plan on CPU: row-block shards, B broadcast in K-panels, accumulate per
panel.  The per-panel multiply is numpy here (the GPU run in bench.py plugs la_gemm_f64_dev into the same plan)."""
import os
import socket
import sys

import numpy as np
import pytest

from la import sharding


def test_row_shard_partition():
    for m, world in [(32768, 8), (10, 3), (7, 8), (1, 1), (65536, 8)]:
        spans = [sharding.row_shard(m, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == m
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_k_panels_cover_and_align():
    for k, p in [(32768, 8), (1024, 8), (100, 8), (17, 4), (16, 8), (5, 3)]:
        plan = sharding.k_panels(k, p)
        assert plan[0][0] == 0 and plan[-1][1] == k and len(plan) <= p
        assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))
        assert all((k0 % 16 == 0) for k0, _ in plan)


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
    r0, r1 = sh.row_shard(m, world, rank)
    a_shard = np.ascontiguousarray(a[r0:r1])
    b_local = torch.from_numpy(b.copy() if rank == 0 else np.zeros((k, n)))  # only the root holds B
    c_shard = np.full((r1 - r0, n), np.nan)

    def bcast(view):
        return dist.broadcast(view, src=0, async_op=True)

    def gemm_panel(k0, k1, accumulate):
        prod = a_shard[:, k0:k1] @ b_local.numpy()[k0:k1]
        if accumulate:
            c_shard[...] += prod
        else:
            c_shard[...] = prod

    sh.sharded_gemm(a_shard, b_local, c_shard, k, 4, bcast, gemm_panel)
    np.save(os.path.join(out_dir, f"c_{rank}.npy"), c_shard)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (64, 96, 40)), (3, (50, 33, 17))])
def test_sharded_gemm_gloo(tmp_path, world, shape):
    torch = pytest.importorskip("torch")
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
