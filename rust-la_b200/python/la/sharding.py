"""Host-side plan of the multi-GPU Mul: row-block shards of A and C, B replicated by a broadcast in K-panels.

Every row of C = A*B depends on the same row of A and on all of B (reference loop nest, src/matrix/mod.rs:965-973), so the
product shards by row blocks with no reduction.  Rank `root` owns B and broadcasts it in `panels` row blocks (contiguous in
row-major storage); each rank multiplies panel p -- C_shard (+)= A_shard[:, panel p] * B[panel p, :] -- while panel p+1 is still
in flight.  The functions here are pure / backend-agnostic so the same plan runs on NCCL + CUDA (bench.py) and on gloo + numpy
(tests/test_multi_rank_cpu.py)."""


def row_shard(m, world, rank):
    """Rows [r0, r1) of A and C owned by `rank`: as even as possible, earlier ranks take the remainder."""
    base, rem = divmod(m, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def k_panels(k, panels):
    """Split the inner dimension into at most `panels` contiguous row blocks of B, each a multiple of 16 (the GEMM k-tile)."""
    panels = max(1, min(panels, (k + 15) // 16))
    step = -(-k // panels)
    step = -(-step // 16) * 16
    out, k0 = [], 0
    while k0 < k:
        out.append((k0, min(k, k0 + step)))
        k0 += step
    return out


def sharded_gemm(a_shard, b_full, c_shard, k, panels, broadcast_async, gemm_panel):
    """Runs the pipeline on one rank.

    broadcast_async(b_rows_view) -> handle with .wait(): starts the broadcast of one K-panel of B (rows k0:k1)
    gemm_panel(k0, k1, accumulate): C_shard (+)= A_shard[:, k0:k1] * B[k0:k1, :]
    All broadcasts are enqueued up front (they serialise on the communication stream); the compute stream waits for panel
    p only, so the multiply of panel p overlaps the transfer of panel p+1.
    """
    plan = k_panels(k, panels)
    handles = [broadcast_async(b_full[k0:k1]) for (k0, k1) in plan]
    for i, (k0, k1) in enumerate(plan):
        handles[i].wait()
        gemm_panel(k0, k1, i > 0)
    return c_shard
