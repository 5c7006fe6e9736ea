"""One rank of the one-process-per-GPU multi-GPU Mul test (started by tests/test_gpu_mg.py, or by hand under torchrun):
contexts exchange their CUDA-IPC handles over gloo, then a device-resident and a host-shard product are checked against
the oracle on sampled rows.  Usage: RANK/WORLD_SIZE/LOCAL_RANK/MASTER_* in the environment, argv[1] = f64 | f32."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from la import _cabi, sharding
    from la._cabi import check, lib
    from oracle import oracle as orc
    from gpu_util import max_rel_err

    dtype = np.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else np.float32
    es = np.dtype(dtype).itemsize
    suf = "f64" if dtype == np.float64 else "f32"
    rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc.build()
    m, k, n = (4096, 2304, 4608) if dtype == np.float64 else (8192, 1024, 8192)
    ctx = sharding.MgContext(rank, world, dev, dtype, k, n)
    mine = torch.frombuffer(bytearray(ctx.handle()), dtype=torch.uint8)
    allh = [torch.empty(_cabi.LA_MG_HANDLE_BYTES, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(allh, mine)
    ctx.connect(b"".join(bytes(t.numpy().tobytes()) for t in allh))
    r0, r1, c0, c1 = sharding.shard(world, rank, m, n, es)
    tol = 1e-12 * k if dtype == np.float64 else 4e-6 + 1.2e-7 * k
    b = orc.fill((k, n), 2, dtype)
    rows = np.unique(np.concatenate([[r0, r1 - 1], np.random.default_rng(rank).integers(r0, r1, 30)]))

    def check_rows(c_shard, what):
        for r in rows:
            a_row = orc.fill((1, k), 1, dtype, first_idx=int(r) * k)
            err = max_rel_err(c_shard[r - r0:r - r0 + 1], orc.gemm_rows(a_row, b, 0, 1))
            assert err <= tol, f"{what}: row {r} error {err}"

    # ---- device-resident: A shard and the own column block generated on the device, two products back to back ----
    torch.cuda.set_device(dev)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    A = torch.empty((r1 - r0, k), dtype=tdt, device=f"cuda:{dev}")
    C = torch.full((r1 - r0, n), float("nan"), dtype=tdt, device=f"cuda:{dev}")
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    fill = getattr(lib(), f"la_fill_hash_{suf}_dev")
    check(fill(A.data_ptr(), A.numel(), 1, r0 * k, sp))
    ptr, ldb, bc0, bc1 = ctx.b_block()
    assert (bc0, bc1, ldb) == (c0, c1, n)
    for i in range(k):
        check(fill(ctypes.c_void_p(ptr + i * ldb * es), c1 - c0, 2, i * n + c0, sp))
    for _ in range(2):
        ctx.gemm(A.data_ptr(), k, C.data_ptr(), n, r1 - r0, sp)
    torch.cuda.synchronize()
    check_rows(C.cpu().numpy(), "device-resident")
    dist.barrier()
    # ---- host shards: two products (the second overwrites the column blocks, so the ack protocol is on the path) ----
    a_sh = orc.fill((r1 - r0, k), 1, dtype, first_idx=r0 * k)
    b_blk = np.ascontiguousarray(b[:, c0:c1])
    for _ in range(2):
        c_sh = np.full((r1 - r0, n), np.nan, dtype=dtype)
        ctx.gemm_host(a_sh, b_blk, c_sh)
        assert np.all(np.isfinite(c_sh))
        check_rows(c_sh, "host shards")
    dist.barrier()
    ctx.destroy()
    dist.destroy_process_group()
    print("MG_WORKER_OK", rank)


if __name__ == "__main__":
    main()
