#!/usr/bin/env python3
"""bench.py -- headline benchmark of the dense hot path (BASELINE.json): f64 GEMM @ n=8192 and f64 LU @ n=16384 on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 : workload "f64 GEMM 8192x8192x8192" (BASELINE configs[1]); the LU n=16384 + 16-RHS solve (configs[2]) is timed
        in the same run and reported under "lu".  A step = one full GEMM over synthetic, HBM-resident inputs.
N > 1 : launched by torchrun, one rank per GPU: "f64 GEMM 32768^3 row-sharded" (configs[3]) through the library's own
        multi-GPU entry points (la_mg_* / la_gemm_f64_mg_rank[_host]): A and C are row-block sharded, every rank owns a
        column block of B and pulls the other blocks over NVLink with the library's peer-memory kernel while it multiplies
        by the blocks it already has.  NCCL (torch.distributed) only carries the 256-byte handles, the barriers and the
        max-over-ranks of the timings.  Total work is fixed => "scaling": "strong".  Every shard of the timed product is
        checked against the oracle on sampled rows x column windows.
--impl reference : the reference's own CPU loop order (oracle/la_oracle.c, canonical i-j-k nest) on the host cores, on a
        bounded sample of the same workload.  The reference is Rust and cannot be built in this image.

PyTorch is plumbing only here (device tensors, streams, events, torch.distributed); all arithmetic is the C-ABI library.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rust-la_b200", "python"))

METRIC = "f64 GEMM TFLOP/s @n=8192 & LU TFLOP/s @n=16384, % of B200 FP64 peak"
NOMINAL_TF32_TFLOPS = 1125.0  # half the nominal dense bf16 rate (2250); used only when nothing was measured
NOMINAL_FP64_TFLOPS = 40.0  # NVIDIA DGX B200 listing, vector == tensor; used only when nothing was measured


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (NVML; nvidia-smi may be absent)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, str(e)

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv:
            self._stop.clear()
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.nv or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------------------------
# CPU reference arm (the oracle's canonical loop nest == the reference's loop order and access pattern)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(n, rows, threads):
    """Times `rows` complete rows of the n^3 f64 product with the reference's i-j-k nest on `threads` host threads."""
    from oracle import oracle as orc
    orc.build()
    a = orc.fill((rows, n), 1)
    b = orc.fill((n, n), 2)
    t0 = time.perf_counter()
    orc.gemm_rows(a, b, 0, rows, form="canon", threads=threads)
    dt = time.perf_counter() - t0
    return 2.0 * rows * n * n / dt / 1e12, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as orc
    orc.build()
    n = 8192 if args.gpus == 1 else 32768
    threads = os.cpu_count() or 1
    # bounded sample: ~2 rows per thread of the same product (rows are independent and equal work)
    rows = max(threads, min(n, 2 * threads))
    if n > 8192:
        rows = threads  # 32768-wide rows are 16x the work each
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(n, max(1, threads // 2), threads)
    vals, times = [], []
    for _ in range(args.steps):
        v, dt = cpu_reference_sample(n, rows, threads)
        vals.append(v)
        times.append(dt)
    value = sum(vals) / len(vals)
    single, _ = cpu_reference_sample(n, 1, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"f64 GEMM {n}x{n}x{n}", "sample": f"{rows} of {n} output rows per step"},
        "cpu_baseline": {"value": value, "unit": "TFLOP/s", "cores": threads, "kind": "port",
                         "sample": f"{rows} of {n} rows of C per step, reference i-j-k loop order (strided walk of B), "
                                   f"rows spread over {threads} threads by us; the reference itself is single-threaded",
                         "single_thread_value": single},
        "e2e": {"value": value, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is a Rust crate and cannot be compiled in this image; this is the oracle's restatement of "
                "src/matrix/mod.rs:965-973 (gcc -O3 -ffp-contract=off)",
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def measure_fp64_peak():
    """Runs the DMMA/DFMA issue-rate microbenchmark (a few seconds) and returns its JSON, or a committed/nominal fallback."""
    exe = os.path.join(ROOT, "rust-la_b200", "build", "peak_fp64")
    out = os.path.join("/tmp", f"peak_fp64_{os.getpid()}.json")
    try:
        subprocess.run([exe, out], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=120)
        d = json.load(open(out))
        d["source"] = "measured in this run by the builder's own microbenchmark rust-la_b200/build/peak_fp64 (DMMA.8x8x4 " \
                      "issue-rate loop on all SMs); there is no driver-side fp64 anchor in MEASURED_PEAKS.json"
        return d
    except Exception as e:
        p = os.path.join(ROOT, "profiles", "peak_fp64_r1.json")
        if os.path.exists(p):
            d = json.load(open(p))
            d["source"] = f"committed profiles/peak_fp64_r1.json (in-run microbenchmark failed: {e})"
            return d
        return {"dmma_tflops": NOMINAL_FP64_TFLOPS, "dmma_tflops_sustained": NOMINAL_FP64_TFLOPS,
                "source": f"nominal 40 TFLOP/s (no measurement available: {e})"}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def window_check(orc, get_c_rows, row0, row1, n, k, seed_a, seed_b, dtype, col_windows, nrows, tol, rng_seed):
    """Parity of a (sharded) product on sampled rows x column windows against the oracle.  Every element of C = A*B is an
    independent dot product in the reference's loop nest (src/matrix/mod.rs:965-973), so a sample of elements is an exact
    check; windows let the checker rebuild pieces of a 32768 x 32768 B without the 8 GiB whole.  Returns the worst
    relative error and the number of elements compared."""
    import numpy as np
    rng = np.random.default_rng(rng_seed)
    rows = np.unique(np.concatenate([[row0, row1 - 1], rng.integers(row0, row1, max(0, nrows - 2))]))
    a_rows = np.concatenate([orc.fill((1, k), seed_a, dtype, first_idx=int(r) * k) for r in rows], axis=0)
    got_rows = get_c_rows(rows - row0)  # [len(rows), n] host array
    worst, count = 0.0, 0
    for (c0, c1) in col_windows:
        b_win = orc.fill_block(0, k, c0, c1 - c0, n, seed_b, dtype)
        ref = orc.gemm(a_rows, b_win).astype(np.float64)
        got = got_rows[:, c0:c1].astype(np.float64)
        worst = max(worst, float(np.max(np.abs(got - ref) / np.abs(ref))))
        count += ref.size
    return worst, count


def column_windows(n, nranks, rank, elem, width=256):
    """Windows that touch every column range a rank multiplies separately (own block, right of it, left of it) and
    straddle the block boundaries."""
    from la import sharding
    _, _, c0, c1 = sharding.shard(nranks, rank, 128, n, elem)
    cand = [0, n - width, c0, max(0, c0 - width // 2), max(0, c1 - width // 2), (c0 + c1) // 2]
    out = []
    for c in cand:
        c = int(min(max(c, 0), n - width)) // 16 * 16
        if (c, c + width) not in out:
            out.append((c, c + width))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-lu", action="store_true", help="N=1 only: do not time the LU n=16384 / Cholesky legs")
    ap.add_argument("--skip-f32", action="store_true", help="do not time the f32 65536x1024x16384 leg")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--skip-lu-mg", action="store_true", help="N>1 only: do not time the LU n=16384 across the N GPUs")
    ap.add_argument("--skip-check", action="store_true", help="skip the in-bench oracle checks (debug)")
    ap.add_argument("--n", type=int, default=0, help="override the GEMM size (debug)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from la import _cabi, sharding
    L = _cabi.lib()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    n_gpus = world

    def chk(st):
        if st != 0:
            raise RuntimeError(L.la_last_error().decode())

    stream = torch.cuda.current_stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    f64 = torch.float64
    warmup = max(args.warmup, 3)
    orc = None
    if not (args.skip_check and args.skip_cpu):
        from oracle import oracle as orc  # the checker; never on a timed GPU path
        orc.build()

    peak = measure_fp64_peak() if rank == 0 else None
    mp = measured_peaks()

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if n_gpus == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_mg(dtype, k, nn):
        """One context per rank; the handles travel through torch.distributed (plumbing), the data never does."""
        ctx = sharding.MgContext(rank, n_gpus, local_rank, dtype, k, nn)
        mine = torch.frombuffer(bytearray(ctx.handle()), dtype=torch.uint8).to(dev)
        allh = [torch.empty(_cabi.LA_MG_HANDLE_BYTES, dtype=torch.uint8, device=dev) for _ in range(n_gpus)]
        dist.all_gather(allh, mine)
        ctx.connect(b"".join(t.cpu().numpy().tobytes() for t in allh))
        return ctx

    def fill_own_block(ctx, fill_fn, es, seed, k, nn):
        """The rank's column block of B generated in place in its replica: element (i, j) is hash(seed, i * n + j)."""
        ptr, ldb, c0, c1 = ctx.b_block()
        for i in range(k):
            chk(fill_fn(ptr + i * ldb * es, c1 - c0, seed, i * nn + c0, sp))
        torch.cuda.synchronize()
        return c0, c1

    # =================================================================================================================
    # f64 GEMM: 8192^3 on one GPU (configs[1]) / 32768^3 row-sharded over N GPUs through la_gemm_f64_mg_rank (configs[3])
    # =================================================================================================================
    mg = None
    if n_gpus == 1:
        n = args.n or 8192
        m_loc, k, nn = n, n, n
        row0 = 0
        A = torch.empty((m_loc, k), dtype=f64, device=dev)
        B = torch.empty((k, nn), dtype=f64, device=dev)
        C = torch.empty((m_loc, nn), dtype=f64, device=dev)
        chk(L.la_fill_hash_f64_dev(A.data_ptr(), A.numel(), 1, 0, sp))
        chk(L.la_fill_hash_f64_dev(B.data_ptr(), B.numel(), 2, 0, sp))
        launches_per_step = 1

        def step():
            chk(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), nn, C.data_ptr(), nn, m_loc, k, nn, 0, sp))
    else:
        n = args.n or 32768
        k, nn = n, n
        row0, row1, bc0, bc1 = sharding.shard(n_gpus, rank, n, nn, 8)
        m_loc = row1 - row0
        mg = make_mg(np.float64, k, nn)
        A = torch.empty((m_loc, k), dtype=f64, device=dev)
        C = torch.empty((m_loc, nn), dtype=f64, device=dev)
        chk(L.la_fill_hash_f64_dev(A.data_ptr(), A.numel(), 1, row0 * k, sp))
        fill_own_block(mg, L.la_fill_hash_f64_dev, 8, 2, k, nn)
        # per rank and step: publish + (N-1) pulls + (N-1) acks + up to 3 GEMMs (own block, right of it, left of it)
        launches_per_step = 1 + 2 * (n_gpus - 1) + (1 + (1 if bc1 < nn else 0) + (1 if bc0 > 0 else 0))

        def step():
            mg.gemm(A.data_ptr(), k, C.data_ptr(), nn, m_loc, sp)

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.stop()
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    flops = 2.0 * n * n * n
    value = flops / (ms_per_step * 1e-3) / 1e12

    # ---- parity of the timed product (N > 1: every shard; N = 1 is covered by tests/test_gpu_gemm_parity.py at full
    #      size, re-checked here on the same sample) ----
    parity = None
    if not args.skip_check:
        wins = column_windows(nn, n_gpus, rank, 8)
        worst, cnt = window_check(orc, lambda rr: C[torch.as_tensor(rr, device=dev)].cpu().numpy(), row0, row0 + m_loc, nn, k,
                                  1, 2, np.float64, wins, 64, 1e-12 * k, 100 + rank)
        worst = max_over_ranks(worst)
        parity = {"checked": f"64 sampled rows x {len(wins)} column windows of 256 per shard vs the oracle "
                             f"(every element is an independent dot product)", "elements_per_shard": cnt,
                  "max_rel_err": worst, "tolerance": 1e-12 * k, "ok": bool(worst <= 1e-12 * k)}
        if not parity["ok"]:
            raise SystemExit(f"bench.py: the timed product FAILED parity: {parity}")

    # ---- end-to-end through the host-pointer C ABI: H2D + kernels + D2H inside the timed region ----
    e2e = None
    if n_gpus == 1:
        hA = torch.empty((n, n), dtype=f64).pin_memory()
        hB = torch.empty((n, n), dtype=f64).pin_memory()
        hC = torch.empty((n, n), dtype=f64).pin_memory()
        hA.copy_(A.cpu())
        hB.copy_(B.cpu())
        for _ in range(2):
            chk(L.la_gemm_f64_host(hA.data_ptr(), hB.data_ptr(), hC.data_ptr(), n, n, n))
        torch.cuda.synchronize()
        reps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(reps):
            chk(L.la_gemm_f64_host(hA.data_ptr(), hB.data_ptr(), hC.data_ptr(), n, n, n))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        e2e = {"value": flops / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": 2 * n * n * 8,
               "d2h_bytes_per_step": n * n * 8, "ms_per_step": dt * 1e3,
               "api": "la_gemm_f64_host (pinned host A,B,C; K-panel then row-block pipelined H2D / DMMA kernel / D2H)"}
        chk(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), nn, C.data_ptr(), nn, m_loc, k, nn, 0, sp))
        torch.cuda.synchronize()
        dC = hC.to(dev)
        rel = float(((dC - C).abs() / C.abs().clamp_min(1e-300)).max())
        e2e["max_rel_diff_vs_device_resident_result"] = rel
        e2e["matches_device_resident_result"] = bool(rel <= 1e-12)
        del dC
        # the same call on PAGEABLE host memory (what a plain Vec<f64> / numpy array is): the driver stages every copy
        pA, pB, pC = hA.numpy().copy(), hB.numpy().copy(), np.empty((n, n))
        chk(L.la_gemm_f64_host(pA.ctypes.data, pB.ctypes.data, pC.ctypes.data, n, n, n))
        t0 = time.perf_counter()
        for _ in range(2):
            chk(L.la_gemm_f64_host(pA.ctypes.data, pB.ctypes.data, pC.ctypes.data, n, n, n))
        dtp = (time.perf_counter() - t0) / 2
        e2e["pageable"] = {"value": flops / dtp / 1e12, "unit": "TFLOP/s", "ms_per_step": dtp * 1e3,
                           "note": "same la_gemm_f64_host call on pageable (malloc) host memory"}
        del hA, hB, hC, pA, pB, pC
    else:
        # every rank: its rows of A, ITS column block of B and its rows of C in pinned host memory; one collective call
        hA = torch.empty((m_loc, k), dtype=f64).pin_memory()
        hC = torch.empty((m_loc, nn), dtype=f64).pin_memory()
        hB = torch.empty((k, bc1 - bc0), dtype=f64).pin_memory()
        hA.copy_(A.cpu())
        hB.copy_(torch.from_numpy(orc.fill_block(0, k, bc0, bc1 - bc0, nn, 2)) if orc is not None else torch.rand((k, bc1 - bc0), dtype=f64))
        e2e_ms = []
        for it in range(3):  # first is the warm-up (scratch allocation)
            barrier()
            t0 = time.perf_counter()
            chk(L.la_gemm_f64_mg_rank_host(mg.h, hA.data_ptr(), hB.data_ptr(), bc1 - bc0, hC.data_ptr(), m_loc))
            barrier()
            if it > 0:
                e2e_ms.append(max_over_ranks((time.perf_counter() - t0) * 1e3))
        dt = sum(e2e_ms) / len(e2e_ms) * 1e-3
        e2e = {"value": flops / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": n * k * 8 + k * nn * 8,
               "d2h_bytes_per_step": n * nn * 8, "ms_per_step": dt * 1e3,
               "api": "la_gemm_f64_mg_rank_host per rank: pinned host rows of A + the rank's 1/N column block of B up its own "
                      "PCIe link, the other blocks pulled over NVLink, rows of C back down (bytes are the whole job's)"}
        if not args.skip_check:
            wins = column_windows(nn, n_gpus, rank, 8)
            worst, _ = window_check(orc, lambda rr: hC[torch.as_tensor(rr)].numpy(), row0, row0 + m_loc, nn, k, 1, 2,
                                    np.float64, wins, 16, 1e-12 * k, 200 + rank)
            worst = max_over_ranks(worst)
            e2e["max_rel_err_vs_oracle"] = worst
            if worst > 1e-12 * k:
                raise SystemExit(f"bench.py: the end-to-end product FAILED parity: {worst}")
        del hA, hB, hC
    if mg is not None:
        barrier()
        mg.destroy()
        mg = None

    # =================================================================================================================
    # LU n=16384 + solve with 16 RHS (configs[2]) and the Cholesky widening step: 1 GPU only
    # =================================================================================================================
    lu = None
    if n_gpus == 1 and not args.skip_lu:
        ln, nx = 16384, 16
        del A, B, C
        torch.cuda.empty_cache()
        A0 = torch.empty((ln, ln), dtype=f64, device=dev)
        LU = torch.empty((ln, ln), dtype=f64, device=dev)
        piv = torch.empty((ln,), dtype=torch.int64, device=dev)
        sign = torch.empty((1,), dtype=torch.int32, device=dev)
        Bx = torch.empty((ln, nx), dtype=f64, device=dev)
        X = torch.empty((ln, nx), dtype=f64, device=dev)
        chk(L.la_fill_hash_f64_dev(A0.data_ptr(), A0.numel(), 1, 0, sp))
        chk(L.la_fill_hash_f64_dev(Bx.data_ptr(), Bx.numel(), 3, 0, sp))
        lu_ms, solve_ms = [], []
        lu_iters = max(10, args.steps)
        for it in range(2 + lu_iters):
            LU.copy_(A0)
            a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a0.record(stream)
            chk(L.la_lu_factor_f64_dev(LU.data_ptr(), ln, ln, piv.data_ptr(), sign.data_ptr(), sp))
            a1.record(stream)
            chk(L.la_lu_solve_f64_dev(LU.data_ptr(), ln, piv.data_ptr(), Bx.data_ptr(), nx, X.data_ptr(), sp))
            a2.record(stream)
            torch.cuda.synchronize()
            if it > 1:
                lu_ms.append(a0.elapsed_time(a1))
                solve_ms.append(a1.elapsed_time(a2))
        lu_t = sum(lu_ms) / len(lu_ms)
        so_t = sum(solve_ms) / len(solve_ms)
        lu_flops = 2.0 / 3.0 * ln ** 3
        R = torch.empty((ln, nx), dtype=f64, device=dev)
        chk(L.la_gemm_f64_dev(A0.data_ptr(), ln, X.data_ptr(), nx, R.data_ptr(), nx, ln, ln, nx, 0, sp))
        torch.cuda.synchronize()
        res = float((R - Bx).norm() / (A0.norm() * X.norm()))
        solve_bytes = 8.0 * ln * ln + 2 * 8.0 * ln * nx  # LU read once (each sweep reads its half) + B in, X out
        hbm_peak = float(mp.get("hbm_gbs", 6650.0))
        lu = {"workload": "f64 LU partial pivoting n=16384 + solve nx=16", "lu_ms": lu_t, "lu_ms_min": min(lu_ms),
              "lu_iters": len(lu_ms), "lu_tflops": lu_flops / (lu_t * 1e-3) / 1e12, "flops_formula": "2/3 n^3",
              "solve_ms": so_t, "solve_ms_min": min(solve_ms), "solve_residual": res,
              "solve_roofline": {"bound": "hbm", "achieved": solve_bytes / (so_t * 1e-3) / 1e9, "peak": hbm_peak,
                                 "unit": "GB/s", "frac": solve_bytes / (so_t * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes": solve_bytes,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in mp else "fallback 6650 GB/s"}}
        del R

    chol = None
    if n_gpus == 1 and not args.skip_lu:
        cn, cnx = 16384, 16
        A0 = LU = R = A = B = C = None  # release the earlier legs' buffers
        torch.cuda.empty_cache()
        G0 = torch.empty((cn, cn), dtype=f64, device=dev)
        chk(L.la_fill_hash_f64_dev(G0.data_ptr(), G0.numel(), 1, 0, sp))
        S0 = G0 @ G0.T  # input generation only (torch): A = G G' + n I, symmetrised exactly
        S0 = (S0 + S0.T) * 0.5
        S0.diagonal().add_(float(cn))
        del G0
        CL = torch.empty_like(S0)
        CBm = torch.empty((cn, cnx), dtype=f64, device=dev)
        CX = torch.empty((cn, cnx), dtype=f64, device=dev)
        cflags = torch.zeros((2,), dtype=torch.int32, device=dev)
        chk(L.la_fill_hash_f64_dev(CBm.data_ptr(), CBm.numel(), 3, 0, sp))
        c_ms, cs_ms = [], []
        for it in range(1 + 3):
            CL.copy_(S0)
            c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            c0.record(stream)
            chk(L.la_chol_factor_f64_dev(CL.data_ptr(), cn, cflags.data_ptr(), sp))
            c1.record(stream)
            chk(L.la_chol_solve_f64_dev(CL.data_ptr(), cn, CBm.data_ptr(), cnx, CX.data_ptr(), sp))
            c2.record(stream)
            torch.cuda.synchronize()
            if it > 0:
                c_ms.append(c0.elapsed_time(c1))
                cs_ms.append(c1.elapsed_time(c2))
        c_t = sum(c_ms) / len(c_ms)
        fl = cflags.cpu().tolist()
        chol = {"workload": "f64 Cholesky n=16384 (A = G G' + n I) + solve nx=16", "ok": int(fl[0] == 0 and fl[1] == 0),
                "chol_ms": c_t, "chol_tflops": cn ** 3 / 3.0 / (c_t * 1e-3) / 1e12, "flops_formula": "1/3 n^3",
                "solve_ms": sum(cs_ms) / len(cs_ms),
                "backward_error": float((CL @ CL.T - S0).norm() / S0.norm()),
                "solve_residual": float((S0 @ CX - CBm).norm() / (S0.norm() * CX.norm()))}
        del S0, CL, CBm, CX

    # =================================================================================================================
    # QR n = 16384 and the streaming elementwise / norm kernels (rank 4), one GPU
    # =================================================================================================================
    qr_leg = None
    stream_leg = None
    if n_gpus == 1 and not args.skip_lu:
        qn = 16384
        A0 = LU = R = A = B = C = None
        torch.cuda.empty_cache()
        QA0 = torch.empty((qn, qn), dtype=f64, device=dev)
        chk(L.la_fill_hash_f64_dev(QA0.data_ptr(), QA0.numel(), 1, 0, sp))
        QRm = torch.empty_like(QA0)
        te = ctypes.c_size_t(0)
        chk(L.la_qr_tmat_elems(qn, qn, local_rank, 8, ctypes.byref(te)))
        Qrd = torch.empty((qn,), dtype=f64, device=dev)
        Qtm = torch.empty((te.value,), dtype=f64, device=dev)
        q_ms = []
        for it in range(1 + 3):
            QRm.copy_(QA0)
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record(stream)
            chk(L.la_qr_factor_f64_dev(QRm.data_ptr(), qn, qn, Qrd.data_ptr(), Qtm.data_ptr(), sp))
            q1.record(stream)
            torch.cuda.synchronize()
            if it > 0:
                q_ms.append(q0.elapsed_time(q1))
        q_t = sum(q_ms) / len(q_ms)
        # R'R = A'A (Q orthogonal): a backward-error style check that needs neither Q nor the oracle at this size
        Rm = torch.triu(QRm, 1)
        Rm.diagonal().copy_(Qrd)
        qerr = float((Rm.T @ Rm - QA0.T @ QA0).norm() / (QA0.T @ QA0).norm())
        q_flops = 4.0 / 3.0 * qn ** 3
        qr_leg = {"workload": "f64 QR (Householder, compact WY) n=16384", "qr_ms": q_t, "qr_ms_min": min(q_ms),
                  "qr_tflops": q_flops / (q_t * 1e-3) / 1e12, "flops_formula": "4/3 n^3",
                  "rtr_minus_ata_rel": qerr}
        del Rm, QRm, Qtm
        # streaming kernels: C = A + B over 2 x 2 GiB in, 2 GiB out; ||A||_F over 2 GiB
        SB = torch.empty_like(QA0)
        SC = torch.empty_like(QA0)
        chk(L.la_fill_hash_f64_dev(SB.data_ptr(), SB.numel(), 2, 0, sp))
        cnt = QA0.numel()
        for _ in range(2):
            chk(L.la_elementwise_f64_dev(_cabi.LA_EW_ADD, QA0.data_ptr(), SB.data_ptr(), 0.0, SC.data_ptr(), cnt, sp))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(stream)
        for _ in range(5):
            chk(L.la_elementwise_f64_dev(_cabi.LA_EW_ADD, QA0.data_ptr(), SB.data_ptr(), 0.0, SC.data_ptr(), cnt, sp))
        s1.record(stream)
        torch.cuda.synchronize()
        add_ms = s0.elapsed_time(s1) / 5
        nrm = ctypes.c_double(0)
        chk(L.la_reduce_f64_dev(_cabi.LA_RED_SUMSQ, QA0.data_ptr(), None, cnt, ctypes.byref(nrm), sp))
        t0 = time.perf_counter()
        for _ in range(5):
            chk(L.la_reduce_f64_dev(_cabi.LA_RED_SUMSQ, QA0.data_ptr(), None, cnt, ctypes.byref(nrm), sp))
        nrm_ms = (time.perf_counter() - t0) * 1e3 / 5
        hbm_peak = float(mp.get("hbm_gbs", 6650.0))
        stream_leg = {"workload": "f64 elementwise add and Frobenius norm over 16384 x 16384 operands",
                      "add_ms": add_ms, "add_gbs": 3 * 8.0 * cnt / (add_ms * 1e-3) / 1e9,
                      "add_frac_of_hbm": 3 * 8.0 * cnt / (add_ms * 1e-3) / 1e9 / hbm_peak,
                      "add_bit_identical_to_torch": bool(torch.equal(SC, QA0 + SB)),
                      "frobenius_ms_incl_d2h_sync": nrm_ms, "frobenius_gbs": 8.0 * cnt / (nrm_ms * 1e-3) / 1e9,
                      "frobenius_rel_diff_vs_torch": abs(nrm.value - float(QA0.norm())) / float(QA0.norm())}
        del QA0, SB, SC

    # =================================================================================================================
    # f32 GEMM 65536x1024 x 1024x16384 (configs[4]): tcgen05 kind::tf32 (the mode the config names) and the default 3xTF32
    # =================================================================================================================
    f32 = None
    if not args.skip_f32:
        A0 = LU = R = A = B = C = None
        torch.cuda.empty_cache()
        fm, fk, fn = 65536, 1024, 16384
        f32t = torch.float32
        fr0, fr1, fc0, fc1 = sharding.shard(n_gpus, rank, fm, fn, 4)
        fm_loc = fr1 - fr0
        FA = torch.empty((fm_loc, fk), dtype=f32t, device=dev)
        FC = torch.empty((fm_loc, fn), dtype=f32t, device=dev)
        chk(L.la_fill_hash_f32_dev(FA.data_ptr(), FA.numel(), 1, fr0 * fk, sp))
        fmg = None
        if n_gpus == 1:
            FB = torch.empty((fk, fn), dtype=f32t, device=dev)
            chk(L.la_fill_hash_f32_dev(FB.data_ptr(), FB.numel(), 2, 0, sp))

            def f32_step():
                chk(L.la_gemm_f32_dev(FA.data_ptr(), fk, FB.data_ptr(), fn, FC.data_ptr(), fn, fm_loc, fk, fn, 0, sp))
        else:
            fmg = make_mg(np.float32, fk, fn)
            fill_own_block(fmg, L.la_fill_hash_f32_dev, 4, 2, fk, fn)

            def f32_step():
                fmg.gemm(FA.data_ptr(), fk, FC.data_ptr(), fn, fm_loc, sp)

        def time_f32(mode):
            chk(L.la_set_gemm_f32_mode(mode))
            for _ in range(3):
                f32_step()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            freps = 10
            f0.record(stream)
            for _ in range(freps):
                f32_step()
            f1.record(stream)
            barrier()
            ms = max_over_ranks(f0.elapsed_time(f1) / freps)
            err = None
            if not args.skip_check:  # 64 sampled full rows per shard against the fp32 oracle (B is only 64 MiB)
                bb = orc.fill((fk, fn), 2, np.float32)
                rows = np.unique(np.concatenate([[0, fm_loc - 1], np.random.default_rng(300 + rank).integers(0, fm_loc, 62)]))
                got = FC[torch.as_tensor(rows, device=dev)].cpu().numpy().astype(np.float64)
                aa = np.concatenate([orc.fill((1, fk), 1, np.float32, first_idx=int(fr0 + r) * fk) for r in rows], axis=0)
                ref = orc.gemm(aa, bb).astype(np.float64)
                err = max_over_ranks(float(np.max(np.abs(got - ref) / np.abs(ref))))
            return ms, err

        tf_ms, tf_err = time_f32(_cabi.LA_F32_TF32)
        x3_ms, x3_err = time_f32(_cabi.LA_F32_3XTF32)
        chk(L.la_set_gemm_f32_mode(_cabi.LA_F32_3XTF32))
        if not args.skip_check:
            if tf_err > 1e-4 * fk or x3_err > 4e-6 + 1.2e-7 * fk:
                raise SystemExit(f"bench.py: the f32 product FAILED parity: tf32 {tf_err}, 3xtf32 {x3_err}")
        tf32_peak, tf32_src = NOMINAL_TF32_TFLOPS, "nominal dense TF32 (half the nominal bf16 rate)"
        if "bf16_tflops" in mp:
            tf32_peak = float(mp["bf16_tflops"]) / 2.0
            tf32_src = "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 issues at half the bf16 rate; no TF32 entry in the file)"
        fl32 = 2.0 * fm * fk * fn
        f_tf = fl32 / (tf_ms * 1e-3) / 1e12
        f32 = {"workload": "f32 GEMM 65536x1024 x 1024x16384" +
                           ("" if n_gpus == 1 else f", rows sharded over {n_gpus} GPUs through la_gemm_f32_mg_rank (column blocks "
                                                   f"of B pulled over NVLink every step)"),
               "ms": tf_ms, "tflops": f_tf, "mode": "LA_F32_TF32 (opt-in; the mode BASELINE's config names)",
               "kernel": "gemm_f32_tf32_pair_kernel (tcgen05.mma cta_group::2 kind::tf32 on CTA pairs, TMEM accumulators, TMA in/out) + B transpose",
               "max_rel_err_vs_fp32_oracle_on_64_rows_per_shard": tf_err, "tolerance": 1e-4 * fk,
               "default_mode_3xtf32": {"ms": x3_ms, "tflops": fl32 / (x3_ms * 1e-3) / 1e12,
                                       "max_rel_err_vs_fp32_oracle_on_64_rows_per_shard": x3_err,
                                       "tolerance": 4e-6 + 1.2e-7 * fk,
                                       "note": "library default: split-compensated, fp32-grade (3 tensor passes)"},
               "roofline": {"bound": "tensor", "achieved": f_tf / n_gpus, "peak": tf32_peak, "unit": "TFLOP/s",
                            "frac": f_tf / n_gpus / tf32_peak, "frac_of_nominal_1125": f_tf / n_gpus / NOMINAL_TF32_TFLOPS,
                            "peak_source": tf32_src}}
        if fmg is not None:
            barrier()
            fmg.destroy()
        del FA, FC

    if rank != 0:
        if n_gpus > 1:
            dist.destroy_process_group()
        return 0

    # the GEMM is timed alone over a sub-second region -> burst figure; the sustained one is reported beside it
    peak_tf = float(peak.get("dmma_tflops") or peak.get("dmma_tflops_sustained"))
    roofline = {"bound": "tensor", "achieved": value / n_gpus, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": value / n_gpus / peak_tf, "traffic": None,
                "kernel": "gemm_f64_tma_kernel (DMMA.8x8x4, TMA-fed)",
                "algorithmic": "2*m*n*k flops per GPU and step",
                "peak_source": peak.get("source"), "peak_sustained": peak.get("dmma_tflops_sustained"),
                "dfma_peak": peak.get("dfma_tflops"), "peak_sm_max_mhz": peak.get("sm_max_mhz"),
                "denominators": {"measured_dmma": peak_tf, "nominal_40": NOMINAL_FP64_TFLOPS, "guide_45": 45.0},
                "frac_of_nominal_40": value / n_gpus / NOMINAL_FP64_TFLOPS, "frac_of_guide_45": value / n_gpus / 45.0,
                "note": "MEASURED_PEAKS.json has no fp64 entry; the fp64 tensor (DMMA) issue-rate ceiling is measured "
                        "by csrc/peak_fp64.cu: 148 SMs x 128 flop/clk x 1.965 GHz = 37.2 TFLOP/s"}
    traffic_file = os.path.join(ROOT, "profiles", "gemm_f64_traffic.json")
    if os.path.exists(traffic_file) and n_gpus == 1:
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        except Exception:
            pass
    if chol:
        chol["frac_of_fp64_peak"] = chol["chol_tflops"] / peak_tf
    if lu:
        lu["frac_of_fp64_peak"] = lu["lu_tflops"] / peak_tf
        lu["frac_of_nominal_40"] = lu["lu_tflops"] / NOMINAL_FP64_TFLOPS
        lu["frac_of_guide_45"] = lu["lu_tflops"] / 45.0

    cpu = None
    if not args.skip_cpu and n_gpus == 1:  # the CPU baseline is a rank-0, N=1 figure
        threads = os.cpu_count() or 1
        rows = min(8192, 2 * threads)
        v, dt = cpu_reference_sample(8192, rows, threads)
        v1, dt1 = cpu_reference_sample(8192, 2, 1)
        cpu = {"value": v, "unit": "TFLOP/s", "cores": threads, "kind": "port",
               "sample": f"{rows} of 8192 output rows of the 8192^3 product ({dt:.1f} s), reference i-j-k loop order, "
                         f"rows spread over {threads} threads by us; single-thread (the reference as shipped): "
                         f"{v1 * 1e3:.3f} GFLOP/s on 2 rows ({dt1:.1f} s)",
               "single_thread_value": v1}
        if lu:
            cpu["lu"] = cpu_lu_baseline(threads)

    lu_mg = None
    if n_gpus > 1 and not args.skip_lu_mg:
        lu_mg = lu_mg_leg(n_gpus)

    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if n_gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"f64 GEMM {n}x{n}x{n}" + ("" if n_gpus == 1 else
                                                             f" row-sharded over {n_gpus} GPUs, column blocks of B pulled over NVLink "
                                                             f"(la_gemm_f64_mg_rank)"),
                   "l2": "inputs (A,B,C = %d + %d + %d MiB per GPU) exceed the 126 MB L2; no explicit flush" %
                         (m_loc * k * 8 >> 20, k * nn * 8 >> 20, m_loc * nn * 8 >> 20),
                   "inputs": "counter-based splitmix64 hash, uniform [0,1), seeds A=1 B=2"},
        "pct_of_fp64_peak": 100.0 * value / n_gpus / peak_tf,
        "clocks": sampler.summary(),
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "parity": parity,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "lu": lu,
        "lu_mg": lu_mg,
        "cholesky": chol,
        "qr": qr_leg,
        "streaming": stream_leg,
        "f32": f32,
    }
    print(json.dumps(line))
    if n_gpus > 1:
        dist.destroy_process_group()
    return 0


def lu_mg_leg(n_gpus):
    """LU n = 16384 across the N GPUs (la_lu_mg_*: one host thread drives all devices, so it runs once, from rank 0, after the
    other ranks have finished -- in a child process with a hard timeout so that nothing here can cost the GEMM line)."""
    devs = ",".join(str(i) for i in range(n_gpus))
    cmd = [sys.executable, os.path.join(ROOT, "tools", "lu_mg_profile.py"), "16384", "4", "0", devs, "--check"]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}
    res = {}
    for ln in out.stdout.splitlines():
        if not ln.startswith("lu_mg "):
            continue
        key = "one_device" if "devices=[0]:" in ln else "all_devices"
        best = float(ln.split("best ")[1].split(" ms")[0])
        res[key] = {"ms": best, "tflops": 2.0 / 3.0 * 16384 ** 3 / (best * 1e-3) / 1e12,
                    "piv_identical_to_oracle_fixture": "piv_identical_to_oracle_fixture=True" in ln}
    if "all_devices" not in res:
        return {"error": (out.stdout + out.stderr)[-400:]}
    return {"workload": f"f64 LU n=16384 over {n_gpus} GPUs: 128-column blocks dealt round-robin, the panel owner's block "
                        f"column copied to every device (la_lu_mg_factor_f64), best of 4, device-timed on the first device "
                        f"which waits for all others", "n_gpus": n_gpus, **res,
            "speedup_vs_one_device_same_driver": res.get("one_device", {}).get("ms", 0.0) / res["all_devices"]["ms"],
            "bound": "the panel chain (panel -> peer copy -> head -> update -> panel), which hops from device to device"}


def cpu_lu_baseline(threads):
    """the reference's LU loop nest (src/decomp/lu.rs:116-161) timed on the host: the literal
    single-threaded nest at n = 1024 and 2048, the order-preserving row-parallel form on all cores at n = 4096, and the n^3
    extrapolation to 16384 -- flagged as generous to the reference, whose rate falls with n (stride-n inner reads)."""
    from oracle import oracle as orc
    out = {"unit": "GFLOP/s", "flops_formula": "2/3 n^3", "kind": "port"}
    for n, form, key in ((1024, "canon", "n1024_1core"), (2048, "canon", "n2048_1core"), (4096, "fast", f"n4096_{threads}threads")):
        a = orc.fill((n, n), 1)
        t0 = time.perf_counter()
        orc.lu(a, form=form)
        dt = time.perf_counter() - t0
        out[key] = {"seconds": dt, "gflops": 2.0 / 3.0 * n ** 3 / dt / 1e9}
    r1 = out["n2048_1core"]["gflops"]
    rp = out[f"n4096_{threads}threads"]["gflops"]
    fl = 2.0 / 3.0 * 16384 ** 3 / 1e9
    out["extrapolated_n16384_seconds"] = {"1core_at_n2048_rate": fl / r1, f"{threads}threads_at_n4096_rate": fl / rp,
                                          "note": "n^3 extrapolation at the measured rate: generous to the reference"}
    return out


if __name__ == "__main__":
    sys.exit(main())
