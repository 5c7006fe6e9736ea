#!/usr/bin/env python3
"""bench.py -- headline benchmark of the dense hot path (BASELINE.json): f64 GEMM @ n=8192 and f64 LU @ n=16384 on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1 : workload "f64 GEMM 8192x8192x8192" (BASELINE configs[1]); the LU n=16384 + 16-RHS solve (configs[2]) is timed
        in the same run and reported under "lu".  A step = one full GEMM over synthetic, HBM-resident inputs.
N > 1 : launched by torchrun, one rank per GPU: "f64 GEMM 32768^3 row-sharded, B broadcast with NCCL" (configs[3]);
        A and C are row-block sharded, rank 0 owns B and broadcasts it in K-panels that overlap the GEMM of the
        previous panel (C += A[:, panel] * B[panel, :]).  Total work is fixed => "scaling": "strong".
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
        d["source"] = "measured in this run: rust-la_b200/build/peak_fp64 (DMMA.8x8x4 issue-rate loop, all SMs)"
        return d
    except Exception as e:
        p = os.path.join(ROOT, "profiles", "peak_fp64_r1.json")
        if os.path.exists(p):
            d = json.load(open(p))
            d["source"] = f"committed profiles/peak_fp64_r1.json (in-run microbenchmark failed: {e})"
            return d
        return {"dmma_tflops": NOMINAL_FP64_TFLOPS, "dmma_tflops_sustained": NOMINAL_FP64_TFLOPS,
                "source": f"nominal 40 TFLOP/s (no measurement available: {e})"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-lu", action="store_true", help="N=1 only: do not time the LU n=16384 leg")
    ap.add_argument("--skip-f32", action="store_true", help="N=1 only: do not time the f32 65536x1024x16384 leg")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--n", type=int, default=0, help="override the GEMM size (debug)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from la import _cabi
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

    peak = measure_fp64_peak() if rank == 0 else None

    if n_gpus == 1:
        n = args.n or 8192
        m_loc, k, nn = n, n, n
        row0 = 0
    else:
        from la import sharding as _sh
        n = args.n or 32768
        row0, row1 = _sh.row_shard(n, n_gpus, rank)
        m_loc, k, nn = row1 - row0, n, n

    A = torch.empty((m_loc, k), dtype=f64, device=dev)
    B = torch.empty((k, nn), dtype=f64, device=dev)
    C = torch.empty((m_loc, nn), dtype=f64, device=dev)
    chk(L.la_fill_hash_f64_dev(A.data_ptr(), A.numel(), 1, row0 * k, sp))
    if rank == 0:
        chk(L.la_fill_hash_f64_dev(B.data_ptr(), B.numel(), 2, 0, sp))

    from la import sharding
    PANELS = 8 if n_gpus > 1 else 1
    plan = sharding.k_panels(k, PANELS)
    launches_per_step = len(plan)

    def gemm_panel(k0, k1, accumulate):
        chk(L.la_gemm_f64_dev(A.data_ptr() + k0 * 8, k, B.data_ptr() + k0 * nn * 8, nn, C.data_ptr(), nn,
                              m_loc, k1 - k0, nn, 2 if accumulate else 0, sp))

    def step():
        if n_gpus == 1:
            chk(L.la_gemm_f64_dev(A.data_ptr(), k, B.data_ptr(), nn, C.data_ptr(), nn, m_loc, k, nn, 0, sp))
            return
        # rank 0 owns B: NCCL broadcast in K-panels; the multiply of panel p overlaps the transfer of panel p+1
        sharding.sharded_gemm(A, B, C, k, PANELS, lambda rows: dist.broadcast(rows, src=0, async_op=True), gemm_panel)

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
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
    ms = e0.elapsed_time(e1)
    if n_gpus > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    flops = 2.0 * n * n * n
    value = flops / (ms_per_step * 1e-3) / 1e12

    # ---- end-to-end through the host-pointer C ABI (what `&a * &b` binds): pinned host buffers, H2D + kernel + D2H ----
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
        # the host path accumulates K in panels (C += A_p B_p): same products, partial sums rounded into C at panel
        # boundaries -> compare to the single-launch result with the f64 parity tolerance instead of bit-equality
        dC = hC.to(dev)
        rel = float(((dC - C).abs() / C.abs().clamp_min(1e-300)).max())
        e2e["max_rel_diff_vs_device_resident_result"] = rel
        e2e["matches_device_resident_result"] = bool(rel <= 1e-12)
        del dC
        del hA, hB, hC
    else:
        # multi-GPU e2e: host shards -> device, B from rank 0's host, result shard back to host
        hA = torch.empty((m_loc, k), dtype=f64).pin_memory()
        hC = torch.empty((m_loc, nn), dtype=f64).pin_memory()
        hA.copy_(A.cpu())
        hB = torch.empty((k, nn), dtype=f64).pin_memory() if rank == 0 else None
        if rank == 0:
            hB.copy_(B.cpu())
        barrier()
        t0 = time.perf_counter()
        A.copy_(hA, non_blocking=True)
        if rank == 0:
            B.copy_(hB, non_blocking=True)
        step()
        hC.copy_(C, non_blocking=True)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": flops / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": n * k * 8 + k * nn * 8,
               "d2h_bytes_per_step": n * nn * 8, "ms_per_step": dt * 1e3,
               "api": "pinned host shards -> la_gemm_f64_dev per rank + NCCL broadcast of B -> pinned host shards"}

    # ---- LU n=16384 + solve with 16 RHS (configs[2]), 1 GPU only ----
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
        for it in range(1 + 3):
            LU.copy_(A0)
            a0, a1, a2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a0.record(stream)
            chk(L.la_lu_factor_f64_dev(LU.data_ptr(), ln, ln, piv.data_ptr(), sign.data_ptr(), sp))
            a1.record(stream)
            chk(L.la_lu_solve_f64_dev(LU.data_ptr(), ln, piv.data_ptr(), Bx.data_ptr(), nx, X.data_ptr(), sp))
            a2.record(stream)
            torch.cuda.synchronize()
            if it > 0:
                lu_ms.append(a0.elapsed_time(a1))
                solve_ms.append(a1.elapsed_time(a2))
        lu_t = sum(lu_ms) / len(lu_ms)
        so_t = sum(solve_ms) / len(solve_ms)
        lu_flops = 2.0 / 3.0 * ln ** 3
        # residual of the solve as a size-independent sanity check: ||A x - b|| / (||A|| ||x||)
        R = torch.empty((ln, nx), dtype=f64, device=dev)
        chk(L.la_gemm_f64_dev(A0.data_ptr(), ln, X.data_ptr(), nx, R.data_ptr(), nx, ln, ln, nx, 0, sp))
        torch.cuda.synchronize()
        res = float((R - Bx).norm() / (A0.norm() * X.norm()))
        lu = {"workload": "f64 LU partial pivoting n=16384 + solve nx=16", "lu_ms": lu_t,
              "lu_tflops": lu_flops / (lu_t * 1e-3) / 1e12, "flops_formula": "2/3 n^3",
              "solve_ms": so_t, "solve_gbs": (8.0 * ln * ln * 2) / (so_t * 1e-3) / 1e9,
              "solve_residual": res}

    # ---- Cholesky n=16384 + solve with 16 RHS (widening step), 1 GPU only ----
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
        for it in range(1 + 2):
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

    # ---- f32 GEMM 65536x1024 x 1024x16384 (configs[4]) on the tcgen05 kind::tf32 kernel, 1 GPU only ----
    f32 = None
    if not args.skip_f32:
        A0 = LU = R = A = B = C = None  # release the f64 buffers
        torch.cuda.empty_cache()
        fm, fk, fn = 65536, 1024, 16384
        fm_loc = fm // n_gpus  # rows of A and C are sharded; rank 0 broadcasts B (64 MiB) every step
        f32t = torch.float32
        FA = torch.empty((fm_loc, fk), dtype=f32t, device=dev)
        FB = torch.empty((fk, fn), dtype=f32t, device=dev)
        FC = torch.empty((fm_loc, fn), dtype=f32t, device=dev)
        chk(L.la_fill_hash_f32_dev(FA.data_ptr(), FA.numel(), 1, rank * fm_loc * fk, sp))
        if rank == 0:
            chk(L.la_fill_hash_f32_dev(FB.data_ptr(), FB.numel(), 2, 0, sp))
        else:
            FB.zero_()

        def f32_step():
            if n_gpus > 1:
                dist.broadcast(FB, src=0)
            chk(L.la_gemm_f32_dev(FA.data_ptr(), fk, FB.data_ptr(), fn, FC.data_ptr(), fn, fm_loc, fk, fn, 0, sp))

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
        f_ms = f0.elapsed_time(f1) / freps
        if n_gpus > 1:
            tt = torch.tensor([f_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            f_ms = float(tt.item())
        rows = torch.arange(0, fm_loc, max(1, fm_loc // 64), device=dev)
        want = FA[rows].double() @ FB.double()
        rel = float(((FC[rows].double() - want).abs() / want.abs().clamp_min(1e-300)).max())
        tf32_peak, tf32_src = NOMINAL_TF32_TFLOPS, "nominal dense TF32 (half the nominal bf16 rate)"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tf32_peak = float(mp["bf16_tflops"]) / 2.0
            tf32_src = "MEASURED_PEAKS.json bf16_tflops / 2 (TF32 issues at half the bf16 rate; no TF32 entry in the file)"
        except Exception:
            pass
        f_tf = 2.0 * fm * fk * fn / (f_ms * 1e-3) / 1e12
        f32 = {"workload": "f32 GEMM 65536x1024 x 1024x16384" +
                           ("" if n_gpus == 1 else f", rows sharded over {n_gpus} GPUs, B broadcast by NCCL every step"),
               "ms": f_ms, "tflops": f_tf,
               "kernel": "gemm_f32_tf32_kernel (tcgen05.mma kind::tf32, TMEM accumulators, TMA in/out) + B transpose",
               "max_rel_err_vs_f64_on_64_rows": rel, "tolerance": 1e-4 * fk,
               "roofline": {"bound": "tensor", "achieved": f_tf / n_gpus, "peak": tf32_peak, "unit": "TFLOP/s",
                            "frac": f_tf / n_gpus / tf32_peak, "peak_source": tf32_src}}
        del FA, FB, FC

    if rank != 0:
        if n_gpus > 1:
            dist.destroy_process_group()
        return 0

    # the GEMM is timed alone over a sub-second region -> burst figure; the sustained one is reported beside it
    peak_tf = float(peak.get("dmma_tflops") or peak.get("dmma_tflops_sustained"))
    roofline = {"bound": "tensor", "achieved": value / n_gpus, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": value / n_gpus / peak_tf, "traffic": None,
                "kernel": "gemm_f64_tma_kernel (DMMA.8x8x4, TMA-fed)",
                "algorithmic": "2*m*n*k flops per launch (one launch per step per GPU)",
                "peak_source": peak.get("source"), "peak_sustained": peak.get("dmma_tflops_sustained"),
                "dfma_peak": peak.get("dfma_tflops"), "frac_of_nominal_40": value / n_gpus / NOMINAL_FP64_TFLOPS,
                "note": "MEASURED_PEAKS.json has no fp64 entry; the fp64 tensor (DMMA) issue-rate ceiling is measured "
                        "by csrc/peak_fp64.cu: 148 SMs x 128 flop/clk x 1.965 GHz = 37.2 TFLOP/s"}
    traffic_file = os.path.join(ROOT, "profiles", "gemm_f64_traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("dram_bytes_per_launch")
        except Exception:
            pass
    if chol:
        chol["frac_of_fp64_peak"] = chol["chol_tflops"] / peak_tf
    if lu:
        lu["frac_of_fp64_peak"] = lu["lu_tflops"] / peak_tf
        lu["frac_of_nominal_40"] = lu["lu_tflops"] / NOMINAL_FP64_TFLOPS

    cpu = None
    if n_gpus > 1:
        roofline["traffic"] = None  # the ncu capture is of the 1-GPU 8192^3 launch
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

    line = {
        "metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if n_gpus == 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"f64 GEMM {n}x{n}x{n}" + ("" if n_gpus == 1 else
                                                             f" row-sharded over {n_gpus} GPUs, B broadcast by NCCL in {PANELS} K-panels"),
                   "l2": "inputs (A,B,C = 3 x %d MiB per GPU) exceed the 126 MB L2; no explicit flush" % (m_loc * k * 8 >> 20),
                   "inputs": "counter-based splitmix64 hash, uniform [0,1), seeds A=1 B=2"},
        "pct_of_fp64_peak": 100.0 * value / n_gpus / peak_tf,
        "clocks": sampler.summary(),
        "e2e": e2e,
        "gpu_launches": launches_per_step * args.steps,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "lu": lu,
        "cholesky": chol,
        "f32": f32,
    }
    print(json.dumps(line))
    if n_gpus > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
