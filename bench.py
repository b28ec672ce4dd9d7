#!/usr/bin/env python
"""bench.py — LSTM fwd + backprop + SGD iterations/sec (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c5|c4] [--product tf32|f32]

One "step" = one pass of the `minimize` loop body (src/optimize.rs:64-71). The workload of the line is BASELINE
config 5 at every N (so the N = 1, 2, 4, 8 series is one strong-scaling curve): LSTM H = I = 4096, T = 32, global batch
8192 rows, f32 parameters, dag mode, TF32 products. N > 1 shards the batch rows over the ranks; every weight gradient is
summed over the ranks before its update by ONE fused kernel per weight — inside the NVSwitch (multimem.ld_reduce /
multimem.st over a symmetric multicast heap) from three ranks up, over NVLink peer memory at two (ncclAllReduce only
when neither is available).

  value      : iterations/s with every input resident in HBM, CUDA events on the library stream, max over ranks
  e2e        : the same metric through the public call `f.minimize(&args, lr, 1)` (C ABI cg_minimize) with HOST buffers:
               H2D of the T inputs + target from pinned memory and D2H of the root value inside the timed region
  roofline   : the dominant kernel (the [B x 4096] . [4096 x 4096] product), timed live with CUDA events
  parity     : N > 1 only, BEFORE anything is timed: the data-parallel run against the same kernels on one GPU and
               against the CPU oracle (f32 and TF32 shapes, every weight through the fused exchange); a failure ends
               the run with a non-zero exit code
  secondary  : N = 1 only: BASELINE config 4 (TF32, FP32-accurate 3 x TF32 on the tensor core, and FP32 SIMT products) and
               configs 1-3, each timed the same way
  same_config: N = 1 only: workloads the REFERENCE can run (README LSTM sweep, exact mode; a square LSTM in dag mode),
               GPU and reference CPU both measured in this run — no extrapolation
  peaks      : TF32 (cuBLAS through torch.matmul; best-of-10 bursts, and ours vs cuBLAS SUSTAINED at the power cap) and
               FP32-FMA peaks measured on this box in this run
  cpu_baseline / --impl reference : the reference's own CPU arithmetic (oracle/_ref, compiled unmodified from the
               reference's src/cpu) under the restated walker on a bounded square sample, 1 host core
"""
from __future__ import annotations

import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# before torch creates the CUDA context: enough hardware work queues that the input copies of an end-to-end call do not
# wait behind kernels of the iteration graph (computational-graph_b200/csrc/graph.cpp)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

H = I = 4096
LR = 0.01
GLOBAL_BATCH_C5 = 8192
METRIC = "LSTM fwd+backprop+SGD iterations/sec"


class Workload:
    """One BASELINE LSTM configuration for `world` ranks."""

    def __init__(self, name: str, world: int):
        self.name, self.world = name, world
        if name == "c5":
            if GLOBAL_BATCH_C5 % world:
                raise SystemExit(f"config 5: the global batch of {GLOBAL_BATCH_C5} rows does not divide over {world} ranks")
            self.T, self.rows, self.scaling, self.seed = 32, GLOBAL_BATCH_C5 // world, "strong", 5
        else:
            self.T, self.rows, self.scaling, self.seed = 8, 1024, "weak", 4
        # (46 (T-1) + 30) B H^2 flop per iteration in dag mode with the dead dX of the constant inputs skipped
        self.flop_per_gpu = (46 * (self.T - 1) + 30) * self.rows * H * H
        self.flop_global = self.flop_per_gpu * world

    @property
    def aggregate(self) -> int:
        """Units of the metric one step of the whole job completes."""
        return 1 if self.scaling == "strong" else self.world

    def config(self) -> dict:
        num = "5" if self.name == "c5" else "4"
        aggregate = ("value = global iterations/s: the 8192-row batch is sharded over the ranks, one step is one iteration "
                     "of the whole batch (strong scaling)" if self.name == "c5" else
                     "value = n_gpus x (iterations/s): each GPU runs the 1024-row LSTM iteration of its batch shard")
        return {"workload": f"BASELINE config {num}: LSTM H=I={H}, T={self.T}, global batch {self.rows * self.world} rows "
                            f"sharded over the ranks, f32 params, (lstm(inputs)-target).sq(), lr={LR}, dag mode",
                "params_13H2": 13 * H * H, "global_batch": self.rows * self.world, "seq_len": self.T,
                "aggregate": aggregate,
                "l2": "per-iteration working set (tens of GB) exceeds the 126 MB L2; no explicit flush"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "sm_max_mhz": 1965.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            c = [x.strip() for x in l.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------
def build_lstm(api, batch, d_in, d_hidden, steps, seed, with_inputs, shard=0):
    """(lstm(inputs) - target).sq() of src/main.rs:177-191 / src/lstm.rs:57-77 with seeded values; inputs and target
    are `Function::input` nodes when `with_inputs` (fed per call through `args`), else constants. `shard` > 0 draws
    a different batch shard (inputs, target) under the SAME parameters: data-parallel replicas start identical."""
    from computational_graph_b200.lstm import lstm_loss, lstm_params
    rng = np.random.default_rng(seed)
    params = lstm_params(d_in, d_hidden, batch, rng, scale=1.0 / np.sqrt(d_hidden))
    if shard:
        rng = np.random.default_rng(seed + 1000 * shard)
    xs = [rng.uniform(-.1, .1, (batch, d_in)).astype(np.float32) for _ in range(steps)]
    tgt = rng.uniform(-.1, .1, (batch, d_hidden)).astype(np.float32)
    if with_inputs:
        fx = [api.Function.input(f"x{t}", [batch, d_in]) for t in range(steps)]
        ft = api.Function.input("target", [batch, d_hidden])
        f = lstm_loss(api, fx, ft, params)
        args = {f"x{t}": xs[t] for t in range(steps)}
        args["target"] = tgt
        return f, args
    f = lstm_loss(api, [api.Constant(x) for x in xs], api.Constant(tgt), params)
    return f, {}


def profile_iteration(api, f):
    """cg_profile_iteration: one eager pass with a CUDA event between every kernel (per-kind device time, and the
    average duration of the largest product for the roofline)."""
    kind_ms = (C.c_double * 4)()
    kind_cnt = (C.c_ulonglong * 4)()
    gflop, gms, gcnt = C.c_double(), C.c_double(), C.c_ulonglong()
    fn = api.lib.cg_profile_iteration
    fn.argtypes = [C.c_void_p, C.c_float, C.c_double * 4, C.c_ulonglong * 4, C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.POINTER(C.c_ulonglong)]
    for _ in range(2):
        api._check(fn(f._h, LR, kind_ms, kind_cnt, C.byref(gflop), C.byref(gms), C.byref(gcnt)))
    return list(kind_ms), [int(x) for x in kind_cnt], gflop.value, gms.value, int(gcnt.value)


def measure_lstm(api, torch, dist, wl: Workload, product: str, steps: int, warmup: int, rank: int, local_rank: int,
                 measured: dict):
    """Builds the workload's graph on this rank and measures it: device-resident, end to end, per-kernel."""
    world = wl.world
    api.set_mode("dag")
    api.set_tf32({"f32": 0, "tf32": 1, "tf32x3": 2}[product])
    f, host_args = build_lstm(api, wl.rows, I, H, wl.T, wl.seed, with_inputs=True, shard=rank)
    pinned, keep = {}, []
    for k, v in host_args.items():
        t = torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = v
        keep.append(t)
        pinned[k] = api.Constant(t.numpy())  # contiguous f32: wraps the pinned memory without copying
    del host_args
    stats = api.plan_stats(f, LR)
    f.minimize(pinned, LR, 1)  # uploads the inputs once; also the first warm-up iteration

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    api.minimize_async(f, LR, warmup)
    barrier()
    launches0 = api.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = api.time_iterations(f, LR, 0, steps)
        barrier()
        launches = api.launch_count() - launches0
        # end to end through the public call, host buffers (same clock record)
        h2d = sum(int(np.asarray(v.value).nbytes) for v in pinned.values())
        d2h = wl.rows * H * 4
        for _ in range(2):
            f.minimize(pinned, LR, 1, trace=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            f.minimize(pinned, LR, 1, trace=True)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1000.0 / steps
    clocks = clk.summary()
    kind_ms, kind_cnt, gflop, gms, gcnt = profile_iteration(api, f)
    if dist is not None:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    pk = peaks()
    if product == "tf32":
        peak = pk["bf16_sustained"] / 2.0
        peak_note = (f"1/2 x {pk['source']} sustained cuBLAS bf16 ({pk['bf16_sustained']} TFLOP/s): tcgen05 kind::tf32 issues "
                     "at half the kind::f16 rate; the TF32 cuBLAS rate measured in this run is under `peaks`")
        bound = "tensor"
    elif product == "tf32x3":
        peak = pk["bf16_sustained"] / 6.0
        peak_note = (f"1/6 x {pk['source']} sustained cuBLAS bf16 ({pk['bf16_sustained']} TFLOP/s): FP32-accurate products on the "
                     "tensor core issue three tcgen05 kind::tf32 MMAs per k-step (operands split in shared memory into the part TF32 "
                     "keeps and the part it drops), so the ceiling is a third of the TF32 rate; flops counted once (2 m n k)")
        bound = "tensor"
    else:
        nominal = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        peak = measured.get("fp32_fma_tflops") or nominal
        peak_note = ("FP32 FMA pipe, measured in this run with a register-resident FFMA loop (tools/probes/fma_peak.cu)"
                     if measured.get("fp32_fma_tflops") else
                     f"FP32 FMA pipe, nominal 148 SMs x 128 lanes x 2 flop x {pk['sm_max_mhz']} MHz (probe library not built)")
        bound = "fma"
    achieved = gflop / (gms * 1e-3) / 1e12 if gms > 0 else 0.0
    agg = wl.aggregate
    cfg = wl.config()
    cfg["product"] = product
    cfg["parallelism"] = f"dp{world}" if world > 1 else "single"
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant product, from the committed `ncu --set full`
    # captures (per launch, like `achieved`)
    traffic = {("tf32", 1024): 88.86e6,      # profiles/r01_gemm_tf32_full.raw.txt: 83.9 MB read + 5.0 MB written
               ("tf32", 8192): 737.3e6,      # profiles/r02d_gemm_pair_full.raw.txt: 617 MB + 120 MB (335 MB algorithmic: B re-read per row-block wave)
               ("tf32x3", 8192): 1027.6e6,   # profiles/r02l_gemm_tf32x3_full.raw.txt: 900 MB + 128 MB
               }.get((product, wl.rows))
    line = {
        "metric": METRIC, "value": agg * 1000.0 / ms, "unit": "iterations/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None,
        "dtype": product, "data": "synthetic", "config": cfg,
        "e2e": {"value": agg * 1000.0 / e2e_ms, "unit": "iterations/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "kernel": f"largest product [{wl.rows} x {I}] . [{I} x {H}] (and its two transposed forms), {gcnt} "
                               f"launches/iter, {gflop / 1e9:.1f} GFLOP each, {gms * 1e3:.1f} us avg (CUDA events around "
                               "each launch, eager pass on the library stream)",
                     "peak_source": peak_note,
                     # every product the plan actually launches (dead products of the last step's output gate are
                     # eliminated) over the whole step; `survey_model` is SURVEY §8d's closed form
                     "whole_step": {"tflops": stats["gemm_flop_per_iter"] / (ms * 1e-3) / 1e12,
                                    "frac": stats["gemm_flop_per_iter"] / (ms * 1e-3) / 1e12 / peak,
                                    "flop_per_iter": stats["gemm_flop_per_iter"],
                                    "survey_model_flop_per_iter": wl.flop_per_gpu}},
        "kernel_share": {"elementwise_ms": kind_ms[0], "reduce_ms": kind_ms[1], "product_ms": kind_ms[2],
                         "exchange_ms_serialised": kind_ms[3], "instrumented_total_ms": sum(kind_ms),
                         "kernels": {"elementwise": kind_cnt[0], "reduce": kind_cnt[1], "product": kind_cnt[2],
                                     "exchange": kind_cnt[3]}},
        "plan": stats,
    }
    if measured.get("tf32_cublas") and product == "tf32":
        sus = (measured.get("tf32_sustained") or {}).get("cases", {})
        line["roofline"]["vs_cublas_tf32"] = {
            "in_step_tflops": achieved,
            "sustained_same_shapes": sus if measured.get("tf32_sustained", {}).get("how", "").find(f"rows = {wl.rows}") >= 0 else None,
            "cublas_burst_best_of_10": {k: v for k, v in measured["tf32_cublas"].items() if f"{wl.rows}x" in k or f"x{wl.rows}" in k},
            "cublas_burst_8192_cubed": measured["tf32_cublas"].get("8192^3"),
            "note": "a dense product runs against the 1 kW power cap: bursts (clocks not yet down) and sustained rates differ by ~20 %; "
                    "the like-for-like comparison is `sustained_same_shapes`"}
    del f, pinned, keep
    gc.collect()
    return line


# ---------------------------------------------------------------------------------------------------------
# peaks measured on this box, in this run
# ---------------------------------------------------------------------------------------------------------
def sustained_vs_cublas(torch, api, rows, seconds=0.8):
    """A dense product on this part runs against the power cap, so a best-of-N burst says little: each of the three hot
    products (forward X.W, dX = E.W^T, dW = X^T.E; src/optimize.rs:135,197-198) runs back to back for `seconds` through
    the library's 26-symbol `gemm` (tcgen05 TF32 kernel) and through cuBLAS TF32 (torch.matmul), the second half timed."""
    from matrix_abi import ProductMatrixAbi
    p = ProductMatrixAbi(api.lib)
    api.set_tf32(True)
    rng = np.random.default_rng(0)
    X = p.new(rows, I, rng.uniform(-1, 1, (rows, I)).astype(np.float32))
    W = p.new(I, H, rng.uniform(-1, 1, (I, H)).astype(np.float32))
    E = p.new(rows, H, rng.uniform(-1, 1, (rows, H)).astype(np.float32))
    Y, dX, dW = p.new(rows, H), p.new(rows, I), p.new(I, H)
    tx, tw, te = torch.randn(rows, I, device="cuda"), torch.randn(I, H, device="cuda"), torch.randn(rows, H, device="cuda")
    flop = 2.0 * rows * I * H

    def run(fn, sync):
        fn(); sync()
        t0, n = time.perf_counter(), 0
        while time.perf_counter() - t0 < seconds / 2:
            for _ in range(10):
                fn()
            sync()
            n += 10
        t1 = time.perf_counter()
        for _ in range(n):
            fn()
        sync()
        return flop / ((time.perf_counter() - t1) / n) / 1e12

    out = {"how": f"each product back to back for {seconds} s, second half timed; rows = {rows}; TFLOP/s", "cases": {}}
    cases = [("NN forward X.W", (X, False, W, False, Y), lambda: torch.matmul(tx, tw)),
             ("NT dX = E.W^T", (E, False, W, True, dX), lambda: torch.matmul(te, tw.t())),
             ("TN dW = X^T.E", (X, True, E, False, dW), lambda: torch.matmul(tx.t(), te))]
    for name, (a, ta, b, tb, r), cub in cases:
        ours = run(lambda: p.lib.gemm(C.byref(a), ta, C.byref(b), tb, C.byref(r)), api.synchronize)
        ref = run(cub, torch.cuda.synchronize)
        out["cases"][name] = {"ours": ours, "cublas": ref, "ours_over_cublas": ours / ref}
    for m in (X, W, E, Y, dX, dW):
        p.free(m)
    del tx, tw, te
    torch.cuda.empty_cache()
    return out


def measure_peaks(torch, api, rows_list):
    out = {}
    # FP32 FMA pipe: register-resident FFMA loop (tools/probes/fma_peak.cu, its own small library)
    lib_path = os.path.join(ROOT, "tools", "probes", "libcg_probes.so")
    if os.path.exists(lib_path):
        lib = C.CDLL(lib_path)
        lib.cg_probe_fma_tflops.restype, lib.cg_probe_fma_tflops.argtypes = C.c_double, [C.c_int, C.c_int]
        v = lib.cg_probe_fma_tflops(1 << 15, 5)
        lib.cg_probe_fma2_tflops.restype, lib.cg_probe_fma2_tflops.argtypes = C.c_double, [C.c_int, C.c_int]
        v2 = lib.cg_probe_fma2_tflops(1 << 15, 5)
        if v > 0:
            # the denominator is whichever instruction form reaches the higher rate on this box
            out["fp32_fma_tflops"] = max(v, v2)
            out["fp32_ffma_scalar_tflops"], out["fp32_ffma2_packed_tflops"] = v, v2
            out["fp32_fma_how"] = ("16 independent FMA chains per thread (scalar FFMA, and packed FFMA2 = __ffma2_rn), 2 x 1024 threads per SM, "
                                   "best of 5 launches each, CUDA events; the larger of the two is the peak")
    # TF32 tensor pipe as the vendor library reaches it: torch.matmul with allow_tf32 (cuBLAS), best of 10
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    res = {}

    def timed(a, b, ta, tb, flop, reps=10):
        x, y = (a.t() if ta else a), (b.t() if tb else b)
        for _ in range(3):
            torch.matmul(x, y)
        best = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(x, y)
            e1.record()
            e1.synchronize()
            best = max(best, flop / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        return best
    try:
        a = torch.randn(8192, 8192, device="cuda")
        b = torch.randn(8192, 8192, device="cuda")
        res["8192^3"] = timed(a, b, False, False, 2.0 * 8192 ** 3)
        del a, b
        for rows in rows_list:
            x = torch.randn(rows, I, device="cuda")
            w = torch.randn(I, H, device="cuda")
            e = torch.randn(rows, H, device="cuda")
            flop = 2.0 * rows * I * H
            res[f"NN_{rows}x{H}x{I}"] = timed(x, w, False, False, flop)     # X . W      (forward)
            res[f"NT_{rows}x{I}x{H}"] = timed(e, w, False, True, flop)      # E . W^T    (dX)
            res[f"TN_{I}x{H}x{rows}"] = timed(x, e, True, False, flop)      # X^T . E    (dW)
            del x, w, e
        out["tf32_cublas"] = res
        out["tf32_cublas_how"] = "torch.matmul, torch.backends.cuda.matmul.allow_tf32 = True, best of 10, CUDA events (TFLOP/s)"
        out["tf32_sustained"] = sustained_vs_cublas(torch, api, max(rows_list))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------------------
# parity of the data-parallel path, checked inside the bench run (N > 1), before anything is timed
# ---------------------------------------------------------------------------------------------------------
def nccl_allreduce_reference(torch, dist, world, n_floats=I * H, reps=20):
    """What the LIBRARY takes for the same exchange: `ncclAllReduce` (through torch.distributed) of one weight gradient
    (4096 x 4096 f32, 64 MiB), back to back, no compute beside it — the figure the fused exchange kernel's serialised time
    per weight (`kernel_share.exchange_ms_serialised` / kernels.exchange) stands against. The fused kernel also applies the
    SGD update, which NCCL leaves to a further pass over the weight."""
    g = torch.ones(n_floats, dtype=torch.float32, device="cuda")
    for _ in range(5):
        dist.all_reduce(g)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dist.all_reduce(g)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    nbytes = 4.0 * n_floats
    del g
    return {"what": "ncclAllReduce (torch.distributed, NCCL %s) of one 64 MiB weight gradient, %d back to back, max over ranks"
                    % (".".join(str(x) for x in torch.cuda.nccl.version()), reps),
            "us": ms * 1e3, "algbw_GBps": nbytes / (ms * 1e-3) / 1e9,
            "busbw_GBps": nbytes / (ms * 1e-3) / 1e9 * 2.0 * (world - 1) / world}


def dp_parity(api, pkg, torch, dist, rank, world, local_rank, multimem, tf32_batch=None, scale_lr=True):
    """The sharded run with every weight through the fused exchange, against (a) the same kernels on ONE GPU over the
    full batch and (b) the CPU oracle (dag mode). One f32 shape (1e-5) and one TF32 shape (1e-3; against the f32 oracle
    only the first iteration is held to it — TF32 rounding is amplified from iteration to iteration on one GPU as on
    several). Per-element criterion |a - b| <= tol * |b| + tol * rms(b) (tests/helpers.py close_err)."""
    from computational_graph_b200.dp import ROW_LOCAL, row_block, shard_lstm_problem
    from computational_graph_b200.lstm import lstm_loss, lstm_params
    from helpers import close_err
    from oracle.oracle import get_oracle

    # The reference's rules make the weight update grow with the batch (root error == 1, no normalisation, and Q3 hands
    # every gate the local derivative): at lr = .01 a 2048-row batch moves a 1/16-scale weight by several times its size
    # per iteration, and the 1e-7 that a different summation order leaves behind is amplified a thousandfold within
    # three iterations — on one GPU exactly as on eight. The check is about the EXCHANGE, so the learning rate shrinks
    # with the batch (same dynamics at every world size), and the first iteration — where both runs start from identical
    # weights and differ by the exchange alone — is held to the tolerance separately, gradients included.
    tb = tf32_batch or 256 * world
    cases = [{"name": "f32", "tf32": 0, "batch": 16 * world, "hidden": 48, "inp": 40, "T": 3, "iters": 4, "tol": 1e-5, "lr": LR},
             {"name": "tf32", "tf32": 1, "batch": tb, "hidden": 256, "inp": 256, "T": 3, "iters": 3, "tol": 1e-3,
              "lr": LR * 256.0 / tb if scale_lr else LR}]
    out = {"cases": [], "criterion": "max over elements of |a-b| / (tol*|b| + tol*rms(b)) <= 1"}
    api.set_mode("dag")
    for c in cases:
        pkg.dp.init_from_torch_distributed(api, local_rank)
        api.dp_set_fused_min(0)  # every weight goes through the fused exchange, however small
        if multimem != "auto":
            api.dp_set_multimem(2 if multimem == "on" and api.dp_multimem_supported() else 0)
        api.set_tf32(bool(c["tf32"]))
        rng = np.random.default_rng(99)
        params = lstm_params(c["inp"], c["hidden"], c["batch"], rng, scale=1.0 / np.sqrt(c["hidden"]))
        xs = [rng.uniform(-.1, .1, (c["batch"], c["inp"])).astype(np.float32) for _ in range(c["T"])]
        tgt = rng.uniform(-.1, .1, (c["batch"], c["hidden"])).astype(np.float32)
        p_r, x_r, t_r = shard_lstm_problem(params, xs, tgt, rank, world)
        f = lstm_loss(api, [api.Constant(x) for x in x_r], api.Constant(t_r), p_r)
        lr = c["lr"]
        st = api.plan_stats(f, lr)
        exchange = api.dp_exchange_name()
        tr1 = np.array(f.minimize({}, lr, 1, trace=True))
        dp_first = [(n, np.array(v)) for n, v in f.leaves()]
        tr = np.concatenate([tr1, np.array(f.minimize({}, lr, c["iters"] - 1, trace=True))])
        dp_leaves = [(n, np.array(v)) for n, v in f.leaves()]
        del f
        api.dp_shutdown()  # the communicator is gone: the next plan has no exchange in it
        api.dp_set_fused_min(-1)
        fs = lstm_loss(api, [api.Constant(x) for x in xs], api.Constant(tgt), params)
        ts1 = np.array(fs.minimize({}, lr, 1, trace=True))
        single_first = [(n, np.array(v)) for n, v in fs.leaves()]
        ts = np.concatenate([ts1, np.array(fs.minimize({}, lr, c["iters"] - 1, trace=True))])
        single_leaves = [(n, np.array(v)) for n, v in fs.leaves()]
        del fs
        orc = get_oracle()
        orc.use_port()
        orc.set_mode("dag")
        fo = lstm_loss(orc, [orc.Constant(x) for x in xs], orc.Constant(tgt), params)
        to = np.asarray(fo.minimize({}, lr, c["iters"], trace=True))
        orc.set_mode("exact")
        lo, hi = row_block(c["batch"], rank, world)
        tol = c["tol"]
        # first iteration: forward value and every leaf's update (p0 - p1) / lr, the exchanged gradients among them
        first = close_err(tr1, ts1[:, lo:hi, :], tol)
        for (n, vg), (ns, vs) in zip(dp_first, single_first):
            p0 = params[n][lo:hi] if n in ROW_LOCAL else params[n]
            want = vs[lo:hi] if n in ROW_LOCAL else vs
            gg, gw = (p0.astype(np.float64) - vg) / lr, (p0.astype(np.float64) - want) / lr
            noise = 4 * np.finfo(np.float32).eps * np.maximum(np.abs(p0), 1e-30) / lr
            bound = tol * np.abs(gw) + tol * max(float(np.sqrt(np.mean(gw * gw))), 1e-30) + noise
            first = max(first, float(np.max(np.abs(gg - gw) / bound)))
        vs_single = close_err(tr, ts[:, lo:hi, :], tol)
        for (n, vg), (ns, vs) in zip(dp_leaves, single_leaves):
            vs_single = max(vs_single, close_err(vg, vs[lo:hi] if n in ROW_LOCAL else vs, tol))
        if c["tf32"]:
            vs_oracle = close_err(tr[:1], to[:1, lo:hi, :], tol)
        else:
            vs_oracle = close_err(tr, to[:, lo:hi, :], tol)
            for (n, vg), (no, vo) in zip(dp_leaves, fo.leaves()):
                vs_oracle = max(vs_oracle, close_err(vg, np.asarray(vo)[lo:hi] if n in ROW_LOCAL else np.asarray(vo), tol))
        t = torch.tensor([vs_single, vs_oracle, first], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out["cases"].append({"case": f"{c['name']}: LSTM B={c['batch']} H={c['hidden']} I={c['inp']} T={c['T']}, {c['iters']} iterations, lr={lr:g}",
                             "tol": tol, "worst_vs_single_gpu": float(t[0]), "worst_vs_oracle": float(t[1]),
                             "worst_first_iteration_vs_single_gpu": float(t[2]),
                             "fused_exchanges_per_iter": st["dp_fused_updates"], "in_switch": st["dp_multimem_updates"],
                             "nccl_allreduces_per_iter": st["allreduces"], "exchange": exchange})
    # The exchange is judged on the first iteration (identical starting weights: only the exchange differs) and against
    # the oracle; the multi-iteration trajectory of the TF32 case is allowed 5x — three iterations of the reference's
    # update rule amplify summation-order noise by two orders of magnitude even at the scaled learning rate.
    out["multi_iteration_allowance"] = {"f32": 1.0, "tf32": 5.0}
    out["ok"] = all(c["worst_vs_single_gpu"] <= (5.0 if c["case"].startswith("tf32") else 1.0) and c["worst_vs_oracle"] <= 1.0
                    and c["worst_first_iteration_vs_single_gpu"] <= 1.0 for c in out["cases"])
    out["max_rel_err"] = max(max(c["worst_vs_single_gpu"], c["worst_vs_oracle"]) * c["tol"] for c in out["cases"])
    out["tol"] = {c["case"].split(":")[0]: c["tol"] for c in out["cases"]}
    out["exchange"] = out["cases"][0]["exchange"]
    return out


# ---------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref = the reference's src/cpu compiled unmodified, under the restated walker)
# ---------------------------------------------------------------------------------------------------------
def ref_oracle(build):
    from oracle.oracle import get_oracle
    orc = get_oracle()
    kind = "reference" if orc.use_ref(build) else "port"
    if kind == "port":
        orc.use_port()
    return orc, kind


BUILD_TAG = {"O0": "reference src/cpu, build.rs flags (-O0)", "O3": "reference src/cpu, -O3 -ffp-contract=off"}


def cpu_time(f, lr, iters, warm):
    if warm:
        f.minimize({}, lr, warm)
    t0 = time.perf_counter()
    f.minimize({}, lr, iters)
    return (time.perf_counter() - t0) / iters


def reference_sample_step(wl: Workload, d: int, steps: int, warmup: int, build="O3"):
    """`steps` passes of the reference CPU path over a BOUNDED SAMPLE of the workload: the same LSTM graph (same T, dag
    mode) at B = H = I = d — the reference's gemm only accepts equal square operands (src/cpu/ops.cpp:92-99, Q7), so
    the bench shape itself cannot run on it. Returns seconds per sample step and the flop fraction the sample is."""
    orc, kind = ref_oracle(build)
    orc.set_mode("dag")
    f, _ = build_lstm(orc, d, d, d, wl.T, wl.seed, with_inputs=False)
    s = cpu_time(f, LR, steps, warmup)
    orc.set_mode("exact")
    orc.use_port()
    sample_flop = (46 * (wl.T - 1) + 30) * d ** 3
    return s, sample_flop / wl.flop_global, kind, sample_flop


def cpu_baseline_block(wl: Workload, d: int, steps: int, warmup: int):
    s, frac, kind, sample_flop = reference_sample_step(wl, d, steps, warmup)
    value = (1.0 / s) * frac
    return {"value": value, "unit": "iterations/s", "cores": 1, "kind": kind,
            "sample": f"{steps} passes (+{warmup} warm-up) of the same LSTM graph (T={wl.T}, dag mode) at B=H=I={d}: "
                      f"{sample_flop / 1e9:.2f} GFLOP per pass, {s * 1e3:.0f} ms per pass = {sample_flop / s / 1e9:.2f} GFLOP/s on one "
                      f"host core ({BUILD_TAG['O3']}; bit-identical to the -O0 build). The reference gemm rejects non-square "
                      f"operands, so the bench shape cannot run on it: value = passes/s x flop fraction {frac:.3e} (EXTRAPOLATED "
                      "by flop count; the measured, un-extrapolated GPU-vs-reference pairs are under `same_config`)",
            "sample_seconds_per_pass": s, "sample_flop_fraction": frac, "host_cpus": os.cpu_count()}, s


def same_config_reference(budget_s_per_point=1.5):
    """Workloads the reference CAN run, measured on its CPU path (one core), both builds: the README LSTM sweep in exact
    mode (BASELINE config 3) and a square LSTM in dag mode."""
    from helpers import readme_lstm
    out = {"c3": {}, "square": {}}
    for build in ("O0", "O3"):
        orc, kind = ref_oracle(build)
        if kind != "reference":
            out["c3"][build] = "oracle/_ref is not built"
            continue
        orc.set_mode("exact")
        rows = {}
        for d in (5, 10, 15, 20, 25, 30):
            f = readme_lstm(orc, d)
            one = cpu_time(f, LR, 2, 1)
            iters = int(max(3, min(1000, budget_s_per_point / one)))
            rows[str(d)] = {"us_per_iter": cpu_time(f, LR, iters, 0) * 1e6, "iterations_timed": iters}
        out["c3"][build] = rows
    orc, kind = ref_oracle("O3")
    orc.set_mode("dag")
    d, T = 256, 8
    f, _ = build_lstm(orc, d, d, d, T, 4, with_inputs=False)
    s = cpu_time(f, LR, 2, 0)
    out["square"] = {"workload": f"LSTM B=H=I={d}, T={T}, dag mode, f32", "build": BUILD_TAG["O3"] if kind == "reference" else "oracle port",
                     "seconds_per_iter": s, "iters_per_s": 1.0 / s, "iterations_timed": 2, "cores": 1}
    orc.set_mode("exact")
    orc.use_port()
    return out


def same_config_block(api):
    """GPU and reference CPU measured on the SAME workloads in this run (no extrapolation)."""
    from helpers import readme_lstm
    ref = same_config_reference()
    out = {"note": "both arms measured in this run on workloads the reference itself can execute; reference = src/cpu compiled "
                   "unmodified (oracle/_ref) under the restated walker, one host core",
           "c3_readme_lstm_T2_exact": [], "square_lstm_dag": None}
    api.set_mode("exact")
    api.set_tf32(False)
    for d in (5, 10, 15, 20, 25, 30):
        f = readme_lstm(api, d)
        f.minimize({}, LR, 5)
        ms = api.time_iterations(f, LR, 20, 1000)
        row = {"d": d, "params_13d2": 13 * d * d, "gpu_us_per_iter": ms * 1e3, "gpu_iterations_timed": 1000}
        for build in ("O0", "O3"):
            r = ref["c3"].get(build)
            if isinstance(r, dict):
                row[f"ref_{build}_us_per_iter"] = r[str(d)]["us_per_iter"]
                row[f"ref_{build}_iterations_timed"] = r[str(d)]["iterations_timed"]
                row[f"gpu_over_ref_{build}"] = r[str(d)]["us_per_iter"] / (ms * 1e3)
        out["c3_readme_lstm_T2_exact"].append(row)
        del f
    api.set_mode("dag")
    sq = ref["square"]
    d, T = 256, 8
    res = {"workload": sq["workload"], "ref_build": sq["build"], "ref_iters_per_s": sq["iters_per_s"], "ref_cores": 1}
    for product in ("f32", "tf32"):
        api.set_tf32(product == "tf32")
        f, _ = build_lstm(api, d, d, d, T, 4, with_inputs=False)
        f.minimize({}, LR, 3)
        ms = api.time_iterations(f, LR, 5, 50)
        res[f"gpu_{product}_iters_per_s"] = 1e3 / ms
        res[f"gpu_{product}_over_ref"] = (1e3 / ms) / sq["iters_per_s"]
        del f
    out["square_lstm_dag"] = res
    api.set_tf32(False)
    api.set_mode("exact")
    return out


def secondary_small_configs(api):
    """BASELINE configs 1-3 on the GPU (parity-test cases, timed for the roofline table of DESIGN.md)."""
    from helpers import c2_graph, readme_lstm
    out = {}
    pk = peaks()
    api.set_mode("exact")
    api.set_tf32(False)
    F = api.Function
    x = F.scalar_param("x", 1.0)
    f = (x + F.scalar(3.0)).sq()
    f.minimize({}, LR, 10)
    ms = api.time_iterations(f, LR, 20, 1000)
    out["c1_scalar"] = {"config": "C1 (x+3)^2, x=1, lr=.01", "us_per_iter": ms * 1e3, "iters_per_s": 1e3 / ms,
                        "executor": "one resident kernel runs all iterations"}
    n, reps = 8192, 100
    f = c2_graph(api, n)
    f.minimize({}, LR, 2)
    ms = api.time_iterations(f, LR, 3, reps)
    alg = 24.0 * n * n
    out["c2_elementwise"] = {"config": f"C2 tanh(sigmoid(W)-sq(V))+abs(W), {n}x{n} f32, {reps} SGD steps", "ms_per_step": ms,
                             "steps_per_s": 1e3 / ms, "algorithmic_bytes_per_step": alg,
                             "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                          "frac": alg / (ms * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk["source"]}}
    del f
    gc.collect()
    rows = []
    api.set_mode("dag")
    for d in (5, 10, 15, 20, 25, 30):
        f = readme_lstm(api, d)
        f.minimize({}, LR, 5)
        ms = api.time_iterations(f, LR, 20, 1000)
        rows.append({"d": d, "us_per_iter": ms * 1e3})
        del f
    out["c3_readme_lstm_T2_dag_mode"] = rows   # (exact mode — the reference's semantics — is in `same_config`)
    api.set_mode("exact")
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.config, max(args.gpus, 1))
    steps, warm = max(1, args.steps), max(0, args.warmup)
    # sample size: one pass should take about 2 s at the ~1.2 GFLOP/s of the reference's naive gemm (-O3)
    d = 128 if wl.T >= 32 else 192
    t0 = time.perf_counter()
    base, s = cpu_baseline_block(wl, d, steps, warm)
    cfg = wl.config()
    cfg["product"] = args.product
    cfg["parallelism"] = f"dp{wl.world}" if wl.world > 1 else "single"
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "iterations/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm,
        # what one step of THIS arm took on the clock: a pass over the bounded sample (the contract's "each step a bounded
        # sample of that workload"); `value` scales it to the whole workload by flop count and says so
        "ms_per_step": s * 1e3, "step_definition": f"one pass of the reference CPU path over the B=H=I={d} sample of the workload",
        "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_same_config:
        line["same_config"] = same_config_reference()
    line["wall_s"] = time.perf_counter() - t0
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--product", default=os.environ.get("CG_BENCH_PRODUCT", "tf32"), choices=["tf32", "tf32x3", "f32"],
                    help="tf32: tcgen05 TF32 products (1e-3); tf32x3: FP32-accurate products on the tensor core, 3 x TF32 split (1e-5); "
                         "f32: FP32 SIMT products (1e-5)")
    ap.add_argument("--config", default=os.environ.get("CG_BENCH_CONFIG", "c5"), choices=["c4", "c5"],
                    help="c5 (default, every N): BASELINE config 5, T=32, global batch 8192, strong scaling; "
                         "c4: BASELINE config 4, 1024 rows per GPU, weak scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="N = 1: skip the secondary configs and the same-config block")
    ap.add_argument("--no-same-config", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the data-parallel parity check before timing")
    ap.add_argument("--parity-only", action="store_true", help="N > 1: run the parity check, print it, time nothing")
    ap.add_argument("--parity-tf32-batch", type=int, default=0, help="(experiments) global batch of the TF32 parity case")
    ap.add_argument("--parity-fixed-lr", action="store_true", help="(experiments) do not shrink the parity learning rate with the batch")
    ap.add_argument("--tf32-pair", type=int, default=1, choices=[0, 1], help="(experiments) 0: single-CTA TF32 product kernel")
    ap.add_argument("--multimem", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: in-switch (NVLS multimem) weight exchange — auto: from three ranks up")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import computational_graph_b200 as pkg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    api = pkg.get_api()
    api.set_device(local_rank)
    api.set_tf32_pair(bool(args.tf32_pair))
    wl = Workload(args.config, world)
    dist = None
    parity = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        if not args.no_parity:
            parity = dp_parity(api, pkg, torch, dist, rank, world, local_rank, args.multimem,
                               tf32_batch=args.parity_tf32_batch or None, scale_lr=not args.parity_fixed_lr)
            if args.parity_only:
                if rank == 0:
                    print(json.dumps({"n_gpus": world, "parity": parity}))
                dist.destroy_process_group()
                sys.exit(0 if parity["ok"] else 1)
            if not parity["ok"]:
                if rank == 0:
                    print(json.dumps({"metric": METRIC, "error": "data-parallel parity check failed; nothing was timed",
                                      "n_gpus": world, "parity": parity}))
                dist.destroy_process_group()
                sys.exit(1)
        pkg.dp.init_from_torch_distributed(api, local_rank)
        if args.multimem != "auto":
            api.dp_set_multimem(2 if args.multimem == "on" and api.dp_multimem_supported() else 0)
    measured = {}
    if rank == 0 and world == 1:
        measured = measure_peaks(torch, api, sorted({wl.rows, 1024}))
    nccl_ref = nccl_allreduce_reference(torch, dist, world) if world > 1 else None

    line = measure_lstm(api, torch, dist, wl, args.product, args.steps, args.warmup, rank, local_rank, measured)
    if world > 1:
        line["exchange"] = api.dp_exchange_name()
        line["multimem_supported"] = api.dp_multimem_supported() or api.dp_multimem_reason()
        if parity is not None:
            line["parity"] = parity
        n_ex = line["kernel_share"]["kernels"]["exchange"]
        line["exchange_vs_nccl"] = {
            "ours_us_per_weight_serialised": line["kernel_share"]["exchange_ms_serialised"] * 1e3 / n_ex if n_ex else None,
            "ours_includes": "sum over ranks + SGD update + new weight in every rank's memory (one kernel)",
            "nccl": nccl_ref}
    if measured:
        line["peaks"] = measured
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_baseline_block(wl, 128 if wl.T >= 32 else 192, 3, 1)
    if rank == 0 and world == 1 and not args.no_secondary:
        sec = {}
        other = Workload("c4" if args.config == "c5" else "c5", 1)
        k = max(3, min(args.steps, 10))
        for product in ("tf32", "tf32x3", "f32"):
            l2 = measure_lstm(api, torch, None, other, product, k if product != "f32" else max(3, k // 2), 3, 0, local_rank, measured)
            sec[f"{other.name}_{product}"] = {key: l2[key] for key in ("value", "unit", "ms_per_step", "steps", "warmup", "scaling", "dtype",
                                                                   "config", "e2e", "gpu_launches", "roofline", "kernel_share", "clocks")}
        sec.update(secondary_small_configs(api))
        line["secondary"] = sec
        if not args.no_same_config:
            line["same_config"] = same_config_block(api)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        api.dp_shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
