#!/usr/bin/env python
"""bench.py — LSTM fwd + backprop + SGD iterations/sec (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--product tf32|f32]

One "step" = one pass of the `minimize` loop body (src/optimize.rs:64-71) over the BASELINE config-4 LSTM
(H = I = 4096, T = 8, batch 1024 rows per GPU, f32, dag mode). N > 1 shards the batch rows over the ranks
(weak scaling: 1024 rows per GPU); every weight gradient is summed over the ranks before its update by the fused
exchange kernel over NVLink peer memory (ncclAllReduce when peer access is unavailable or CG_DP_FUSED=0).
`--config c5` selects BASELINE config 5 instead (T = 32, global batch 8192 sharded over the ranks, strong scaling).

  value    : iterations/sec with every input resident in HBM, CUDA events on the library stream, max over ranks
  e2e      : the same metric through the public call `f.minimize(&args, lr, 1)` (C ABI cg_minimize) with HOST
             buffers: H2D of the T inputs + target from pinned memory and D2H of the root value inside the
             timed region
  roofline : the dominant kernel (the [1024 x 4096] . [4096 x 4096] product), timed live with CUDA events
  cpu_baseline / --impl reference : the reference's own CPU arithmetic (oracle/_ref, compiled unmodified from
             the reference's src/cpu) under the restated walker, on a bounded square sample, 1 host core
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

H = I = 4096
T = 8
B_PER_GPU = 1024
CONFIG = "c4"   # "c4": BASELINE config 4, weak scaling (1024 rows per GPU); "c5": BASELINE config 5 (T = 32, global batch
                # 8192 sharded over the ranks), strong scaling — select with --config
GLOBAL_BATCH_C5 = 8192
LR = 0.01
SEED = 4
# (46 (T-1) + 30) B H^2 flop per iteration in dag mode with the dead dX of the constant inputs skipped
FLOP_PER_ITER_PER_GPU = (46 * (T - 1) + 30) * B_PER_GPU * H * H
SAMPLE_D = 256  # CPU sample: square LSTM (the reference gemm only accepts equal square operands, Q7)


def configure(name: str, world: int) -> None:
    """Sets the module-level workload constants for the chosen BASELINE config."""
    global CONFIG, T, B_PER_GPU, FLOP_PER_ITER_PER_GPU
    CONFIG = name
    if name == "c5":
        if GLOBAL_BATCH_C5 % world:
            raise SystemExit(f"config 5: the global batch of {GLOBAL_BATCH_C5} rows does not divide over {world} ranks")
        T, B_PER_GPU = 32, GLOBAL_BATCH_C5 // world
    else:
        T, B_PER_GPU = 8, 1024
    FLOP_PER_ITER_PER_GPU = (46 * (T - 1) + 30) * B_PER_GPU * H * H


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "sm_max_mhz": 1965.0, "source": "fallback"}


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


def cpu_sample(steps, warmup):
    """The reference's CPU arithmetic (oracle/_ref O3 build: same results bit for bit as the -O0 build.rs flags) under
    the oracle walker, dag mode, square d = SAMPLE_D sample, one host core. Returns (it/s of the sample, it/s scaled
    to the bench workload by flop count, description)."""
    from oracle.oracle import get_oracle
    orc = get_oracle()
    kind = "reference" if orc.use_ref("O3") else "port"
    if kind == "port":
        orc.use_port()
    orc.set_mode("dag")
    d = SAMPLE_D
    f, _ = build_lstm(orc, d, d, d, T, SEED, with_inputs=False)
    if warmup:
        f.minimize({}, LR, warmup)
    t0 = time.perf_counter()
    f.minimize({}, LR, steps)
    dt = time.perf_counter() - t0
    orc.set_mode("exact")
    sample_flop = (46 * (T - 1) + 30) * d * d * d
    its = steps / dt
    scaled = its * sample_flop / FLOP_PER_ITER_PER_GPU
    desc = (f"{steps} iterations of the same LSTM graph at B=H=I={d}, T={T}, dag mode ({sample_flop / 1e9:.2f} GFLOP/iter; "
            f"the reference gemm only accepts square operands); value = sample it/s x flop ratio {sample_flop / FLOP_PER_ITER_PER_GPU:.3e} "
            f"(extrapolated to B={B_PER_GPU}, H=I={H}); sample ran at {its:.3f} it/s")
    return its, scaled, kind, desc


def cpu_config_lines():
    """SURVEY §8d "CPU reference timing": the reference's CPU arithmetic (oracle/_ref, compiled unmodified from
    src/cpu with build.rs's flags = -O0, and with -O3 -ffp-contract=off as the fair column; bit-identical results)
    under the restated walker, one host core, on the secondary BASELINE configs C1-C3. One JSON line per config."""
    from helpers import c2_graph, readme_lstm
    from oracle.oracle import get_oracle
    orc = get_oracle()
    orc.set_mode("exact")
    ncpu = os.cpu_count()

    def timed(f, lr, iters, warm=1):
        if warm:
            f.minimize({}, lr, warm)
        t0 = time.perf_counter()
        f.minimize({}, lr, iters)
        return (time.perf_counter() - t0) / iters

    for build in ("O0", "O3"):
        if not orc.use_ref(build):
            print(json.dumps({"impl": "reference", "build": build, "unavailable": "oracle/_ref is not built"}))
            continue
        tag = {"O0": "reference src/cpu, build.rs flags (-O0)", "O3": "reference src/cpu, -O3 -ffp-contract=off"}[build]
        x = orc.Function.scalar_param("x", 1.0)
        s = timed((x + orc.Function.scalar(3.0)).sq(), LR, 1000, warm=10)
        print(json.dumps({"impl": "reference", "build": tag, "cores": 1, "host_cpus": ncpu, "config": "C1 (x+3)^2 scalar",
                          "us_per_iter": s * 1e6, "iters_per_s": 1.0 / s}))
        n = 8192 if build == "O3" else 4096
        s = timed(c2_graph(orc, n), LR, 2, warm=0) * (8192 / n) ** 2
        print(json.dumps({"impl": "reference", "build": tag, "cores": 1, "host_cpus": ncpu,
                          "config": "C2 tanh(sigmoid(W)-sq(V))+abs(W), 8192x8192",
                          "sample": f"2 steps at {n}x{n}" + ("" if n == 8192 else ", scaled by element count to 8192x8192"),
                          "ms_per_step": s * 1e3, "steps_per_s": 1.0 / s}))
        for d in (5, 10, 15, 20, 25, 30):
            iters = 200 if d <= 15 else 50
            s = timed(readme_lstm(orc, d), LR, iters, warm=2)
            print(json.dumps({"impl": "reference", "build": tag, "cores": 1, "host_cpus": ncpu,
                              "config": f"C3 README LSTM T=2 d={d} ({13 * d * d} params), exact mode",
                              "sample": f"{iters} iterations", "us_per_iter": s * 1e6, "iters_per_s": 1.0 / s}))
    orc.use_port()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.cpu_configs:
        return cpu_config_lines()
    steps = max(1, min(args.steps, 20))
    warm = min(args.warmup, 1)
    its, scaled, kind, desc = cpu_sample(steps, warm)
    line = {
        "impl": "reference", "metric": "LSTM fwd+backprop+SGD iterations/sec", "value": scaled, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1000.0 / scaled, "higher_is_better": True,
        "scaling": "strong" if CONFIG == "c5" else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, "f32"),
        "cpu_baseline": {"value": scaled, "unit": "iterations/s", "cores": 1, "kind": kind, "sample": desc},
        "e2e": {"value": scaled, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus, product):
    num = "5" if CONFIG == "c5" else "4"
    aggregate = ("value = global iterations/s: the 8192-row batch is sharded over the ranks, one step is one iteration "
                 "of the whole batch (strong scaling)" if CONFIG == "c5" else
                 "value = n_gpus x (global iterations/s): each GPU runs the 1024-row LSTM iteration of its batch shard")
    return {"workload": f"BASELINE config {num}: LSTM H=I={H}, T={T}, batch {B_PER_GPU} rows per GPU (global batch "
                        f"{B_PER_GPU * n_gpus}), f32 params, (lstm(inputs)-target).sq(), lr={LR}, dag mode",
            "product": product, "params_13H2": 13 * H * H, "global_batch": B_PER_GPU * n_gpus, "seq_len": T,
            "parallelism": f"dp{n_gpus}" if n_gpus > 1 else "single",
            "aggregate": aggregate,
            "l2": "per-iteration working set (>8 GB) exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--product", default=os.environ.get("CG_BENCH_PRODUCT", "tf32"), choices=["tf32", "f32"])
    ap.add_argument("--config", default=os.environ.get("CG_BENCH_CONFIG", "c4"), choices=["c4", "c5"],
                    help="c4 (default): BASELINE config 4, weak scaling; c5: BASELINE config 5, T=32, global batch 8192, strong scaling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multimem", default="auto", choices=["auto", "on", "off"],
                    help="N > 1: in-switch (NVLS multimem) weight exchange — auto: from three ranks up")
    ap.add_argument("--cpu-configs", action="store_true",
                    help="with --impl reference: time the reference CPU path on the secondary configs C1-C3 instead")
    args = ap.parse_args()
    configure(args.config, int(os.environ.get("WORLD_SIZE", "1")) if args.impl == "ours" else max(args.gpus, 1))
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        pkg.dp.init_from_torch_distributed(api, local_rank)
        if args.multimem != "auto":
            api.dp_set_multimem(2 if args.multimem == "on" and api.dp_multimem_supported() else 0)
    api.set_mode("dag")
    api.set_tf32(args.product == "tf32")

    f, host_args = build_lstm(api, B_PER_GPU, I, H, T, SEED, with_inputs=True, shard=rank)
    # pinned host buffers for the end-to-end leg
    pinned, keep = {}, []
    for k, v in host_args.items():
        t = torch.empty(v.shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = v
        keep.append(t)
        pinned[k] = api.Constant(t.numpy())  # contiguous f32: wraps the pinned memory without copying
    stats = api.plan_stats(f, LR)
    f.minimize(pinned, LR, 1)  # uploads the inputs once; also the first warm-up iteration

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------------
    api.minimize_async(f, LR, args.warmup)
    barrier()
    launches0 = api.launch_count()
    with ClockSampler(local_rank) as clk:
        ms = api.time_iterations(f, LR, 0, args.steps)
        barrier()
        launches = api.launch_count() - launches0
        # ---- end to end through the public call, host buffers (same clock record) -----------------------
        h2d = sum(int(np.asarray(v.value).nbytes) for v in pinned.values())
        d2h = B_PER_GPU * H * 4
        for _ in range(2):
            f.minimize(pinned, LR, 1, trace=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            f.minimize(pinned, LR, 1, trace=True)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1000.0 / args.steps
    clocks = clk.summary()
    # ---- roofline of the dominant kernel ----------------------------------------------------------------
    import ctypes as C
    kind_ms = (C.c_double * 4)()
    kind_cnt = (C.c_ulonglong * 4)()
    gflop, gms, gcnt = C.c_double(), C.c_double(), C.c_ulonglong()
    api.lib.cg_profile_iteration.argtypes = [C.c_void_p, C.c_float, C.c_double * 4, C.c_ulonglong * 4,
                                             C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    api._check(api.lib.cg_profile_iteration(f._h, LR, kind_ms, kind_cnt, C.byref(gflop), C.byref(gms), C.byref(gcnt)))
    api._check(api.lib.cg_profile_iteration(f._h, LR, kind_ms, kind_cnt, C.byref(gflop), C.byref(gms), C.byref(gcnt)))

    if dist is not None:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    pk = peaks()
    if args.product == "tf32":
        peak = pk["bf16_sustained"] / 2.0
        peak_note = (f"1/2 x {pk['source']} sustained cuBLAS bf16 ({pk['bf16_sustained']} TFLOP/s): tcgen05 kind::tf32 "
                     "issues at half the kind::f16 rate; MEASURED_PEAKS.json holds no TF32 figure")
    else:
        peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        peak_note = f"FP32 FMA pipe: 148 SMs x 128 lanes x 2 flop x {pk['sm_max_mhz']} MHz (nominal; not in MEASURED_PEAKS.json)"
    achieved = gflop.value / (gms.value * 1e-3) / 1e12 if gms.value > 0 else 0.0
    agg = 1 if CONFIG == "c5" else world  # config 5: one step = one iteration of the whole (sharded) batch
    total_ms = sum(kind_ms)
    line = {
        # whole-job aggregate: every rank completes one 1024-row LSTM iteration per step (weak scaling), so the job
        # processes `world` shard-iterations per step; at N = 1 this is plain iterations/s
        "metric": "LSTM fwd+backprop+SGD iterations/sec", "value": agg * 1000.0 / ms, "unit": "iterations/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if CONFIG == "c5" else "weak", "vs_baseline": None, "dtype": args.product, "data": "synthetic",
        "config": workload_config(world, args.product),
        "e2e": {"value": agg * 1000.0 / e2e_ms, "unit": "iterations/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor" if args.product == "tf32" else "fma", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel (ncu --set full,
                     # profiles/r01_gemm_tf32_full.raw.txt): 83.9 MB + 5.0 MB; algorithmic A + B + C = 100.7 MB
                     "traffic": 88.86e6 if args.product == "tf32" else None,
                     "kernel": f"largest product, {gcnt.value} launches/iter, {gflop.value / 1e9:.1f} GFLOP each, "
                               f"{gms.value * 1e3:.1f} us avg (CUDA events around each launch, eager pass)",
                     "peak_source": peak_note,
                     # every product the plan actually launches (dead products of the last step's output gate are
                     # eliminated: 5.12 TFLOP) over the whole step; `survey_model` is SURVEY §8d's closed form
                     "whole_step": {"tflops": stats["gemm_flop_per_iter"] / (ms * 1e-3) / 1e12,
                                    "frac": stats["gemm_flop_per_iter"] / (ms * 1e-3) / 1e12 / peak,
                                    "flop_per_iter": stats["gemm_flop_per_iter"],
                                    "survey_model_flop_per_iter": FLOP_PER_ITER_PER_GPU}},
        "kernel_share": {"elementwise_ms": kind_ms[0], "reduce_ms": kind_ms[1], "product_ms": kind_ms[2],
                         "allreduce_ms": kind_ms[3], "instrumented_total_ms": total_ms,
                         "kernels": {"elementwise": int(kind_cnt[0]), "reduce": int(kind_cnt[1]),
                                     "product": int(kind_cnt[2]), "allreduce": int(kind_cnt[3])}},
        "plan": stats,
    }
    if world > 1:
        line["exchange"] = api.dp_exchange_name()
        line["multimem_supported"] = api.dp_multimem_supported() or api.dp_multimem_reason()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        its, scaled, kind, desc = cpu_sample(2, 0)
        line["cpu_baseline"] = {"value": scaled, "unit": "iterations/s", "cores": 1, "kind": kind, "sample": desc}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        api.dp_shutdown()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
