# round 2, 8 GPUs: one exchange lane vs two (pair kernel, 32 work queues); the first run carries the parity check
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CG_DP_LANES=1 timeout 600 $TR --master-port 29821 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02j_bench_c5_n8_lanes1.json 2> gpurun_out/r02j_bench_c5_n8_lanes1.err; echo "lanes1 rc=$?"
CG_DP_LANES=2 timeout 600 $TR --master-port 29822 bench.py --gpus $N --steps 10 --warmup 3 --no-parity > gpurun_out/r02j_bench_c5_n8_lanes2.json 2> gpurun_out/r02j_bench_c5_n8_lanes2.err; echo "lanes2 rc=$?"
for f in gpurun_out/r02j_bench_c5_n8_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],3), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "exchange_ms(serialised)", round(d["kernel_share"]["exchange_ms_serialised"],2), "product_ms", round(d["kernel_share"]["product_ms"],2), "clock", d["clocks"]["sm_mhz"], "|", d.get("exchange","")[:40])
    p=d.get("parity")
    if p: print("  parity ok", p["ok"], [(c["case"][:4], round(c["worst_vs_single_gpu"],3), round(c["worst_vs_oracle"],3), round(c["worst_first_iteration_vs_single_gpu"],3), c["in_switch"]) for c in p["cases"]])
except Exception as e:
    print("  unreadable:", e)
PY
done
