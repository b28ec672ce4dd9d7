# round 2, N GPUs: the driver's command line at head (pair kernel, 32 work queues, in-run parity, exchange vs ncclAllReduce)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29801 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02p_bench_c5_n$N.json 2> gpurun_out/r02p_bench_c5_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r02p_bench_c5_n$N.err | cut -c1-300
python - gpurun_out/r02p_bench_c5_n$N.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],3), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "exchange_ms(serialised)", round(d["kernel_share"]["exchange_ms_serialised"],2), "product_ms", round(d["kernel_share"]["product_ms"],2), "ew_ms", round(d["kernel_share"]["elementwise_ms"],2), "|", d.get("exchange"))
    p=d.get("parity")
    if p: print("  parity ok", p["ok"], [(c["case"][:4], round(c["worst_vs_single_gpu"],3), round(c["worst_vs_oracle"],3), round(c["worst_first_iteration_vs_single_gpu"],3), c["in_switch"]) for c in p["cases"]])
    print("  clocks", d["clocks"])
except Exception as e:
    print("  unreadable:", e)
PY
python - gpurun_out/r02p_bench_c5_n$N.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  exchange_vs_nccl", d.get("exchange_vs_nccl"))
PY
