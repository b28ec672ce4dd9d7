# round 2, N GPUs (default 2): the driver's own command line — config 5, parity check inside the run — with the default
# exchange and with the in-switch exchange forced
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( time timeout 900 $TR --master-port 29751 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02c_bench_c5_n$N.json 2> gpurun_out/r02c_bench_c5_n$N.err ) 2> gpurun_out/r02c_bench_c5_n$N.time; echo "bench default rc=$?"; grep real gpurun_out/r02c_bench_c5_n$N.time
tail -3 gpurun_out/r02c_bench_c5_n$N.err
if [ "$N" = "2" ]; then
timeout 900 $TR --master-port 29752 bench.py --gpus $N --steps 10 --warmup 3 --multimem on > gpurun_out/r02c_bench_c5_n${N}_mm.json 2> gpurun_out/r02c_bench_c5_n${N}_mm.err; echo "bench multimem rc=$?"
fi
for f in gpurun_out/r02c_bench_c5_n$N*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],3), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "exchange_ms(serialised)", round(d["kernel_share"]["exchange_ms_serialised"],2), "n", d["kernel_share"]["kernels"]["exchange"], "product_ms", round(d["kernel_share"]["product_ms"],2), "|", d.get("exchange"))
    print("  parity", json.dumps(d.get("parity"))[:900])
    print("  clocks", d["clocks"])
except Exception as e:
    print("  unreadable:", e)
PY
done
