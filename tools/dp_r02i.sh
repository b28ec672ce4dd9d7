# round 2, 2 GPUs: which of the three changes (pair kernel, two exchange lanes, 32 work queues) costs overlap at N = 2?
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; env "$@" timeout 400 $TR --master-port 29811 bench.py --gpus $N --steps 10 --warmup 3 --no-parity $EXTRA > gpurun_out/r02i_n2_$name.json 2> gpurun_out/r02i_n2_$name.err; echo "$name rc=$?"; }
EXTRA="" run lanes1 CG_DP_LANES=1
EXTRA="" run conn8 CUDA_DEVICE_MAX_CONNECTIONS=8
EXTRA="--tf32-pair 0" run single_cta X=1
EXTRA="" run ctas32 CG_DP_CTAS=32
for f in gpurun_out/r02i_n2_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],3), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "exchange_ms(serialised)", round(d["kernel_share"]["exchange_ms_serialised"],2), "product_ms", round(d["kernel_share"]["product_ms"],2), "clock", d["clocks"]["sm_mhz"])
except Exception as e:
    print("  unreadable:", e)
PY
done
