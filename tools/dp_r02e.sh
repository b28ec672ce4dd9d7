# round 2, 4 GPUs: what bounds the overlapped exchange — bytes in flight (CTAs x unroll), SMs left to it, or the products beside it
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 400 $TR --master-port 29771 bench.py --gpus $N --steps 10 --warmup 3 --no-parity > gpurun_out/r02e_n${N}_$name.json 2> gpurun_out/r02e_n${N}_$name.err; echo "$name rc=$?"
}
run ctas16 CG_DP_MC_CTAS=16
run ctas64 CG_DP_MC_CTAS=64
run ctas64_u4 CG_DP_MC_CTAS=64 CG_DP_MC_UNROLL=4
run ctas128_u4 CG_DP_MC_CTAS=128 CG_DP_MC_UNROLL=4
run reserve8 CG_DP_RESERVE_SMS=8 CG_DP_MC_CTAS=64 CG_DP_MC_UNROLL=4
for f in gpurun_out/r02e_n${N}_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],3), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],3), "exchange_ms(serialised)", round(d["kernel_share"]["exchange_ms_serialised"],2), "product_ms", round(d["kernel_share"]["product_ms"],2), "clock", d["clocks"]["sm_mhz"])
except Exception as e:
    print("  unreadable:", e)
PY
done
