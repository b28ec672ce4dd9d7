# round 2, 2 GPUs: is the N = 8 parity failure the exchange or the dynamics? Same TF32 case (global batch 2048) at 2 ranks,
# with both exchanges, with the learning rate fixed (.01) and scaled with the batch.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for mm in off on; do
  for lr in fixed scaled; do
    extra=""; [ "$lr" = "fixed" ] && extra="--parity-fixed-lr"
    timeout 300 $TR --master-port 29781 bench.py --gpus 2 --parity-only --parity-tf32-batch 2048 --multimem $mm $extra > gpurun_out/r02f_parity_b2048_mm${mm}_lr$lr.json 2> gpurun_out/r02f_parity_b2048_mm${mm}_lr$lr.err; echo "mm=$mm lr=$lr rc=$?"
    python - gpurun_out/r02f_parity_b2048_mm${mm}_lr$lr.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    for c in d["parity"]["cases"]: print("   ", c["case"], "| single", round(c["worst_vs_single_gpu"],3), "oracle", round(c["worst_vs_oracle"],3), "first", round(c["worst_first_iteration_vs_single_gpu"],3), "in_switch", c["in_switch"])
except Exception as e: print("   unreadable", e)
PY
  done
done
