mkdir -p gpurun_out
for cfg in c5 c4; do for skip in 0 1; do
CG_X_SKIP_EW=$skip timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-secondary --no-cpu-baseline --no-same-config > gpurun_out/x_skip_${cfg}_$skip.json 2> gpurun_out/x_skip_${cfg}_$skip.err
python - gpurun_out/x_skip_${cfg}_$skip.json $cfg $skip <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "skip_ew", sys.argv[3], "ms", round(d["ms_per_step"],2), "e2e ms", round(d["e2e"]["ms_per_step"],2), "products", round(d["kernel_share"]["product_ms"],1), "ew", round(d["kernel_share"]["elementwise_ms"],1), d["clocks"]["sm_mhz"])
PY
done; done
