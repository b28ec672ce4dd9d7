# round 2, 1 GPU: the driver's bench command at head (tf32x3 products in `secondary`); the GPU suite ran in the call before (profiles/r02n_pytest_gpu.log)
mkdir -p gpurun_out
t0=$(date +%s); timeout 1200 python bench.py > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r02n_bench_n1.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02n_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "whole", d["roofline"]["whole_step"]["frac"], d["clocks"])
for k in ("c4_tf32","c4_tf32x3","c4_f32"):
    s=d["secondary"][k]; print(k, round(s["value"],2), round(s["e2e"]["value"],2), round(s["roofline"]["frac"],3), round(s["roofline"]["achieved"],1), s["clocks"]["sm_mhz"])
PY
