# round 2, 1 GPU: the GPU test tier, then the new bench line (config 5 at N = 1 with the secondary / same-config blocks)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02a_pytest_gpu.log
( time timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err ) 2> gpurun_out/r02a_bench_n1.time; echo "bench rc=$?"; cat gpurun_out/r02a_bench_n1.time; tail -5 gpurun_out/r02a_bench_n1.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err ) 2> gpurun_out/r02a_bench_ref.time; echo "ref rc=$?"; cat gpurun_out/r02a_bench_ref.time
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02a_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"])
print("peaks", json.dumps(d.get("peaks"))[:800])
for k,v in d.get("secondary",{}).items():
    print(k, json.dumps(v)[:300])
print(json.dumps(d.get("same_config"))[:2500])
PY
