# round 2, 1 GPU: FP32 SIMT product with the cp.async ring; 32 work queues by default; then the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matrix_abi.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02g_pytest.log
timeout 300 python tools/gemm_bench.py --product f32 --reps 30 > gpurun_out/r02g_gemm_f32.json 2>&1; cat gpurun_out/r02g_gemm_f32.json
timeout 300 python tools/gemm_bench.py --product f32 --reps 10 --b 8192 >> gpurun_out/r02g_gemm_f32.json 2>&1; tail -1 gpurun_out/r02g_gemm_f32.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02g_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"], "whole", d["roofline"]["whole_step"]["frac"], d["clocks"])
for k in ("c4_tf32","c4_f32"):
    s=d["secondary"][k]; print(k, s["value"], s["e2e"]["value"], s["roofline"]["frac"], s["roofline"]["achieved"], s["clocks"])
PY
