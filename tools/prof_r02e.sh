# round 2, 1 GPU: (1) graph tests after the transitive reduction, (2) the end-to-end probe at T = 32, (3) FP32 SIMT product with FFMA2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_fullsize.py tests/test_gpu_matrix_abi.py -m gpu -x -q > gpurun_out/r02e_pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02e_pytest_graph.log
timeout 600 python tools/e2e_probe2.py 32 8192 > gpurun_out/r02e_e2e_probe_T32.json 2> gpurun_out/r02e_e2e_probe_T32.err; echo "e2e probe rc=$?"; cat gpurun_out/r02e_e2e_probe_T32.json
timeout 300 python tools/gemm_bench.py --product f32 --reps 30 > gpurun_out/r02e_gemm_f32.json 2>&1; cat gpurun_out/r02e_gemm_f32.json
timeout 300 python tools/gemm_bench.py --product f32 --reps 10 --b 8192 >> gpurun_out/r02e_gemm_f32.json 2>&1; tail -1 gpurun_out/r02e_gemm_f32.json
timeout 600 python bench.py --config c4 --product f32 --no-secondary --no-cpu-baseline --steps 5 > gpurun_out/r02e_bench_c4_f32.json 2> gpurun_out/r02e_bench_c4_f32.err; echo "f32 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02e_bench_c4_f32.json").read().strip().splitlines()[-1])
print("c4 f32: value", d["value"], "roofline", d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"]["frac"], d["clocks"], {k:v for k,v in d["peaks"].items() if "fp32" in k and "how" not in k})
PY
