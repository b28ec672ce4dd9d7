# round 2, 1 GPU: pair kernel as the default — GPU test tier, the bench line, launch list + ncu captures of the dominant product and of the FP32 SIMT product
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02d_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02d_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"], "whole", d["roofline"]["whole_step"]["frac"], d["clocks"])
for k in ("c4_tf32","c4_f32"):
    s=d["secondary"][k]; print(k, s["value"], s["e2e"]["value"], s["roofline"]["frac"], s["roofline"]["achieved"], s["clocks"])
PY
timeout 600 python tools/e2e_probe2.py 8 8192 > gpurun_out/r02d_e2e_probe.json 2> gpurun_out/r02d_e2e_probe.err; echo "e2e probe rc=$?"; cat gpurun_out/r02d_e2e_probe.json
# launch list of the config-4 line (short) and full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02d_launches_c4_tf32.csv python bench.py --config c4 --no-secondary --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r02d_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair_kernel -s 6 -c 3 -o gpurun_out/r02d_gemm_pair python tools/gemm_bench.py --reps 3 --b 8192 > gpurun_out/r02d_ncu_pair.log 2>&1; echo "ncu pair rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tiled_kernel -s 3 -c 2 -o gpurun_out/r02d_gemm_f32 python tools/gemm_bench.py --reps 2 --product f32 > gpurun_out/r02d_ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
ls -la gpurun_out/*.ncu-rep
