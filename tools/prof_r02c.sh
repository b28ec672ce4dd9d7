# round 2, 1 GPU: tcgen05 TF32 product vs cuBLAS TF32, both sustained (single-CTA and CTA-pair kernels); FP32 SIMT product after the occupancy fix
mkdir -p gpurun_out
timeout 600 python tools/gemm_vs_cublas.py --pair 0 > gpurun_out/r02c_gemm_vs_cublas_pair0.jsonl 2> gpurun_out/r02c_gemm_pair0.err; echo "pair0 rc=$?"
timeout 600 python tools/gemm_vs_cublas.py --pair 1 > gpurun_out/r02c_gemm_vs_cublas_pair1.jsonl 2> gpurun_out/r02c_gemm_pair1.err; echo "pair1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02c_gemm_vs_cublas_pair0.jsonl","gpurun_out/r02c_gemm_vs_cublas_pair1.jsonl"):
    for l in open(f):
        d=json.loads(l)
        print(d["rows"], d["case"], "pair" if d["pair_kernel"] else "single", "ours %.0f TF @%s MHz %sW | cublas %.0f TF @%s MHz %sW | ratio %.3f" % (d["ours_tflops"], d["ours_sm_mhz"], d["ours_watts"], d["cublas_tflops"], d["cublas_sm_mhz"], d["cublas_watts"], d["ours_over_cublas"]))
PY
timeout 600 python bench.py --config c4 --product f32 --no-secondary --no-cpu-baseline --steps 5 > gpurun_out/r02c_bench_c4_f32.json 2> gpurun_out/r02c_bench_c4_f32.err; echo "f32 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02c_bench_c4_f32.json").read().strip().splitlines()[-1])
print("c4 f32: value", d["value"], "roofline", d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"]["frac"], d["clocks"])
PY
timeout 300 python tools/h2d_probe.py 8192 33 > gpurun_out/r02c_h2d_probe.json 2> gpurun_out/r02c_h2d_probe.err; echo "h2d rc=$?"; cat gpurun_out/r02c_h2d_probe.json
