# round-1 (fourth pass), 1 GPU: the whole GPU suite (incl. full-size C2/C4 properties and hoisting bit-identity), then
# the bench with and without hoisting
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu_d.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_d_hoist1.json 2> gpurun_out/bench_d_hoist1.err; echo "bench hoist rc=$?"
CG_HOIST=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_d_hoist0.json 2> gpurun_out/bench_d_hoist0.err; echo "bench nohoist rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --product f32 --no-cpu-baseline > gpurun_out/bench_d_f32.json 2> gpurun_out/bench_d_f32.err; echo "bench f32 rc=$?"
for f in gpurun_out/bench_d_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"e2e": {[^}]*}' $f; grep -o '"kernel_share": {[^}]*}' $f; done
