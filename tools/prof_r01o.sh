mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -x -q -m gpu 2>&1 | tail -3
python tools/e2e_probe.py 2>&1 | sed -n 1,4p
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_o_exact1.json 2> gpurun_out/bench_o_exact1.err; echo "bench exact rc=$?"
CG_GRAPH_EXACT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_o_exact0.json 2> gpurun_out/bench_o_exact0.err; echo "bench streams rc=$?"
for f in gpurun_out/bench_o_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"e2e": {[^}]*}' $f; done
