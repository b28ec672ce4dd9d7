# round-1 (ninth pass), 1 GPU: root-store elision (C2), small-shape threshold of the compiled resident executor
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -x -q -m gpu > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu_i.log
timeout 300 python tools/bench_configs.py > gpurun_out/configs_i.jsonl 2> gpurun_out/configs_i.err; echo "configs rc=$?"
cut -c1-330 gpurun_out/configs_i.jsonl | head -4
grep -o '"config": "C3[^}]*"us_per_iter": [0-9.]*' gpurun_out/configs_i.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
