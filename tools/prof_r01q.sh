mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/configs_q.jsonl 2> gpurun_out/configs_q.err; echo "configs rc=$?"
grep -o '"config": "C3[^}]*"us_per_iter": [0-9.]*' gpurun_out/configs_q.jsonl
