# round-1 (eighth pass), 1 GPU: the compiled resident executor — parity tests, then the C1-C3 timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -x -q -m gpu > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_h.log
timeout 300 python tools/bench_configs.py > gpurun_out/configs_h_jit.jsonl 2> gpurun_out/configs_h_jit.err; echo "configs rc=$?"
cut -c1-200 gpurun_out/configs_h_jit.jsonl
tail -3 gpurun_out/configs_h_jit.err
