# round-1 (seventh pass), 1 GPU: the resident tape executor — parity tests, then the C1-C3 timings with it on and off
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -x -q -m gpu > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_g.log
timeout 300 python tools/bench_configs.py > gpurun_out/configs_g_vm.jsonl 2> gpurun_out/configs_g_vm.err; echo "configs vm rc=$?"
CG_VM_MAX_DIM=0 timeout 300 python tools/bench_configs.py --only c3 > gpurun_out/configs_g_novm.jsonl 2> gpurun_out/configs_g_novm.err; echo "configs novm rc=$?"
cut -c1-220 gpurun_out/configs_g_vm.jsonl
