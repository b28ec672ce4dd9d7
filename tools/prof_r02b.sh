# round 2, 1 GPU: the new parity tests and the f3 tests only
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_graph.py -m gpu -q --durations=8 -k "T8 or full_size or checkpoint or print_freq" > gpurun_out/r02b_pytest_new2.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r02b_pytest_new2.log
