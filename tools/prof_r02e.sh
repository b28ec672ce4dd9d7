# round 2, 1 GPU: after the transitive reduction of the iteration graph's edges — graph tests, then the end-to-end probe at T = 32
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02e_pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02e_pytest_graph.log
timeout 600 python tools/e2e_probe2.py 32 8192 > gpurun_out/r02e_e2e_probe_T32.json 2> gpurun_out/r02e_e2e_probe_T32.err; echo "e2e probe rc=$?"; cat gpurun_out/r02e_e2e_probe_T32.json
