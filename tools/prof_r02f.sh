# round 2, 1 GPU: tied-parameter tests; does the number of hardware work queues decide how well the H2D copies of an end-to-end call overlap the iteration?
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -x -q -k "tied or corrected" > gpurun_out/r02f_pytest_tied.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02f_pytest_tied.log
for c in 32 2; do
  CUDA_DEVICE_MAX_CONNECTIONS=$c timeout 600 python tools/e2e_probe2.py 32 8192 > gpurun_out/r02f_e2e_probe_T32_conn$c.json 2> gpurun_out/r02f_e2e_probe_conn$c.err; echo "conn=$c rc=$?"; cat gpurun_out/r02f_e2e_probe_T32_conn$c.json
done
