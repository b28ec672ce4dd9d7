# round 2, 1 GPU: FP32-accurate products on the tensor core (3 x TF32 split in shared memory, cg_set_tf32(2))
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_matrix_abi.py -m gpu -x -q -k "tf32x3" > gpurun_out/r02l_pytest_x3.log 2>&1; echo "pytest x3 rc=$?"; tail -5 gpurun_out/r02l_pytest_x3.log | cut -c1-400
timeout 200 python tools/gemm_bench.py --product tf32x3 --reps 30 > gpurun_out/r02l_gemm_x3.json 2>&1; tail -1 gpurun_out/r02l_gemm_x3.json
timeout 200 python tools/gemm_bench.py --product tf32x3 --reps 10 --b 8192 >> gpurun_out/r02l_gemm_x3.json 2>&1; tail -1 gpurun_out/r02l_gemm_x3.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_pair -s 3 -c 1 -o gpurun_out/r02l_gemm_x3 -f python tools/gemm_bench.py --product tf32x3 --reps 1 --b 8192 > gpurun_out/r02l_ncu.log 2>&1; echo "ncu rc=$?"
