# round 2, 1 GPU: ncu of the FP32 SIMT product (cp.async ring + FFMA2), NN and TN at 8192 rows
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tiled_kernel -s 3 -c 6 -o gpurun_out/r02h_gemm_f32 python tools/gemm_bench.py --reps 1 --product f32 --b 8192 > gpurun_out/r02h_ncu_f32.log 2>&1; echo "ncu f32 rc=$?"
ls -la gpurun_out/r02h_gemm_f32.ncu-rep
