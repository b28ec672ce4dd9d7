# round-1 final evidence, 1 GPU: the whole GPU suite, smoke, the bench lines (ours TF32 / FP32, reference arm), the
# CPU reference timings of C1-C3 on this box's host cores, C1-C3 on the GPU, the ncu launch list of the bench command
# and full captures of the compiled resident kernel (C3 d=30, exact) and of the quiet C2 kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,power.limit --format=csv > gpurun_out/final_gpu.txt 2>&1; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/final_gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_tf32.json 2> gpurun_out/final_bench_tf32.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --product f32 --no-cpu-baseline > gpurun_out/final_bench_f32.json 2> gpurun_out/final_bench_f32.err; echo "bench f32 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?"
timeout 600 python bench.py --impl reference --cpu-configs > gpurun_out/final_cpu_configs.jsonl 2> gpurun_out/final_cpu_configs.err; echo "cpu configs rc=$?"
timeout 300 python tools/bench_configs.py > gpurun_out/final_configs.jsonl 2> gpurun_out/final_configs.err; echo "configs rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches_tf32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/bench_configs.py --only c3 > gpurun_out/final_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tape_jit -s 17 -c 1 -o gpurun_out/final_prof_tape_c3_d30 python tools/bench_configs.py --only c3 > gpurun_out/final_ncu_tape.log 2>&1
echo "tape full rc=$?"
python tools/bench_configs.py --only c2 --reps 3 > gpurun_out/final_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ew_jit -s 4 -c 2 -o gpurun_out/final_prof_ew_c2 python tools/bench_configs.py --only c2 --reps 3 > gpurun_out/final_ncu_ew.log 2>&1
echo "ew full rc=$?"
cut -c1-300 gpurun_out/final_bench_tf32.json; cut -c1-200 gpurun_out/final_bench_reference.json
