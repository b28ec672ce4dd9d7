# round-1 (second pass) evidence: bench line, launch list of the same command, full capture of the specialised
# elementwise kernel on C2 and of the product kernel
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python bench.py --steps 5 --warmup 3 --product f32 --no-cpu-baseline > gpurun_out/bench_final_f32.json 2>> gpurun_out/bench_final.err
python tools/bench_configs.py > gpurun_out/configs_b.jsonl 2> gpurun_out/configs_b.err; echo "configs rc=$?"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32_b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/bench_configs.py --only c2 --reps 2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ew_jit -s 3 -c 2 -o gpurun_out/prof_ew_jit_c2 python tools/bench_configs.py --only c2 --reps 2 > gpurun_out/ncu_ew.log 2>&1
echo "ew full rc=$?"
cat gpurun_out/bench_final.json | cut -c1-400
