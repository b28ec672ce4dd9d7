# round 2, 1 GPU: ncu launch list (gpu__time_duration) of the driver's bench command (config 5), after the command itself exited 0 without ncu
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02o_bench_plain.json 2> gpurun_out/r02o_bench_plain.err; echo "plain rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02o_launches_c5.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/r02o_ncu.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py launches gpurun_out/r02o_launches_c5.csv > gpurun_out/r02o_launches_c5.summary.txt 2>&1; head -12 gpurun_out/r02o_launches_c5.summary.txt | cut -c1-200
