mkdir -p gpurun_out
python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
python tools/gemm_bench.py > gpurun_out/gemm_tf32.json 2> gpurun_out/gemm_tf32.err; echo "gemm rc=$?"
python tools/gemm_bench.py --product f32 --reps 10 > gpurun_out/gemm_f32.json 2>> gpurun_out/gemm_tf32.err
python - > gpurun_out/h2d.txt 2>&1 <<'PY'
import torch, time
x = torch.empty(16*1024*1024, dtype=torch.float32, pin_memory=True)  # 64 MiB
d = torch.empty_like(x, device="cuda")
for n in (4*1024*1024, 16*1024*1024):
    for _ in range(3): d[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): d[:n].copy_(x[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    print("H2D pinned", n*4/2**20, "MiB", n*4/dt/1e9, "GB/s")
    for _ in range(3): x[:n].copy_(d[:n], non_blocking=True)
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(10): x[:n].copy_(d[:n], non_blocking=True)
    torch.cuda.synchronize(); dt=(time.perf_counter()-t)/10
    print("D2H pinned", n*4/2**20, "MiB", n*4/dt/1e9, "GB/s")
PY
cat gpurun_out/configs.jsonl gpurun_out/gemm_tf32.json gpurun_out/gemm_f32.json gpurun_out/h2d.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tf32.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/gemm_bench.py --reps 1 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 -c 12 -o gpurun_out/prof_gemm_tf32 python tools/gemm_bench.py --reps 1 > gpurun_out/ncu_gemm.log 2>&1
echo "gemm full rc=$?"
python tools/bench_configs.py --only c2 --reps 2 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ew_program -s 3 -c 2 -o gpurun_out/prof_ew_c2 python tools/bench_configs.py --only c2 --reps 2 > gpurun_out/ncu_ew.log 2>&1
echo "ew full rc=$?"
ls -la gpurun_out
