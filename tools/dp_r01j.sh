# round-1 (tenth pass), 4 GPUs: parity of the fused exchange at world = 4 and the weak-scaling point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29731 tests/dp_check.py --tf32 0 --batch 64 --hidden 48 --inp 40 --steps 2 --iters 3 --exchange fused > gpurun_out/dpj_fused_small_n4.log 2>&1; echo "fused small n4 rc=$?"
grep -E "dp_check world" gpurun_out/dpj_fused_small_n4.log | tail -2
timeout 400 $TR --master-port 29734 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_j_n4_fused.json 2> gpurun_out/bench_j_n4_fused.err; echo "bench fused rc=$?"
for f in gpurun_out/bench_j_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"e2e": {[^}]*}' $f; grep -o '"kernel_share": {[^}]*}' $f; done
