# round-1 (sixth pass), 8 GPUs: parity of the fused exchange at world = 8, then the weak-scaling bench with either exchange
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29721 tests/dp_check.py --tf32 0 --batch 64 --hidden 48 --inp 40 --steps 2 --iters 3 --exchange fused > gpurun_out/dpf_fused_small_n8.log 2>&1; echo "fused small n8 rc=$?"
grep -E "dp_check|Error|error" gpurun_out/dpf_fused_small_n8.log | tail -4
timeout 400 $TR --master-port 29724 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_f_n8_fused.json 2> gpurun_out/bench_f_n8_fused.err; echo "bench fused rc=$?"
CG_DP_FUSED=0 timeout 400 $TR --master-port 29725 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_f_n8_nccl.json 2> gpurun_out/bench_f_n8_nccl.err; echo "bench nccl rc=$?"
for f in gpurun_out/bench_f_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"e2e": {[^}]*}' $f; grep -o '"kernel_share": {[^}]*}' $f; done
tail -3 gpurun_out/bench_f_n8_fused.err
