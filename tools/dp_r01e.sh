# round-1 (fifth pass), 2 GPUs: the fused exchange with the hoisted tape
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29711 tests/dp_check.py --tf32 1 --batch 512 --hidden 256 --inp 256 --steps 3 --iters 3 --exchange fused > gpurun_out/dpe_fused_tf32.log 2>&1; echo "fused tf32 rc=$?"
tail -2 gpurun_out/dpe_fused_tf32.log
timeout 300 $TR --master-port 29714 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_e_n2_hoist1.json 2> gpurun_out/bench_e_n2_hoist1.err; echo "bench hoist rc=$?"
CG_HOIST=0 timeout 300 $TR --master-port 29715 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_e_n2_hoist0.json 2> gpurun_out/bench_e_n2_hoist0.err; echo "bench nohoist rc=$?"
CG_DP_FUSED=0 timeout 300 $TR --master-port 29716 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_e_n2_nccl_hoist1.json 2> gpurun_out/bench_e_n2_nccl_hoist1.err; echo "bench nccl hoist rc=$?"
CG_DP_RESERVE_SMS=8 timeout 300 $TR --master-port 29717 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_e_n2_hoist1_res8.json 2> gpurun_out/bench_e_n2_hoist1_res8.err; echo "bench reserve8 rc=$?"
for f in gpurun_out/bench_e_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"e2e": {[^}]*}' $f; grep -o '"kernel_share": {[^}]*}' $f; done
