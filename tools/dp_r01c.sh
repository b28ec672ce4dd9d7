# round-1 (third pass), 2 GPUs: parity of the fused all-reduce + update kernel (and of the NCCL exchange), then the
# weak-scaling bench with either exchange
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 240 $TR --master-port 29701 tests/dp_check.py --tf32 0 --batch 64 --hidden 48 --inp 40 --steps 2 --iters 3 --exchange fused > gpurun_out/dpc_fused_small.log 2>&1; echo "fused small rc=$?"
tail -3 gpurun_out/dpc_fused_small.log
timeout 240 $TR --master-port 29702 tests/dp_check.py --tf32 1 --batch 512 --hidden 256 --inp 256 --steps 2 --iters 3 --exchange fused > gpurun_out/dpc_fused_tf32.log 2>&1; echo "fused tf32 rc=$?"
tail -2 gpurun_out/dpc_fused_tf32.log
timeout 240 $TR --master-port 29703 tests/dp_check.py --tf32 0 --batch 64 --hidden 48 --inp 40 --steps 2 --iters 3 --exchange nccl > gpurun_out/dpc_nccl_small.log 2>&1; echo "nccl small rc=$?"
tail -2 gpurun_out/dpc_nccl_small.log
timeout 300 $TR --master-port 29704 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err; echo "bench fused rc=$?"
CG_DP_FUSED=0 timeout 300 $TR --master-port 29705 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "bench nccl rc=$?"
CG_DP_CTAS=128 timeout 300 $TR --master-port 29706 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_fused_c128.json 2> gpurun_out/bench_n2_fused_c128.err; echo "bench fused 128 ctas rc=$?"
CG_DP_CTAS=32 timeout 300 $TR --master-port 29707 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_fused_c32.json 2> gpurun_out/bench_n2_fused_c32.err; echo "bench fused 32 ctas rc=$?"
for f in gpurun_out/bench_n2_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"kernel_share": {[^}]*}' $f; done
