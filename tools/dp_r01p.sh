# round-1 (11th pass), 2 GPUs: data parallel under the exact-DAG capture (both exchanges), weak-scaling point
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29741 tests/dp_check.py --tf32 1 --batch 512 --hidden 256 --inp 256 --steps 3 --iters 3 --exchange fused > gpurun_out/dpp_fused_tf32.log 2>&1; echo "fused tf32 rc=$?"
grep "dp_check world" gpurun_out/dpp_fused_tf32.log
timeout 240 $TR --master-port 29742 tests/dp_check.py --tf32 0 --batch 64 --hidden 48 --inp 40 --steps 2 --iters 3 --exchange nccl > gpurun_out/dpp_nccl_small.log 2>&1; echo "nccl small rc=$?"
grep "dp_check world" gpurun_out/dpp_nccl_small.log
timeout 300 $TR --master-port 29744 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_p_n2.json 2> gpurun_out/bench_p_n2.err; echo "bench rc=$?"
CG_GRAPH_EXACT=0 timeout 300 $TR --master-port 29745 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_p_n2_streams.json 2> gpurun_out/bench_p_n2_streams.err; echo "bench streams rc=$?"
for f in gpurun_out/bench_p_*.json; do echo $f; grep -o '"value": [0-9.]*, "unit": "iterations/s", "n_gpus"' $f | head -1; grep -o '"e2e": {[^}]*}' $f; done
