# round 2, 2 GPUs: does the box have NVLS multicast; parity of the three exchanges; exchange timings (config 4 shape)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1
nvidia-smi --query-gpu=index,name,power.limit --format=csv > gpurun_out/r02a_gpus.txt 2>&1
timeout 240 $TR --master-port 29741 tests/dp_check.py --exchange multimem > gpurun_out/r02a_mc_small.log 2>&1; echo "multimem small rc=$?"
grep "dp_check\|multicast" gpurun_out/r02a_mc_small.log | tail -5
timeout 240 $TR --master-port 29742 tests/dp_check.py --tf32 1 --batch 512 --hidden 256 --inp 256 --steps 3 --iters 3 --exchange multimem > gpurun_out/r02a_mc_tf32.log 2>&1; echo "multimem tf32 rc=$?"
grep "dp_check world" gpurun_out/r02a_mc_tf32.log
timeout 240 $TR --master-port 29743 tests/dp_check.py --exchange fused > gpurun_out/r02a_fused_small.log 2>&1; echo "fused small rc=$?"
grep "dp_check world" gpurun_out/r02a_fused_small.log
timeout 240 $TR --master-port 29744 tests/dp_check.py --exchange nccl > gpurun_out/r02a_nccl_small.log 2>&1; echo "nccl small rc=$?"
grep "dp_check world" gpurun_out/r02a_nccl_small.log
for mm in off on; do
  timeout 300 $TR --master-port 29745 bench.py --gpus 2 --steps 20 --warmup 3 --multimem $mm --no-cpu-baseline > gpurun_out/r02a_bench_n2_mm_$mm.json 2> gpurun_out/r02a_bench_n2_mm_$mm.err; echo "bench multimem=$mm rc=$?"
done
c=128
CG_DP_CTAS=$c timeout 300 $TR --master-port 29746 bench.py --gpus 2 --steps 20 --warmup 3 --multimem off --no-cpu-baseline > gpurun_out/r02a_bench_n2_ipc_ctas$c.json 2> gpurun_out/r02a_bench_n2_ipc_ctas$c.err; echo "bench ipc ctas=$c rc=$?"
c=32
CG_DP_CTAS=$c timeout 300 $TR --master-port 29747 bench.py --gpus 2 --steps 20 --warmup 3 --multimem on --no-cpu-baseline > gpurun_out/r02a_bench_n2_mc_ctas$c.json 2> gpurun_out/r02a_bench_n2_mc_ctas$c.err; echo "bench mc ctas=$c rc=$?"
for f in gpurun_out/r02a_bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "allreduce_ms(serialised)", round(d["kernel_share"]["allreduce_ms"],3), "n", d["kernel_share"]["kernels"]["allreduce"], "|", d.get("exchange"), "|", d.get("multimem_supported"))
except Exception as e:
    print("  unreadable:", e)
PY
done
