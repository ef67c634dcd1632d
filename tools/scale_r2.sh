# Multi-GPU evidence of round 2 (run under `gpurun --gpus N -- bash tools/scale_r2.sh N`).
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err; echo "bench exit $?"
tail -c 400 gpurun_out/scale_${N}.err
timeout 300 $TR --master-port 29517 tools/check_multi_gpu.py 2>&1 | tail -1 | tee gpurun_out/check_multi_gpu_${N}.json
for ex in resident nccl; do timeout 300 $TR --master-port 29541 tools/bench_root_parallel.py --exchange $ex --seeds 8 2>&1 | tail -1 | tee gpurun_out/root_parallel_${N}_$ex.json; done
timeout 300 $TR --master-port 29543 tools/bench_root_parallel.py --exchange resident --seeds 2 --n-sim 100000 --reps 2 2>&1 | tail -1 | tee gpurun_out/root_parallel_N100k_${N}_resident.json
python - <<PY
import json
d=json.loads(open("gpurun_out/scale_${N}.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","ms_per_step_fastest_rank","scaling","search_mode","solved","shard_equals_single_gpu"): print(k, d[k])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("weak", d["weak"])
for k in ("root_parallel","puct_root_parallel","selfplay"): print(k, json.dumps(d.get(k))[:500])
PY
