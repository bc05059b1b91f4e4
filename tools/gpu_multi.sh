# usage: bash tools/gpu_multi.sh N [check]
N=$1
cd /root/repo; mkdir -p gpurun_out
if [ "$2" = "check" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/sharded_pruned_check.py > gpurun_out/r2_check_n$N.log 2>&1; echo "check rc=$?"
grep -E "world=|Error|error|Traceback" gpurun_out/r2_check_n$N.log | tail -12
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
    for k in ('value','phases','parity_max_rel_err','replicas_equal','exchange','row_ranges','c5_scaled'):
        print(k, json.dumps(d.get(k))[:700])
    print('e2e', json.dumps(d.get('e2e'))[:200])
except Exception as e: print("no json", e)
PY
