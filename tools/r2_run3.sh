cd /root/repo; mkdir -p gpurun_out
python tools/r2_debug7.py 2>&1 | grep "^sync" | cut -c1-330
for i in 1 2; do
python -m pytest tests/test_gpu_pruning.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_plain_$i.log 2>&1; echo "plain run $i rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_plain_$i.log | tail -8
done
KGAT_POISON=1 python -m pytest tests/test_gpu_pruning.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_poison_1.log 2>&1; echo "poison run rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_poison_1.log | tail -8
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hbm-regime > gpurun_out/r2_b2.json 2> gpurun_out/r2_b2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_b2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','phases','gpu_launches')}); print(d['e2e'])
PY
