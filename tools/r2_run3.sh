cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_pruning.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_plain_1.log 2>&1; echo "plain run rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_plain_1.log | tail -8
KGAT_POISON=1 python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_poison_1.log 2>&1; echo "poison run rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_poison_1.log | tail -8
for pm in 600 1000 300; do
KGAT_SPMM_STATIC_PERMILLE=$pm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hbm-regime --no-e2e > gpurun_out/r2_b4_$pm.json 2> gpurun_out/r2_b4_$pm.err; echo "bench $pm rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_b4_$pm.json').read().strip().splitlines()[-1])
print($pm, {k:d[k] for k in ('value','phases')})
PY
done
python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/r2_ncu.log 2>&1; echo "ncu rc=$?"
