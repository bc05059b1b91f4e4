cd /root/repo; mkdir -p gpurun_out
for cfg in "U4 KGAT_SPMM_U=4" "U8 KGAT_SPMM_U=8" "U4grid KGAT_SPMM_U=4 KGAT_L1_GRID=1" "U8grid KGAT_SPMM_U=8 KGAT_L1_GRID=1"; do
set -- $cfg; name=$1; shift
env "$@" python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --no-e2e > gpurun_out/r2_ab_$name.json 2> gpurun_out/r2_ab_$name.err; echo "$name rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_ab_$name.json').read().strip().splitlines()[-1])
print("$name", d['phases']['cf_step_us'], {k:round(v['avg_us'],1) for k,v in d['kernels'].items() if k.startswith('spmm')})
PY
done
