cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "rolling or lazy or engine_epoch or engine_kg" > gpurun_out/r2_t7.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_t7.log
for mode in dense rolling; do
KGAT_KG_ADAM=$mode timeout 600 python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline --no-hbm-regime > gpurun_out/r2_b7_$mode.json 2> gpurun_out/r2_b7_$mode.err; echo "bench $mode rc=$?"
python - <<PY
import json
j=json.loads(open("gpurun_out/r2_b7_$mode.json").read().strip().splitlines()[-1]); print("$mode", j["value"], j["e2e"]["value"], j["phases"], j["cf_loss"], j["kg_loss"])
PY
done
