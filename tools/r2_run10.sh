cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "rolling or lazy or adam or engine_epoch or engine_kg or trajectory or sharded_engine_world1" > gpurun_out/r2_t10.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_t10.log
for f in 0 1; do for w in 16 32; do KGAT_KG_FORK=$f KGAT_KG_WINDOW=$w python tools/prof_kg.py --kg 3000 --epochs 3 2>&1 | grep mode | tail -1 | sed "s/^/fork=$f /"; done; done
timeout 600 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-hbm-regime > gpurun_out/r2_b10.json 2> gpurun_out/r2_b10.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads(open("gpurun_out/r2_b10.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j["phases"], j["cf_loss"], j["kg_loss"])
PY
