cd /root/repo; mkdir -p gpurun_out
for v in 0 1; do KGAT_EXPAND_PLAIN=$v python tools/prof_frontier.py 2>&1 | tail -2; done | tee gpurun_out/r2_frontier_variants_b.txt
python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t8.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t8.log
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2_b8.json 2> gpurun_out/r2_b8.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r2_b8.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j["phases"])
PY
