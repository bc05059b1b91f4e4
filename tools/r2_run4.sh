cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_pruning.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "frontier or pruned or biagg or golden or trajectory" > gpurun_out/r2_t7.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_t7.log
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2_b7.json 2> gpurun_out/r2_b7.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b7.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"], j["e2e"]["value"], j["phases"])
        ks=j["kernels"]
        print("   ", {k:v["avg_us"] for k,v in list(ks.items())[:40]})
    except Exception as e: print(f, "ERR", e)
PY
