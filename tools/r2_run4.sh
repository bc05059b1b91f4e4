cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/r2_full4.log 2>&1; echo "full gpu suite rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_full4.log | tail -12
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2_b6_c5.json 2> gpurun_out/r2_b6_c5.err; echo "bench ctas5 rc=$?"
KGAT_SPMM_CTAS=4 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2_b6_c4.json 2> gpurun_out/r2_b6_c4.err; echo "bench ctas4 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b6_c5.json","gpurun_out/r2_b6_c4.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"], j.get("e2e",{}).get("value"), j.get("phases") or j.get("breakdown"))
        ks=j.get("kernels") or {}
        for k,v in sorted(ks.items(), key=lambda kv:-kv[1] if isinstance(kv[1],(int,float)) else 0)[:14]: print("   ",k,v)
    except Exception as e: print(f, "ERR", e)
PY
