cd /root/repo; mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2_b6_c5.json 2> gpurun_out/r2_b6_c5.err; echo "bench ctas5 rc=$?"
KGAT_SPMM_CTAS=4 python bench.py --steps 3 --warmup 3 --no-extra > gpurun_out/r2_b6_c4.json 2> gpurun_out/r2_b6_c4.err; echo "bench ctas4 rc=$?"
python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches6.csv python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/r2_ncu6.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_b6_c5.json","gpurun_out/r2_b6_c4.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["value"], j.get("e2e",{}).get("value"), {k:v for k,v in j.items() if k in ("cf_step_us","kg_step_us","phases","breakdown")})
    except Exception as e: print(f, "ERR", e)
PY
python - <<'PY'
import json
for f in ("gpurun_out/r2_b6_c5.json","gpurun_out/r2_b6_c4.json"):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["phases"])
        ks=j["kernels"]
        print("   ", {k:(v if not isinstance(v,dict) else v.get("us") or v) for k,v in list(ks.items())[:16]})
    except Exception as e: print(f, "ERR", e)
PY
