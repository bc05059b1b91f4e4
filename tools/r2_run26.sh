cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_t26.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_t26.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j.get("phases"), j.get("bench_wall_s"))
PY
