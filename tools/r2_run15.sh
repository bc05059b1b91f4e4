cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t15.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_t15.log
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2_n1.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"], j.get("phases"), j.get("roofline"), j.get("bench_wall_s"))
PY
