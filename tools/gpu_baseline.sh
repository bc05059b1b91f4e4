# baseline of the re-entered session: full GPU suite, default N=1 bench line, launch list of tools/prof_steps.py
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_tests_a.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests_a.log
timeout 600 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2_n1.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j.get("phases"), j.get("roofline"))
PY
timeout 300 python tools/prof_steps.py > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_prof_steps.csv python tools/prof_steps.py > gpurun_out/ncu_list.log 2>&1; echo "ncu rc=$?"
