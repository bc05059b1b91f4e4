cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2_n1.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j.get("phases"), j.get("bench_wall_s"))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_prof_steps.csv python tools/prof_steps.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_rows_kernel|biagg_fwd_tc5|biagg_bwd_mma|adam_kernel|spmm_scatter|frontier_expand" -s 34 -c 16 -o gpurun_out/r2_full_cf python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/ncu_full_cf.log 2>&1; echo "ncu full cf rc=$?"
ncu -i gpurun_out/r2_full_cf.ncu-rep --page raw --csv > gpurun_out/r2_full_cf_raw.csv 2>/dev/null; wc -c gpurun_out/r2_full_cf_raw.csv; rm -f gpurun_out/r2_full_cf.ncu-rep
