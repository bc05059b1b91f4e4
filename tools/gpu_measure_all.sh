# round-2 final measurements: default bench line, launch list, ncu --set full of the CF and KG kernels (exports only)
cd /root/repo; mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2_n1.err | cut -c1-300
python - <<'PY'
import json
j=json.loads(open("gpurun_out/bench_r2_n1.json").read().strip().splitlines()[-1]); print(j["value"], j["e2e"]["value"], j.get("phases"), j.get("bench_wall_s"))
PY
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err; echo "ref arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_prof_steps.csv python tools/prof_steps.py > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_rows_kernel|biagg_fwd_tc5|biagg_bwd_mma|adam_kernel|spmm_scatter|frontier_expand" -s 34 -c 16 -o gpurun_out/r2_full_cf python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/ncu_full_cf.log 2>&1; echo "ncu full cf rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"adam_rolling|transr_bwd" --launch-skip 9000 -c 6 -o gpurun_out/r2_full_kg python tools/prof_kg.py --kg 3200 --epochs 1 > gpurun_out/ncu_full_kg.log 2>&1; echo "ncu full kg rc=$?"
for f in r2_full_cf r2_full_kg; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; wc -c gpurun_out/${f}_raw.csv; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 18000 -c 600 --csv --log-file gpurun_out/r2_kg_rolling_launches.csv python tools/prof_kg.py --kg 3200 --epochs 1 > gpurun_out/ncu_kg.log 2>&1; echo "ncu kg list rc=$?"
