cd /root/repo; mkdir -p gpurun_out
for w in 8 32 128; do KGAT_KG_WINDOW=$w python tools/prof_kg.py --kg 3000 --epochs 3 2>&1 | grep mode; done
KGAT_KG_ADAM=dense python tools/prof_kg.py --kg 3000 --epochs 2 2>&1 | grep mode
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 8000 -c 1200 --csv --log-file gpurun_out/r2_kg_rolling_launches.csv python tools/prof_kg.py --kg 1500 --epochs 2 > gpurun_out/ncu_kg.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_kg_rolling_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4][:60]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:60s} n={len(v):5d} avg={sum(v)/len(v)/1e3:8.2f} us  max={max(v)/1e3:8.2f}")
PY
