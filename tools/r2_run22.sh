cd /root/repo; mkdir -p gpurun_out
for p in 0 1 0 1; do KGAT_PDL=$p python tools/prof_kg.py --kg 3000 --epochs 3 2>&1 | grep mode | tail -1 | sed "s/^/PDL(wait only)=$p /"; done
