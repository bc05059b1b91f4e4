cd /root/repo; mkdir -p gpurun_out
for p in 0 1 0 1; do KGAT_PDL=$p python tools/prof_kg.py --kg 3000 --epochs 3 2>&1 | grep mode | tail -1 | sed "s/^/PDL=$p /"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "rolling or adam or engine_epoch or engine_kg or deferred or kg_loss or transr" > gpurun_out/r2_t21.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_t21.log
