cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t23.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_t23.log
python tools/prof_kg.py --kg 3000 --epochs 3 2>&1 | grep mode | tail -1
python tools/prof_cf.py 2>&1 | grep "epoch 2"
