cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "kg or transr or engine or deferred or rolling or golden or smoke" > gpurun_out/r2_t25.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_t25.log
python tools/prof_kg.py --kg 3000 --epochs 3 2>&1 | grep mode | tail -1
