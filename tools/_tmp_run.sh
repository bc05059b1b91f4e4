cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t29.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_t29.log
python tools/prof_api.py --cf 200 --kg 6000 --top 8 2>&1 | grep -E "us/step" | head -4
