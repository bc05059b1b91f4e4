cd /root/repo; mkdir -p gpurun_out
python tools/prof_api.py --cf 50 --kg 4000 --top 14 2>&1 | grep -E "us/step|tottime|functions.py|optim.py|model.py|trainer.py" | head -30
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t13.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/r2_t13.log
KGAT_PLAN_SORT=1 timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_t13s.log 2>&1; echo "all gpu tests (sorted plan) rc=$?"; tail -12 gpurun_out/r2_t13s.log | grep -E "FAILED|passed|failed"
