cd /root/repo; mkdir -p gpurun_out
python tools/prof_cf.py 2>&1 | grep "epoch 2"
for u in 4 8; do KGAT_SPMM_U=$u python tools/prof_cf.py 2>&1 | grep "epoch 2"; done
timeout 600 python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t19.log 2>&1; echo "pruning tests rc=$?"; tail -2 gpurun_out/r2_t19.log
