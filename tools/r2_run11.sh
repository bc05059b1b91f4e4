cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t11.log 2>&1; echo "pruning tests rc=$?"; tail -3 gpurun_out/r2_t11.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "spmm or cf_loss or propag or engine" > gpurun_out/r2_t11b.log 2>&1; echo "parity tests rc=$?"; tail -3 gpurun_out/r2_t11b.log
for h in 0 4 5; do for u in 4 8; do KGAT_SPMM_HALF=$h KGAT_SPMM_U=$u python tools/prof_cf.py 2>&1 | grep "epoch 2"; done; done
