cd /root/repo; mkdir -p gpurun_out
for s in 0 1; do for h in 0 4; do KGAT_PLAN_SORT=$s KGAT_SPMM_HALF=$h python tools/prof_cf.py 2>&1 | grep "epoch 2"; done; done
KGAT_PLAN_SORT=1 KGAT_SPMM_HALF=4 timeout 600 python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_t12.log 2>&1; echo "pruning tests (sorted plan, half) rc=$?"; tail -4 gpurun_out/r2_t12.log
