cd /root/repo; mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t24_$i.log 2>&1; echo "pruning tests run $i rc=$?"; tail -1 gpurun_out/r2_t24_$i.log; done
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_t24.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_t24.log
