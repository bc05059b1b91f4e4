cd /root/repo; mkdir -p gpurun_out
for p in 0 1 0 1; do KGAT_BIAGG_PREFETCH=$p python tools/prof_cf.py 2>&1 | grep "epoch 2"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "biagg or cf_loss or engine_epoch" > gpurun_out/r2_t20.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_t20.log
