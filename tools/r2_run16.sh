cd /root/repo; mkdir -p gpurun_out
for c in 5 6; do KGAT_SPMM_CTAS=$c python tools/prof_cf.py 2>&1 | grep "epoch 2"; done
for c in 5 6; do KGAT_SPMM_CTAS=$c KGAT_SPMM_U=8 python tools/prof_cf.py 2>&1 | grep "epoch 2"; done
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_t16.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_t16.log
