cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_pruning.py -q -m gpu -p no:cacheprovider -k "range_sharded" -x > gpurun_out/r2_rs1.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/r2_rs1.log
