set -x
cd /root/repo
mkdir -p gpurun_out
python tools/r2_debug1.py > gpurun_out/r2_dbg1.log 2>&1; tail -40 gpurun_out/r2_dbg1.log
python -m pytest tests/test_gpu_pruning.py -q -m gpu > gpurun_out/r2_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_t1.log | tail -20
python -m pytest tests/test_gpu_parity.py -q -m gpu > gpurun_out/r2_t2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2_t2.log | tail -30
