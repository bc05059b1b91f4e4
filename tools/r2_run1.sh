set -x
cd /root/repo
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pruning.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t1.log
tail -30 gpurun_out/r2_t1.log
python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_t2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_t2.log
tail -15 gpurun_out/r2_t2.log
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-hbm-regime > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_b1.json; tail -5 gpurun_out/r2_b1.err
