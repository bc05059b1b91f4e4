cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "api or graphed or deferred or engine_epoch or state_dict or trajectory or golden" > gpurun_out/r2_t14.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2_t14.log
python tools/prof_api.py --cf 50 --kg 4000 --top 12 2>&1 | grep -E "us/step|tottime|functions.py|optim.py|model.py|trainer.py" | head -14
