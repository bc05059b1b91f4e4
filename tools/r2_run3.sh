cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/ -q -m gpu -p no:cacheprovider > gpurun_out/r2_full.log 2>&1; echo "full gpu suite rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2_full.log | tail -12
python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/r2_prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"spmm_rows_kernel|spmm_task_kernel|biagg_fwd_tc5|biagg_bwd_mma|adam_kernel|spmm_scatter" -s 34 -c 14 -o gpurun_out/r2_prof_cf python tools/prof_steps.py --cf 2 --kg 2 > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_ncu_full.log
ls -la gpurun_out/r2_prof_cf.ncu-rep
