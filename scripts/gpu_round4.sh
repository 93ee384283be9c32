mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=900 -p no:cacheprovider -k "groupnorm" > gpurun_out/pytest_gn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
(DADD_SILU_EXACT=1 python scripts/exp_silu.py; python scripts/exp_silu.py) > gpurun_out/exp_silu.log 2>&1
(DADD_SILU_EXACT=1 python __graft_entry__.py --smoke; python __graft_entry__.py --smoke) >> gpurun_out/exp_silu.log 2>&1
DADD_SILU_EXACT=1 timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/model_exact.log 2>&1; echo "model exact rc=$?" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -s --timeout=900 -p no:cacheprovider > gpurun_out/model_tanh.log 2>&1; echo "model tanh rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/kbench.py --kernel gn > gpurun_out/kbench3.log 2>&1; echo "kbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/profile_step.py --patients 2 > gpurun_out/profile_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python scripts/profile_step.py --patients 2 > gpurun_out/profile_ncu.log 2>&1; echo "profile rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gn.log; cat gpurun_out/exp_silu.log; grep -h "PSNR\|rel err" gpurun_out/model_exact.log gpurun_out/model_tanh.log; cat gpurun_out/kbench3.log; tail -3 gpurun_out/profile_plain.log
