mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=900 -p no:cacheprovider -k "groupnorm" > gpurun_out/pytest_gn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/kbench.py --kernel gn > gpurun_out/kbench3.log 2>&1; echo "kbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/profile_step.py --patients 2 > gpurun_out/profile_plain.log 2>&1; echo "profile rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gn.log; cat gpurun_out/kbench3.log; tail -3 gpurun_out/profile_plain.log
