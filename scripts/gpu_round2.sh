mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/kbench.py --kernel gn,ln,add_ln,bias_res,geglu > gpurun_out/kbench2.log 2>&1; echo "kbench rc=$?" >> gpurun_out/summary.txt
timeout 600 python scripts/profile_step.py --patients 2 > gpurun_out/profile_plain.log 2>&1; echo "profile rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/kbench2.log; tail -3 gpurun_out/profile_plain.log
