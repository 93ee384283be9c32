mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/kb_gn2.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "groupnorm or linear" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
for B in 26 104; do
timeout 300 python scripts/kbench.py --kernel gn --batch $B >> gpurun_out/kb_gn2.log 2>&1
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "model or smoke" > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?" >> gpurun_out/summary.txt
for P in 2 8; do
timeout 600 python scripts/profile_step.py --patients $P > gpurun_out/profile_p$P.log 2>&1; echo "profile P=$P rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_model.log; cat gpurun_out/kb_gn2.log; cat gpurun_out/profile_p2.log gpurun_out/profile_p8.log
