mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "groupnorm or upsample or model or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
for P in 2 8; do
timeout 600 python scripts/profile_step.py --patients $P > gpurun_out/profile_p$P.log 2>&1; echo "profile P=$P rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/profile_p2.log gpurun_out/profile_p8.log
