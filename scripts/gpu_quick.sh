mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "layernorm or groupnorm_expanded" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=900 -p no:cacheprovider -s > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?" >> gpurun_out/summary.txt
for P in 2 8; do
timeout 600 python scripts/profile_step.py --patients $P > gpurun_out/profile_p$P.log 2>&1; echo "profile P=$P rc=$?" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 2 > gpurun_out/bench_r22.json 2> gpurun_out/bench_r22.err; echo "bench rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu.log; grep -E "rel err|PSNR|passed|failed" gpurun_out/pytest_model.log | tail -22; cat gpurun_out/profile_p2.log gpurun_out/profile_p8.log; wc -l gpurun_out/bench_r22.json; cut -c1-330 gpurun_out/bench_r22.json
