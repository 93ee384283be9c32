mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/kb_ff1b.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "ff_geglu" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
for B in 26 104; do
timeout 300 python scripts/kbench.py --kernel ff1 --batch $B >> gpurun_out/kb_ff1b.log 2>&1
done
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; grep fused gpurun_out/kb_ff1b.log
