mkdir -p gpurun_out
timeout 120 scripts/micro/tma_rows > gpurun_out/tma_rows.log 2>&1; echo "micro rc=$?"
cat gpurun_out/tma_rows.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "self_attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_attn.log
timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 2>&1 | tail -4
