mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=120 -p no:cacheprovider -k "self_attention_core and tc and bf16 and 256-80" > gpurun_out/tc_probe1.log 2>&1; echo "probe1 rc=$?"
tail -25 gpurun_out/tc_probe1.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider -k "self_attention" > gpurun_out/tc_probe2.log 2>&1; echo "probe2 rc=$?"
tail -30 gpurun_out/tc_probe2.log
