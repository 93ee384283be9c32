mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm" > gpurun_out/r2/gn_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2/gn_tests.log
