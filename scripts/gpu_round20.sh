mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=600 -p no:cacheprovider -s -k "quick_gelu or clip_tower or image_projections or front_end or purifier" > gpurun_out/pytest_fe.log 2>&1; echo "pytest front end rc=$?" >> gpurun_out/summary.txt
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x --timeout=900 -p no:cacheprovider -s -k "512 or sweep" > gpurun_out/pytest_model2.log 2>&1; echo "pytest model2 rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "rel err|passed|failed|Error|^E " gpurun_out/pytest_fe.log | head -20; grep -E "rel err|passed|failed|Error|^E " gpurun_out/pytest_model2.log | head
