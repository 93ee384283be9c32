mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=120 -p no:cacheprovider -k "self_attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/kbench.py --kernel self_attn > gpurun_out/kb_attn.log 2>&1; echo "kbench rc=$?" >> gpurun_out/summary.txt
DADD_SELF_ATTN=tc1 timeout 300 python scripts/kbench.py --kernel self_attn > gpurun_out/kb_attn_v1.log 2>&1
timeout 300 python scripts/kbench.py --kernel self_attn --res 64 --batch 4 > gpurun_out/kb_attn_512.log 2>&1
cat gpurun_out/summary.txt; tail -15 gpurun_out/pytest_attn.log; cat gpurun_out/kb_attn.log; echo v1; cat gpurun_out/kb_attn_v1.log; echo 512; cat gpurun_out/kb_attn_512.log
