mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/kb_attn_mc.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "self_attention" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
for B in 26 104; do
for MC in 0 1; do
echo "== B=$B DADD_ATTN_MC=$MC" >> gpurun_out/kb_attn_mc.log
DADD_ATTN_MC=$MC timeout 300 python scripts/kbench.py --kernel self_attn --batch $B >> gpurun_out/kb_attn_mc.log 2>&1
done
done
DADD_ATTN_MC=1 timeout 300 python scripts/kbench.py --kernel self_attn --res 64 --batch 4 >> gpurun_out/kb_attn_mc.log 2>&1
DADD_ATTN_MC=0 timeout 300 python scripts/kbench.py --kernel self_attn --res 64 --batch 4 >> gpurun_out/kb_attn_mc.log 2>&1
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/kb_attn_mc.log
