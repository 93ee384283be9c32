mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=120 -p no:cacheprovider -k "self_attention" > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
for PL in 0 2 3 4; do
echo "== POLY=$PL" >> gpurun_out/kb_attn_poly.log
DADD_ATTN_POLY=$PL timeout 300 python scripts/kbench.py --kernel self_attn >> gpurun_out/kb_attn_poly.log 2>&1
done
echo "== POLY=3 B=104" >> gpurun_out/kb_attn_poly.log
DADD_ATTN_POLY=3 timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 >> gpurun_out/kb_attn_poly.log 2>&1
cat gpurun_out/summary.txt; tail -5 gpurun_out/pytest_attn.log; cat gpurun_out/kb_attn_poly.log
