mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "self_attention" --timeout=600 -p no:cacheprovider > gpurun_out/r2/pytest_attn2.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r2/pytest_attn2.log
for poly in 0 1 2 3 4; do
  echo "== tc3 POLY=$poly"
  DADD_ATTN_POLY=$poly timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep -v "^$" | cut -c1-200
done
echo "== tc2"
DADD_SELF_ATTN=tc2 timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | cut -c1-200
echo "== tc3 POLY=2 res64 B=16"
timeout 300 python scripts/kbench.py --kernel self_attn --batch 16 --res 64 --dtype fp16 2>&1 | cut -c1-200
echo "== tc3 POLY=2 B=13"
timeout 300 python scripts/kbench.py --kernel self_attn --batch 13 --dtype fp16 2>&1 | cut -c1-200
