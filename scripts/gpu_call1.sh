mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/r2/pytest1.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2/pytest1.log
for dt in fp16 bf16; do timeout 600 python scripts/kbench.py --batch 104 --dtype $dt > gpurun_out/r2/kbench_base_B104_$dt.jsonl 2>&1; done
timeout 600 python scripts/kbench.py --batch 13 --dtype fp16 > gpurun_out/r2/kbench_base_B13_fp16.jsonl 2>&1
cat gpurun_out/r2/kbench_base_B104_fp16.jsonl
