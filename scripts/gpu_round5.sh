mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 300 python scripts/kbench.py --kernel self_attn > gpurun_out/kb_attn.log 2>&1
for P in 2 4 8; do
timeout 900 python bench.py --steps 2 --warmup 3 --patients $P > gpurun_out/bench_p$P.log 2>&1; echo "bench P=$P rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -4 gpurun_out/pytest_gpu.log; head -2 gpurun_out/kb_attn.log
for P in 2 4 8; do tail -1 gpurun_out/bench_p$P.log | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ('value','ms_per_step','e2e','unet_tflops_achieved','clocks')}, d['cpu_baseline']['value'] if 'cpu_baseline' in d else None)"; done
