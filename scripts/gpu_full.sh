mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/r2/pytest_full.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2/pytest_full.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r2/bench_n1.err
cat gpurun_out/r2/bench_n1.json
