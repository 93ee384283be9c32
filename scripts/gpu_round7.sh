mkdir -p gpurun_out
timeout 120 scripts/micro/exp_variants > gpurun_out/exp_variants.log 2>&1; echo "micro rc=$?"
cat gpurun_out/exp_variants.log
timeout 900 python bench.py > gpurun_out/bench_r7.log 2>gpurun_out/bench_r7.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench_r7.log
