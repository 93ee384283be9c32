mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1800 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_r21.json 2> gpurun_out/bench_ref_r21.err; echo "bench ref rc=$?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench_r21.json 2> gpurun_out/bench_r21.err; echo "bench rc=$?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/pytest_gpu_full.log; tail -1 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_ref_r21.json; cut -c1-3500 gpurun_out/bench_r21.json; tail -2 gpurun_out/bench_r21.err
# ncu --set full captures of the hand-written hot kernels at the bench batch (B = 104)
cap() {  # name, kernel regex, kbench kernel
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$2" -c 1 -f -o gpurun_out/r01_$1 python scripts/kbench.py --kernel $3 --batch 104 --iters 1 --no-flush > gpurun_out/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
}
cap self_attn_tc2 self_attn_tc2 self_attn
cap gn_cluster gn_cluster gn
cap cross_attn_tc cross_attn_tc cross
cap ff_geglu ff_geglu ff1
