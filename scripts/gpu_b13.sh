mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -q -x --timeout=900 -p no:cacheprovider > gpurun_out/r2/pytest_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2/pytest_train.log
for p in 1 2; do
timeout 600 python bench.py --steps 3 --warmup 3 --patients $p --no-extras > gpurun_out/r2/bench_p$p.json 2> gpurun_out/r2/bench_p$p.err; echo "bench p=$p rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2/bench_p$p.json')); print('patients $p', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_bench_step_B13.csv python bench.py --steps 1 --warmup 3 --patients 1 --profile-step --no-extras > gpurun_out/r2/bench_under_ncu_b13.log 2>&1; echo "ncu rc=$?"
