mkdir -p gpurun_out/r2
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2/bench_n1.json 2> gpurun_out/r2/bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2/bench_n1.json'))
for k in ('value','e2e','strong','config3','config4','config5','parity','roofline'):
    print(k, json.dumps(d.get(k))[:300])
PY
