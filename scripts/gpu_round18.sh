mkdir -p gpurun_out; rm -f gpurun_out/summary.txt gpurun_out/kb_gn3.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout=300 -p no:cacheprovider -k "groupnorm or linear" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/summary.txt
for B in 26 52 104; do
for F in 0 1; do
echo "== B=$B DADD_GN_FLAT=$F" >> gpurun_out/kb_gn3.log
DADD_GN_FLAT=$F timeout 300 python scripts/kbench.py --kernel gn --batch $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['kernel'], d['us_cold'], d['us_hot_l2'])
    except Exception: print(l.strip())" >> gpurun_out/kb_gn3.log
done
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider -k "model or smoke" > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?" >> gpurun_out/summary.txt
for P in 2 8; do
timeout 600 python scripts/profile_step.py --patients $P > gpurun_out/profile_p$P.log 2>&1; echo "profile P=$P rc=$?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -8 gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_model.log; cat gpurun_out/kb_gn3.log; cat gpurun_out/profile_p2.log gpurun_out/profile_p8.log
