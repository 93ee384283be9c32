mkdir -p gpurun_out; rm -f gpurun_out/kb_gn4.log
run() { echo "== $1" >> gpurun_out/kb_gn4.log; env $1 timeout 300 python scripts/kbench.py --kernel gn --batch 104 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print(d['kernel'], d['us_cold'], d['us_hot_l2'], d['frac_hbm_peak'])
    except Exception: print(l.strip()[:200])" >> gpurun_out/kb_gn4.log; }
run "X=1"
run "DADD_GN_THREADS=256 DADD_GN_SLAB_KB=48 DADD_GN_SMAX=16"
run "DADD_GN_THREADS=256 DADD_GN_SLAB_KB=96 DADD_GN_SMAX=8"
run "DADD_GN_THREADS=512 DADD_GN_SLAB_KB=48 DADD_GN_SMAX=16"
run "DADD_GN_THREADS=384 DADD_GN_SLAB_KB=64 DADD_GN_SMAX=16"
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "groupnorm" > gpurun_out/pytest_gn.log 2>&1; echo "pytest default rc=$?" >> gpurun_out/kb_gn4.log
DADD_GN_THREADS=256 DADD_GN_SLAB_KB=48 DADD_GN_SMAX=16 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "groupnorm" > gpurun_out/pytest_gn16.log 2>&1; echo "pytest S16 rc=$?" >> gpurun_out/kb_gn4.log
cat gpurun_out/kb_gn4.log; tail -3 gpurun_out/pytest_gn16.log
