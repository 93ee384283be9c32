# GroupNorm: parity tests of every path, then the same-box A/B of the streaming kernel against the cluster kernel
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "groupnorm" > gpurun_out/r2/gn_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2/gn_tests.log
kb() { timeout 300 python scripts/kbench.py --kernel gn --batch $1 --dtype fp16 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['kernel'], d.get('us_cold'), d.get('us_hot_l2'), d.get('frac_hbm_peak'))"; }
for B in ${GN_BATCHES:-104 13}; do
  echo "== B=$B stream occ2"; DADD_GN_IMPL=stream kb $B
  echo "== B=$B stream occ1"; DADD_GN_IMPL=stream DADD_GN_OCC=1 kb $B
  echo "== B=$B cluster"; DADD_GN_IMPL=cluster kb $B
done 2>&1 | tee gpurun_out/r2/gn_ab.log
