# Round-2 (second half) evidence: ncu --set full captures (cold caches) of the kernels added this half: streaming GroupNorm, the
# CTA-pair linear GEMM, the wide-head attention.  Summaries -> gpurun_out/r2/ncu/*.txt (copied to profiles/ by hand).
mkdir -p gpurun_out/r2/ncu
O=gpurun_out/r2/ncu
cap() {  # name, kernel regex, kbench kernel, env
  env $4 timeout 600 ncu --set full --import-source on --cache-control all --clock-control none -k "regex:$2" -c 1 -f -o $O/$1 python scripts/kbench.py --kernel $3 --batch 104 --iters 1 --no-flush --dtype fp16 > $O/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python scripts/ncu_pick.py $O/$1.ncu-rep > $O/$1.txt 2>&1
  rm -f $O/$1.ncu-rep
}
cap gn_stream gn_stream_kernel gn DADD_GN_IMPL=stream
cap linear_pair linear_kernel lin DADD_LIN_PAIR=1
cap linear_single linear_kernel lin DADD_LIN_PAIR=0
timeout 600 ncu --set full --import-source on --cache-control all --clock-control none -k "regex:self_attn_wide" -c 1 -f -o $O/attn_wide python - > $O/ncu_attn_wide.log 2>&1 <<'PY'
import torch
from progressive_stable_diffusion_b200 import ops
q, k, v = (torch.randn(104, 1024, 512, device="cuda", dtype=torch.float16) * 0.3 for _ in range(3))
ops.self_attention(q, k, v, 1); torch.cuda.synchronize()
PY
echo "ncu attn_wide rc=$?"; python scripts/ncu_pick.py $O/attn_wide.ncu-rep > $O/attn_wide.txt 2>&1; rm -f $O/attn_wide.ncu-rep
ls -la $O
