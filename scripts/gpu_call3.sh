mkdir -p gpurun_out/r2
export DADD_ATTN_DEBUG=1
for shape in "256 80 2" "1024 80 1" "256 160 2" "512 80 2" "384 72 1"; do
  set -- $shape
  echo "== N=$1 d=$2 b=$3"
  timeout 120 python - "$@" <<'PY' 2>&1 | tail -4
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from progressive_stable_diffusion_b200 import ops
n, d, b = map(int, sys.argv[1:4]); h = 8; c = h * d
g = torch.Generator().manual_seed(n + d)
qkv = (torch.randn(b, n, 3 * c, generator=g) * 1.2).to(torch.float16)
q, k, v = (qkv[..., i * c:(i + 1) * c].float().view(b, n, h, d).transpose(1, 2) for i in range(3))
ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
qd = qkv.cuda()
o = ops.self_attention(qd[..., :c], qd[..., c:2 * c], qd[..., 2 * c:], h, impl="tc")
torch.cuda.synchronize()
print("rel err", ((o.float().cpu() - ref).abs().max() / ref.abs().max()).item())
PY
done
unset DADD_ATTN_DEBUG
echo "== trace POLY=0"
DADD_ATTN_POLY=0 DADD_ATTN_TRACE=gpurun_out/r2/attn_trace_p0.txt timeout 120 python scripts/kbench.py --kernel self_attn --batch 26 --iters 1 --no-flush --dtype fp16 2>&1 | tail -2
python scripts/attn_trace.py gpurun_out/r2/attn_trace_p0.txt
echo "== kbench POLY=0 B=104"
DADD_ATTN_POLY=0 timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep "^{" | cut -c1-200
