mkdir -p gpurun_out/r2
./scripts/micro/mma_rate 2>&1 | tee gpurun_out/r2/mma_rate.log
export DADD_ATTN_DEBUG=1
for shape in "256 80 2" "1024 80 1" "384 72 1" "256 160 2"; do
  set -- $shape
  echo "== N=$1 d=$2 b=$3"
  timeout 120 python - "$@" <<'PY' 2>&1 | grep -v "^  File\|^    \|Traceback" | tail -6
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
