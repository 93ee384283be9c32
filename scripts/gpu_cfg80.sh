for c in 1 2; do
  echo "== cfg80 $c"
  DADD_ATTN_CFG80=$c timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep "^{" | sed -n 2p | cut -c1-150
  DADD_ATTN_CFG80=$c timeout 300 python scripts/kbench.py --kernel self_attn --batch 13 --dtype fp16 2>&1 | grep "^{" | sed -n 2p | cut -c1-150
  DADD_ATTN_CFG80=$c timeout 300 python scripts/kbench.py --kernel self_attn --batch 16 --res 64 --dtype fp16 2>&1 | grep "^{" | sed -n 2p | cut -c1-150
  DADD_ATTN_CFG80=$c DADD_ATTN_DEBUG=1 timeout 120 python - <<'PY' 2>&1 | tail -2
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from progressive_stable_diffusion_b200 import ops
for n, d, b in ((256, 80, 24), (1000, 80, 3)):
    h = 8; c = h * d
    g = torch.Generator().manual_seed(n + d)
    qkv = (torch.randn(b, n, 3 * c, generator=g) * 1.2).to(torch.float16)
    q, k, v = (qkv[..., i * c:(i + 1) * c].float().view(b, n, h, d).transpose(1, 2) for i in range(3))
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, n, c)
    qd = qkv.cuda()
    o = ops.self_attention(qd[..., :c], qd[..., c:2 * c], qd[..., 2 * c:], h, impl="tc")
    torch.cuda.synchronize()
    print(n, d, b, "rel err", ((o.float().cpu() - ref).abs().max() / ref.abs().max()).item())
PY
done
