"""GroupNorm+SiLU kernel error against fp64 for the two SiLU forms (run twice: DADD_SILU_EXACT=1 and unset)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progressive_stable_diffusion_b200 import ops
torch.manual_seed(0)
for dt in (torch.bfloat16, torch.float16):
    x = (torch.randn(8, 320, 32, 32) * 1.0).to(dt)
    g = 1 + 0.5 * torch.randn(320); b = 0.5 * torch.randn(320)
    g[:40] *= 6.0      # a few channels with wide pre-activations (|o| up to ~25)
    ref = F.silu(F.group_norm(x.double(), 32, g.double(), b.double(), 1e-5))
    y = ops.group_norm(x.cuda().contiguous(memory_format=torch.channels_last), g.cuda(), b.cuda(), 32, 1e-5, True).double().cpu()
    ideal = ref.to(dt).double()
    e, e0 = (y - ref).abs(), (ideal - ref).abs()
    print(f"{dt} mode={'exact' if os.environ.get('DADD_SILU_EXACT') else 'tanh'}: max abs err {e.max():.3e} (rounding alone {e0.max():.3e}), "
          f"rms err {e.pow(2).mean().sqrt():.3e} (rounding alone {e0.pow(2).mean().sqrt():.3e}), "
          f"rms err where ref<0: {e[ref<0].pow(2).mean().sqrt():.3e} (rounding alone {e0[ref<0].pow(2).mean().sqrt():.3e})")
