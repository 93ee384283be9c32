"""VAE decode of the bench batch (104 latents -> 256x256 images) in slices of S images: time per image vs S."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_stable_diffusion_b200 as P  # noqa: E402
from progressive_stable_diffusion_b200.inference_pipeline_ip import _latents_to_images  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
module = P.DiffusionModuleWithIP(P.default_config()).to(dev).eval()
lat = torch.randn(104, 4, 32, 32, device=dev)
with torch.no_grad():
    for s in (104, 52, 26, 13, 8, 4, 2):
        def run():
            return torch.cat([_latents_to_images(module, lat[i:i + s]) for i in range(0, 104, s)])
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            out = run()
        e1.record()
        torch.cuda.synchronize()
        print(f"slice {s:3d}: {e0.elapsed_time(e1) / 3:.1f} ms for 104 images", flush=True)
