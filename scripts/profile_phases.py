"""Where one progression batch (8 patients x 13 levels x 50 steps, the bench's `value` leg) spends its time: device time (CUDA
events) and host wall time per phase.  python scripts/profile_phases.py [--dtype fp16]"""
import argparse, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_stable_diffusion_b200 as P
from progressive_stable_diffusion_b200 import inference_pipeline_ip as ip

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--patients", type=int, default=8)
args = ap.parse_args()
dev = torch.device("cuda", 0)
P.set_compute_dtype(torch.bfloat16 if args.dtype == "bf16" else torch.float16)
torch.manual_seed(0)
module = P.DiffusionModuleWithIP(P.default_config(), build_image_encoder=True).to(dev).eval()
L, p = 13, args.patients
g = torch.Generator().manual_seed(1)
tokens = torch.randn(p, 16, 768, generator=g).to(dev).repeat_interleave(L, 0)
source = torch.zeros(p * L, device=dev)
target = ip._build_labels(L, 0.0, 3.0, dev).repeat(p)
noise = torch.randn(p, 4, 32, 32, generator=g).to(dev).repeat_interleave(L, 0)
pixels = torch.randn(p, 3, 224, 224, generator=g).to(dev)


class Phase:
    def __init__(self): self.rows = []
    def run(self, name, fn):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        out = fn()
        e1.record(); t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        self.rows.append((name, e0.elapsed_time(e1), t_issue * 1e3, (time.perf_counter() - t0) * 1e3))
        return out


with torch.no_grad():
    for _ in range(2):
        lat = ip._sample(module, target, source, tokens, noise, 50, dev, 0.0, 1.0, None, 3.0, 1.0, False, True)
        ip._latents_to_images(module, lat)
    ph = Phase()
    ph.run("front end: CLIP ViT-L/14 + resampler (8 patients)", lambda: module._get_image_embeds(pixels))
    cond = ph.run("_prepare_conditioning (purifier, AOE, delta)", lambda: ip._prepare_conditioning(module, target, source, tokens, image_scale=1.0, leace=None))
    ph.run("_set_delta_scale_on_processors", lambda: ip._set_delta_scale_on_processors(module, 3.0))
    eng = ip._engine_for(module, p * L, 50, 0.0, False, 1.0, cond.shape[1], dev, 3.0, True)
    ph.run("engine: copy inputs", lambda: (eng.x.copy_(noise), eng.ehs.copy_(cond)))
    ph.run("engine: _revalidate_weights", eng._revalidate_weights)
    ph.run("engine: _refresh_kv (K/V projections of 16 sites, gate vectors)", eng._refresh_kv)
    ph.run("engine: state.zero_", lambda: eng.state.zero_())
    ph.run("engine: 50 graph replays", lambda: [eng.graph.replay() for _ in range(50)])
    lat = ph.run("engine: x.clone", lambda: eng.x.clone())
    ph.run("whole eng.run()", lambda: eng.run(noise, cond, None, None))
    ph.run("whole _sample()", lambda: ip._sample(module, target, source, tokens, noise, 50, dev, 0.0, 1.0, None, 3.0, 1.0, False, True))
    img = ph.run("VAE decode + image post", lambda: ip._latents_to_images(module, lat))
    host = torch.empty(img.shape, dtype=img.dtype).pin_memory()
    ph.run("D2H of the images (82 MB, pinned)", lambda: host.copy_(img, non_blocking=True))
print(f"{'phase':70s} {'device ms':>10s} {'host issue ms':>14s} {'wall ms':>9s}")
for name, d, i, w in ph.rows:
    print(f"{name:70s} {d:10.2f} {i:14.2f} {w:9.2f}")
