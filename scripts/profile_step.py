"""One eager denoising step (UNet + fused DDIM) of the bench workload between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python scripts/profile_step.py
Without ncu it prints the CUDA-event time of the step (eager and graph replay)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import progressive_stable_diffusion_b200 as P  # noqa: E402
from progressive_stable_diffusion_b200.inference_pipeline_ip import _build_labels, _sample  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--patients", type=int, default=2)
ap.add_argument("--dtype", default="bf16")
ap.add_argument("--steps", type=int, default=1)
args = ap.parse_args()
P.set_compute_dtype(torch.bfloat16 if args.dtype == "bf16" else torch.float16)
dev = torch.device("cuda", 0)
torch.manual_seed(0)
module = P.DiffusionModuleWithIP(P.default_config(), build_vae=False).to(dev).eval()
b = args.patients * 13
g = torch.Generator().manual_seed(1)
tokens = torch.randn(args.patients, 16, 768, generator=g).to(dev).repeat_interleave(13, 0)
noise = torch.randn(args.patients, 4, 32, 32, generator=g).to(dev).repeat_interleave(13, 0)
src = torch.zeros(b, device=dev)
tgt = _build_labels(13, 0.0, 3.0, dev).repeat(args.patients)
with torch.no_grad():
    _sample(module, tgt, src, tokens, noise, 50, dev, 0.0, 1.0, None, 3.0, 1.0, False, True)
    eng = next(iter(module.__dict__["_b200_engines"].values()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.state.zero_()
    e0.record()
    for _ in range(10):
        eng.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay: {e0.elapsed_time(e1) / 10:.3f} ms per denoising step at B={b} ({eng.launches_per_step} dadd launches)")
    eng.state.zero_()
    eng._step()
    torch.cuda.synchronize()
    eng.state.zero_()
    torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        eng._step()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print(f"eager: {e0.elapsed_time(e1) / args.steps:.3f} ms per denoising step")
