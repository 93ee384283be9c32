"""Experiment (not product): how far can reduced precision drift over 50 DDIM steps?  Runs the ORACLE graph on the GPU in
fp32 (TF32 off) and under torch.autocast(bf16 / fp16), same weights and noise, and prints eps error and final PSNR."""
import math, sys, torch
sys.path.insert(0, ".")
from oracle import weights, unet as ou, sampler, conditioning
import oracle.unet, oracle.conditioning
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
# make the oracle's CPU-side constant creation land on the GPU
_ar = torch.arange
state = {k: v.to(dev) for k, v in weights.make_module_state(seed=0).items()}
torch.set_default_device(dev)
uw, aw, pw, vw = (weights.sub_state(state, p) for p in ("unet.unet.", "ordinal_embedder.", "feature_purifier.", "vae.vae."))
g = torch.Generator(device="cpu").manual_seed(3)
noise = torch.randn(1, 4, 32, 32, generator=g, device="cpu").to(dev).repeat(2, 1, 1, 1)
img = torch.randn(1, 16, 768, generator=g, device="cpu").to(dev).expand(2, -1, -1).contiguous()
tgt, src = torch.tensor([0.75, 3.0]), torch.ones(2)
_, ac = sampler.build_noise_schedule(); ts = sampler.ddim_timesteps()
ac = ac.to(dev)

def run(ctx):
    x = noise.clone()
    cond = conditioning.prepare_conditioning(aw, pw, tgt, src, img)
    cfg = ou.CrossCfg(True, 3.0)
    eps_all = []
    for i in range(50):
        t = int(ts[i])
        with ctx():
            eps = ou.unet_forward(uw, x, torch.full((2,), t), cond, cfg).float()
        eps_all.append(eps)
        x = sampler.ddim_update(x, eps, ac, t, None if i == 49 else int(ts[i + 1]))
    return x, eps_all

import contextlib
with torch.no_grad():
    ref, eref = run(contextlib.nullcontext)
    imr = ou.latents_to_images(vw, ref)
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        x, e = run(lambda: torch.autocast("cuda", dtype=dt))
        im = ou.latents_to_images(vw, x)
        mse = ((im - imr) ** 2).mean().item()
        e0 = ((e[0] - eref[0]).abs().max() / eref[0].abs().max()).item()
        r0 = ((e[0] - eref[0]).pow(2).mean().sqrt() / eref[0].pow(2).mean().sqrt()).item()
        print(f"autocast {name}: step0 eps max-rel {e0:.4g} rms-rel {r0:.4g}; final latent max-rel "
              f"{((x - ref).abs().max() / ref.abs().max()).item():.4g}; PSNR {10 * math.log10(1 / mse):.1f} dB")

# ---- variants: bf16 (or fp16) GEMM operands only, everything else fp32 -------------------------------------------
import torch.nn.functional as F
_conv2d, _linear, _sdpa, _matmul = F.conv2d, F.linear, F.scaled_dot_product_attention, torch.matmul

def patch(dt, round_out):
    def r(t):
        return t.to(dt).float()
    def conv2d(x, w, b=None, **kw):
        y = _conv2d(r(x), r(w), None if b is None else b, **kw)
        return r(y) if round_out else y
    def linear(x, w, b=None):
        y = _linear(r(x), r(w), b)
        return r(y) if round_out else y
    def sdpa(q, k, v):
        p = torch.softmax(_matmul(r(q), r(k).transpose(-1, -2)) / math.sqrt(q.shape[-1]), -1)
        y = _matmul(r(p), r(v))
        return r(y) if round_out else y
    def matmul(a, b):
        y = _matmul(r(a), r(b))
        return y
    F.conv2d, F.linear, F.scaled_dot_product_attention, torch.matmul = conv2d, linear, sdpa, matmul

def unpatch():
    F.conv2d, F.linear, F.scaled_dot_product_attention, torch.matmul = _conv2d, _linear, _sdpa, _matmul

with torch.no_grad():
    for name, dt, ro in (("bf16 operands, 16-bit GEMM outputs, fp32 elsewhere", torch.bfloat16, True),
                         ("bf16 operands, fp32 GEMM outputs, fp32 elsewhere", torch.bfloat16, False),
                         ("fp16 operands, 16-bit GEMM outputs, fp32 elsewhere", torch.float16, True)):
        patch(dt, ro)
        try:
            x, e = run(contextlib.nullcontext)
        finally:
            unpatch()
        im = ou.latents_to_images(vw, x)
        mse = ((im - imr) ** 2).mean().item()
        r0 = ((e[0] - eref[0]).pow(2).mean().sqrt() / eref[0].pow(2).mean().sqrt()).item()
        print(f"{name}: step0 eps rms-rel {r0:.4g}; PSNR {10 * math.log10(1 / mse):.1f} dB")
