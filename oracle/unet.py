"""Oracle restatement of the SD-1.x ``UNet2DConditionModel`` graph + AutoencoderKL decoder (TEST INFRASTRUCTURE).

PARITY UNPINNED: the arithmetic lives in the un-vendored dependency ``diffusers`` (>=0.31, pyproject.toml:27,
model ``CompVis/stable-diffusion-v1-4``); it is absent from /root/reference and from this image and the reference
has no test at that boundary.  This file restates the published architecture and is anchored
on the reference's call sites: ``src/models/unet/unet.py:96-146`` (wrapper), ``attention_processor_routing_gates.py``
(attn2 processors installed by name) and ``src/models/vae/vae.py:90-112``.

fp32, CPU, functional over a flat state dict with diffusers' key names (oracle/weights.py).
"""

from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import processors as P
from .weights import BLOCK_OUT, sub_state

W = Dict[str, torch.Tensor]


def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """diffusers ``Timesteps(320, flip_sin_to_cos=True, downscale_freq_shift=0)``: [cos | sin], fp32 (A.2 step 1)."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    args = t.to(torch.float32)[:, None] * freqs[None, :]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x, w, name, eps, silu):
    y = F.group_norm(x, 32, w[name + ".weight"], w[name + ".bias"], eps)
    return F.silu(y) if silu else y


def _conv(x, w, name, stride=1, padding=1):
    return F.conv2d(x, w[name + ".weight"], w[name + ".bias"], stride=stride, padding=padding)


def resnet_block(w: W, p: str, x: torch.Tensor, emb: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    """``ResnetBlock2D`` (A.3)."""
    h = _conv(_gn(x, w, p + ".norm1", eps, True), w, p + ".conv1")
    if emb is not None:
        h = h + F.linear(F.silu(emb), w[p + ".time_emb_proj.weight"], w[p + ".time_emb_proj.bias"])[:, :, None, None]
    h = _conv(_gn(h, w, p + ".norm2", eps, True), w, p + ".conv2")
    if (p + ".conv_shortcut.weight") in w:
        x = _conv(x, w, p + ".conv_shortcut", padding=0)
    return x + h


class CrossCfg:
    """Which attn2 processor is installed (mirrors DiffusionModuleWithIP._setup_attention_processors,
    diffusion_module_ip.py:203-233) and its mutable ``delta_scale`` (inference_pipeline_ip.py:311-318)."""

    def __init__(self, use_routing_gates: bool = True, delta_scale: float = 0.0, use_frequency_strategy: bool = True,
                 use_sdpa: bool = True) -> None:
        self.use_routing_gates = use_routing_gates
        self.delta_scale = delta_scale
        self.use_frequency_strategy = use_frequency_strategy
        self.use_sdpa = use_sdpa


def transformer_block(w: W, p: str, x: torch.Tensor, ehs: torch.Tensor, cfg: CrossCfg) -> torch.Tensor:
    """``Transformer2DModel`` depth 1 with conv projections (A.4) + ``BasicTransformerBlock`` (A.5)."""
    b, c, hh, ww = x.shape
    res = x
    h = _gn(x, w, p + ".norm", 1e-6, False)
    h = _conv(h, w, p + ".proj_in", padding=0)
    h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, c)
    t = p + ".transformer_blocks.0"
    n1 = F.layer_norm(h, (c,), w[t + ".norm1.weight"], w[t + ".norm1.bias"], 1e-5)
    h = h + P.self_attention(sub_state(w, t + ".attn1."), n1, use_sdpa=cfg.use_sdpa)
    n2 = F.layer_norm(h, (c,), w[t + ".norm2.weight"], w[t + ".norm2.bias"], 1e-5)
    w2 = sub_state(w, t + ".attn2.")
    if cfg.use_routing_gates:
        h = h + P.split_injection_attention(w2, n2, ehs, cfg.delta_scale)
    else:
        mode = P.frequency_mode_of(p) if cfg.use_frequency_strategy else "both"
        h = h + P.ordinal_ip_attention(w2, n2, ehs, mode)
    n3 = F.layer_norm(h, (c,), w[t + ".norm3.weight"], w[t + ".norm3.bias"], 1e-5)
    g = F.linear(n3, w[t + ".ff.net.0.proj.weight"], w[t + ".ff.net.0.proj.bias"])
    a, gate = g.chunk(2, dim=-1)                                  # GEGLU: value first, gate second
    h = h + F.linear(a * F.gelu(gate), w[t + ".ff.net.2.weight"], w[t + ".ff.net.2.bias"])
    h = h.reshape(b, hh, ww, c).permute(0, 3, 1, 2).contiguous()
    return _conv(h, w, p + ".proj_out", padding=0) + res


def unet_forward(w: W, sample: torch.Tensor, timesteps: torch.Tensor, cond_embed: torch.Tensor,
                 cfg: Optional[CrossCfg] = None) -> torch.Tensor:
    """``OrdinalUNet.forward`` (src/models/unet/unet.py:96-146) + the diffusers graph (A.2).  Returns eps."""
    cfg = cfg or CrossCfg()
    if cond_embed.ndim == 2:
        ehs = cond_embed.unsqueeze(1)
    elif cond_embed.ndim == 3:
        ehs = cond_embed
    else:
        raise ValueError(f"cond_embed must have shape (B, D) or (B, seq_len, D), got {cond_embed.shape}")
    if timesteps.ndim == 0:
        timesteps = timesteps[None]
    elif timesteps.ndim > 1:
        timesteps = timesteps.view(-1)
    timesteps = timesteps.expand(sample.shape[0])
    temb = timestep_embedding(timesteps)
    emb = F.linear(F.silu(F.linear(temb, w["time_embedding.linear_1.weight"], w["time_embedding.linear_1.bias"])),
                   w["time_embedding.linear_2.weight"], w["time_embedding.linear_2.bias"])
    h = _conv(sample, w, "conv_in")
    skips = [h]
    for i in range(4):
        for j in range(2):
            h = resnet_block(w, f"down_blocks.{i}.resnets.{j}", h, emb)
            if i < 3:
                h = transformer_block(w, f"down_blocks.{i}.attentions.{j}", h, ehs, cfg)
            skips.append(h)
        if i < 3:
            h = _conv(h, w, f"down_blocks.{i}.downsamplers.0.conv", stride=2, padding=1)
            skips.append(h)
    h = resnet_block(w, "mid_block.resnets.0", h, emb)
    h = transformer_block(w, "mid_block.attentions.0", h, ehs, cfg)
    h = resnet_block(w, "mid_block.resnets.1", h, emb)
    for i in range(4):
        for j in range(3):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet_block(w, f"up_blocks.{i}.resnets.{j}", h, emb)
            if i > 0:
                h = transformer_block(w, f"up_blocks.{i}.attentions.{j}", h, ehs, cfg)
        if i < 3:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, w, f"up_blocks.{i}.upsamplers.0.conv")
    h = _gn(h, w, "conv_norm_out", 1e-5, True)
    return _conv(h, w, "conv_out")


def vae_decode(w: W, z: torch.Tensor) -> torch.Tensor:
    """``AutoencoderKL.decode(z).sample`` (called at src/models/vae/vae.py:112): post_quant_conv -> Decoder."""
    h = _conv(z, w, "post_quant_conv", padding=0)
    h = _conv(h, w, "decoder.conv_in")
    h = resnet_block(w, "decoder.mid_block.resnets.0", h, None, eps=1e-6)
    a = "decoder.mid_block.attentions.0"
    b, c, hh, ww = h.shape
    x = F.group_norm(h.view(b, c, hh * ww), 32, w[a + ".group_norm.weight"], w[a + ".group_norm.bias"], 1e-6)
    x = x.transpose(1, 2)
    q = F.linear(x, w[a + ".to_q.weight"], w[a + ".to_q.bias"])
    k = F.linear(x, w[a + ".to_k.weight"], w[a + ".to_k.bias"])
    v = F.linear(x, w[a + ".to_v.weight"], w[a + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]      # one head of d=512
    o = F.linear(o, w[a + ".to_out.0.weight"], w[a + ".to_out.0.bias"])
    h = h + o.transpose(1, 2).reshape(b, c, hh, ww)                                   # residual_connection=True
    h = resnet_block(w, "decoder.mid_block.resnets.1", h, None, eps=1e-6)
    for i in range(4):
        for j in range(3):
            h = resnet_block(w, f"decoder.up_blocks.{i}.resnets.{j}", h, None, eps=1e-6)
        if i < 3:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, w, f"decoder.up_blocks.{i}.upsamplers.0.conv")
    h = _gn(h, w, "decoder.conv_norm_out", 1e-6, True)
    return _conv(h, w, "decoder.conv_out")


def _vae_mid_attention(w: W, a: str, h: torch.Tensor) -> torch.Tensor:
    b, c, hh, ww = h.shape
    x = F.group_norm(h.view(b, c, hh * ww), 32, w[a + ".group_norm.weight"], w[a + ".group_norm.bias"], 1e-6).transpose(1, 2)
    q = F.linear(x, w[a + ".to_q.weight"], w[a + ".to_q.bias"])
    k = F.linear(x, w[a + ".to_k.weight"], w[a + ".to_k.bias"])
    v = F.linear(x, w[a + ".to_v.weight"], w[a + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, w[a + ".to_out.0.weight"], w[a + ".to_out.0.bias"])
    return h + o.transpose(1, 2).reshape(b, c, hh, ww)


def vae_encode_moments(w: W, images: torch.Tensor) -> torch.Tensor:
    """``AutoencoderKL.encode(x)`` up to the posterior's moments [mean | logvar] (B, 8, H/8, W/8): Encoder -> quant_conv
    (diffusers graph; called at src/models/diffusion_module_ip.py:419).  Downsample2D(padding=0) pads right / bottom by one."""
    h = _conv(images, w, "encoder.conv_in")
    for i in range(4):
        for j in range(2):
            h = resnet_block(w, f"encoder.down_blocks.{i}.resnets.{j}", h, None, eps=1e-6)
        if i < 3:
            h = _conv(F.pad(h, (0, 1, 0, 1)), w, f"encoder.down_blocks.{i}.downsamplers.0.conv", stride=2, padding=0)
    h = resnet_block(w, "encoder.mid_block.resnets.0", h, None, eps=1e-6)
    h = _vae_mid_attention(w, "encoder.mid_block.attentions.0", h)
    h = resnet_block(w, "encoder.mid_block.resnets.1", h, None, eps=1e-6)
    h = _conv(_gn(h, w, "encoder.conv_norm_out", 1e-6, True), w, "encoder.conv_out")
    return _conv(h, w, "quant_conv", padding=0)


def latents_to_images(vae_w: W, latents: torch.Tensor, latent_scale: float = 0.18215) -> torch.Tensor:
    """``_latents_to_images`` (inference_pipeline_ip.py:473-486)."""
    img = vae_decode(vae_w, latents / latent_scale).clamp(-1.0, 1.0)
    return ((img + 1.0) / 2.0).clamp(0.0, 1.0)


def attention_shapes(latent_hw: int = 32):
    """(site prefix, C, N, d) table of SURVEY.md Appendix B.3."""
    from .weights import attention_sites
    res = {320: latent_hw, 640: latent_hw // 2}
    out = []
    for p, c in attention_sites():
        if c == 1280:
            r = latent_hw // 8 if p.startswith("mid_block") else latent_hw // 4
        else:
            r = res[c]
        out.append((p, c, r * r, c // 8))
    return out


assert BLOCK_OUT == (320, 640, 1280, 1280)
