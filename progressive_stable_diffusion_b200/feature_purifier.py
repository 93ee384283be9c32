"""Feature Purifier on B200: same module, parameters and state-dict keys as
``/root/reference/src/models/feature_purifier.py`` (:29-95), executed with the fused fp32 kernels of ``libdadd_b200``.

    img_n, aoe_n = LN(img), LN(aoe)                       -> dadd_layernorm_fwd
    q | k, v     = in_proj(img_n) | in_proj(aoe_n)         (cuBLAS, packed nn.MultiheadAttention weights)
    disease      = out_proj(MHA core)                      -> dadd_purifier_attn_fwd (16x16 per head, smem only)
    logits       = W2 gelu(W1 [disease | img_n])           (cuBLAS)
    out          = LN(img - sigmoid(logits) * disease)     -> dadd_purifier_gate_ln_fwd (gate + subtract + LN fused)

It runs once per sampling call on (B,16,768) tokens (inference_pipeline_ip.py:288-289), in fp32.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, wcache


class FeaturePurifier(nn.Module):
    def __init__(self, dim: int = 768, num_heads: int = 8, ff_mult: int = 2) -> None:
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.norm_img = nn.LayerNorm(dim)
        self.norm_aoe = nn.LayerNorm(dim)
        self.cross_attn = nn.MultiheadAttention(embed_dim=dim, num_heads=num_heads, batch_first=True)
        self.gate = nn.Sequential(nn.Linear(dim * 2, dim * ff_mult), nn.GELU(), nn.Linear(dim * ff_mult, dim), nn.Sigmoid())
        self.norm_out = nn.LayerNorm(dim)

    def _f32(self, owner, tag, p):
        return wcache.cast(owner, tag, p, torch.float32)

    def forward(self, image_embeds: torch.Tensor, source_aoe: torch.Tensor) -> torch.Tensor:
        out_dtype = image_embeds.dtype
        img = image_embeds.to(torch.float32).contiguous()
        aoe = source_aoe.to(torch.float32).contiguous()
        d = self.dim
        img_n = ops.layer_norm(img, self._f32(self.norm_img, "w", self.norm_img.weight), self._f32(self.norm_img, "b", self.norm_img.bias), self.norm_img.eps)
        aoe_n = ops.layer_norm(aoe, self._f32(self.norm_aoe, "w", self.norm_aoe.weight), self._f32(self.norm_aoe, "b", self.norm_aoe.bias), self.norm_aoe.eps)
        wi = self._f32(self.cross_attn, "wi", self.cross_attn.in_proj_weight)
        bi = self._f32(self.cross_attn, "bi", self.cross_attn.in_proj_bias)
        q = F.linear(img_n, wi[:d], bi[:d])
        kv = F.linear(aoe_n, wi[d:], bi[d:])                      # K and V in one GEMM
        k, v = kv[..., :d].contiguous(), kv[..., d:].contiguous()
        core = ops.purifier_attention(q, k, v, self.num_heads)
        op = self.cross_attn.out_proj
        disease = F.linear(core, self._f32(op, "w", op.weight), self._f32(op, "b", op.bias))
        g0, g2 = self.gate[0], self.gate[2]
        hid = F.gelu(F.linear(torch.cat([disease, img_n], dim=-1), self._f32(g0, "w", g0.weight), self._f32(g0, "b", g0.bias)))
        logits = F.linear(hid, self._f32(g2, "w", g2.weight), self._f32(g2, "b", g2.bias))
        out = ops.purifier_gate_ln(img, logits, disease.contiguous(), self._f32(self.norm_out, "w", self.norm_out.weight),
                                   self._f32(self.norm_out, "b", self.norm_out.bias), self.norm_out.eps)
        return out.to(out_dtype)
