"""Oracle restatement of the once-per-call conditioning front end (TEST INFRASTRUCTURE):
AdditiveOrdinalEmbedder, FeaturePurifier and the pipeline's ``_prepare_conditioning``.

Pinned against the verbatim reference modules: tests/golden/{aoe,purifier}_*.npz.
"""

from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- AOE
def aoe_class_table(w: W) -> torch.Tensor:
    """E[k] = base + sum_{j<k} deltas[j]  (ordinal_embedder.py:107-127)."""
    cum = torch.cumsum(w["deltas"], dim=0)
    off = torch.cat([torch.zeros(1, cum.shape[1], dtype=cum.dtype), cum], dim=0)
    return w["base"].unsqueeze(0) + off


def aoe_interp(w: W, labels: torch.Tensor) -> torch.Tensor:
    """clamp -> floor/ceil gather -> lerp (ordinal_embedder.py:155-171, 15-40)."""
    table = aoe_class_table(w)
    kmax = table.shape[0] - 1
    y = torch.clamp(labels.to(table.dtype), 0.0, float(kmax))
    lo = torch.floor(y)
    hi = torch.clamp(lo + 1, max=kmax)
    a = (y - lo).unsqueeze(-1)
    return F.embedding(lo.long(), table) * (1.0 - a) + F.embedding(hi.long(), table) * a


def aoe_project(w: W, emb: torch.Tensor, num_tokens: int = 16) -> torch.Tensor:
    """projector = Linear(D,2D) -> GELU(erf) -> Linear(2D, T*D); view (B,T,D) (ordinal_embedder.py:79-84,177-178).
    ``self.norm`` (:85) is never applied."""
    h = F.gelu(F.linear(emb, w["projector.0.weight"], w["projector.0.bias"]))
    out = F.linear(h, w["projector.2.weight"], w["projector.2.bias"])
    return out.view(-1, num_tokens, emb.shape[-1])


def aoe_forward(w: W, labels: torch.Tensor, num_tokens: int = 16) -> torch.Tensor:
    """``AdditiveOrdinalEmbedder.forward(labels, is_training=False)`` (ordinal_embedder.py:129-180)."""
    return aoe_project(w, aoe_interp(w, labels), num_tokens)


def aoe_negative(w: W, labels: torch.Tensor, num_tokens: int = 16) -> torch.Tensor:
    """``get_negative_embedding``: labels -> clamp(1 - y, 0, 1) (ordinal_embedder.py:182-221)."""
    return aoe_forward(w, torch.clamp(1.0 - labels, min=0.0, max=1.0), num_tokens)


def aoe_delta(w: W, source: torch.Tensor, target: torch.Tensor, num_tokens: int = 16) -> torch.Tensor:
    """``get_ordinal_delta_embedding`` = proj(E[target]) - proj(E[source]) (ordinal_embedder.py:246-294)."""
    return aoe_project(w, aoe_interp(w, target), num_tokens) - aoe_project(w, aoe_interp(w, source), num_tokens)


# ----------------------------------------------------------------------------- purifier
def purifier_forward(w: W, image_embeds: torch.Tensor, source_aoe: torch.Tensor, num_heads: int = 8) -> torch.Tensor:
    """``FeaturePurifier.forward`` (feature_purifier.py:64-95).  nn.MultiheadAttention(batch_first) restated:
    packed in_proj rows [0:D]=q, [D:2D]=k, [2D:3D]=v (+bias), scale 1/sqrt(D/heads), out_proj (+bias)."""
    d = image_embeds.shape[-1]
    img_n = F.layer_norm(image_embeds, (d,), w["norm_img.weight"], w["norm_img.bias"], 1e-5)
    aoe_n = F.layer_norm(source_aoe, (d,), w["norm_aoe.weight"], w["norm_aoe.bias"], 1e-5)
    wi, bi = w["cross_attn.in_proj_weight"], w["cross_attn.in_proj_bias"]
    q = F.linear(img_n, wi[:d], bi[:d])
    k = F.linear(aoe_n, wi[d:2 * d], bi[d:2 * d])
    v = F.linear(aoe_n, wi[2 * d:], bi[2 * d:])
    b, n, _ = q.shape
    hd = d // num_heads
    qh = q.view(b, n, num_heads, hd).transpose(1, 2)
    kh = k.view(b, -1, num_heads, hd).transpose(1, 2)
    vh = v.view(b, -1, num_heads, hd).transpose(1, 2)
    p = F.softmax(torch.matmul(qh, kh.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    z = torch.matmul(p, vh).transpose(1, 2).reshape(b, n, d)
    disease = F.linear(z, w["cross_attn.out_proj.weight"], w["cross_attn.out_proj.bias"])
    gate_in = torch.cat([disease, img_n], dim=-1)                               # :88
    h = F.gelu(F.linear(gate_in, w["gate.0.weight"], w["gate.0.bias"]))
    mask = torch.sigmoid(F.linear(h, w["gate.2.weight"], w["gate.2.bias"]))     # :89
    clean = image_embeds - mask * disease                                       # :92
    return F.layer_norm(clean, (d,), w["norm_out.weight"], w["norm_out.bias"], 1e-5)


# ----------------------------------------------------------------------------- pipeline assembly
def prepare_conditioning(
    aoe_w: W,
    purifier_w: Optional[W],
    target_labels: torch.Tensor,
    source_labels: torch.Tensor,
    image_embeds: torch.Tensor,
    use_routing_gates: bool = True,
    image_scale: float = 1.0,
    zero_aoe: bool = False,
) -> torch.Tensor:
    """``_prepare_conditioning`` (inference_pipeline_ip.py:232-308 / evaluation_pipeline.py:406-461) from the point
    where ``image_embeds = module._get_image_embeds(...)`` (B,16,768) is available (CLIP + resampler are off-path)."""
    t_aoe = aoe_negative(aoe_w, target_labels) if zero_aoe else aoe_forward(aoe_w, target_labels)
    s_aoe = aoe_forward(aoe_w, source_labels)
    if purifier_w is not None:
        image_embeds = purifier_forward(purifier_w, image_embeds, s_aoe)
    if image_scale != 1.0:
        image_embeds = image_embeds * image_scale
    if use_routing_gates:
        delta = aoe_delta(aoe_w, source_labels, target_labels)
        return torch.cat([s_aoe, image_embeds, delta], dim=1)     # [Source_AOE | E_clean | Delta_AOE]
    return torch.cat([t_aoe, image_embeds], dim=1)                # [AOE | Image]
