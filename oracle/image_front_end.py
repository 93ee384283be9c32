"""CPU oracle of the conditioning front end (TEST INFRASTRUCTURE - only tests/, smoke() and bench.py's cpu_baseline leg may
import this): CLIP vision tower -> hidden states -> ImageProjectionPlus resampler -> (B, 16, 768) anatomy tokens.

Restates, in plain fp32 PyTorch on state dicts:
  * ``CLIPVisionModelWithProjection`` of the un-vendored dependency ``transformers`` (pyproject.toml: ``transformers``; the
    reference calls it at /root/reference/src/models/image_encoder.py:34-38 (construction), :63-68 (``image_embeds``) and
    :82-87 (``hidden_states[-1]``: the last encoder layer's output, BEFORE ``post_layernorm``); model
    ``openai/clip-vit-large-patch14``: hidden 1024, 24 layers, 16 heads, MLP 4096, 14x14 patches of a 224x224 image,
    ``quick_gelu``, LayerNorm eps 1e-5, projection 768 without bias).  Published algorithm: x = [class; patches] + position
    embeddings -> pre_layrnorm -> L x {x += out_proj(softmax(q k^T / sqrt d) v) on LN1(x); x += fc2(quick_gelu(fc1(LN2(x))))};
    image_embeds = visual_projection(post_layernorm(x[:, 0])).
    PINNED: ``transformers`` 5.5 is installed in the build container, so tests/golden/make_golden.py runs the real
    ``CLIPVisionModelWithProjection`` on seeded weights / pixels and stores its outputs (tests/golden/reference_modules.npz).
  * ``ImageProjectionPlus.forward`` (/root/reference/src/models/image_encoder.py:193-228) and ``ImageProjection.forward``
    (:119-133).  PINNED against the verbatim reference classes the same way.
"""

from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]

CLIP_L14 = dict(hidden=1024, inter=4096, layers=24, heads=16, image=224, patch=14, proj=768)
CLIP_TINY = dict(hidden=64, inter=256, layers=2, heads=4, image=28, patch=14, proj=32)     # CPU-sized fixture


def _ln(x: torch.Tensor, w: W, name: str, eps: float = 1e-5) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], eps)


def _lin(x: torch.Tensor, w: W, name: str) -> torch.Tensor:
    return F.linear(x, w[name + ".weight"], w.get(name + ".bias"))


def clip_hidden_states(w: W, pixels: torch.Tensor, heads: int, patch: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(last encoder layer output (B, 1 + n_patches, hidden), image_embeds (B, proj)); ``w`` has the keys of
    ``CLIPVisionModelWithProjection.state_dict()``."""
    p = "vision_model."
    x = F.conv2d(pixels, w[p + "embeddings.patch_embedding.weight"], None, stride=patch)          # (B, hidden, g, g)
    b, c = x.shape[:2]
    x = x.flatten(2).transpose(1, 2)                                                               # (B, g*g, hidden)
    x = torch.cat([w[p + "embeddings.class_embedding"].expand(b, 1, c), x], dim=1)
    x = x + w[p + "embeddings.position_embedding.weight"][None]
    x = _ln(x, w, p + "pre_layrnorm")
    d = c // heads
    layer = 0
    while f"{p}encoder.layers.{layer}.layer_norm1.weight" in w:
        q = f"{p}encoder.layers.{layer}."
        h = _ln(x, w, q + "layer_norm1")
        n = h.shape[1]
        qh, kh, vh = (_lin(h, w, q + f"self_attn.{t}_proj").view(b, n, heads, d).transpose(1, 2) for t in "qkv")
        a = torch.softmax(qh @ kh.transpose(-1, -2) * d ** -0.5, dim=-1) @ vh
        x = x + _lin(a.transpose(1, 2).reshape(b, n, c), w, q + "self_attn.out_proj")
        h = _lin(_ln(x, w, q + "layer_norm2"), w, q + "mlp.fc1")
        x = x + _lin(h * torch.sigmoid(1.702 * h), w, q + "mlp.fc2")                              # quick_gelu
        layer += 1
    embeds = F.linear(_ln(x[:, 0], w, p + "post_layernorm"), w["visual_projection.weight"])
    return x, embeds


def projection_plus(w: W, hidden: torch.Tensor, heads: int = 8) -> torch.Tensor:
    """ImageProjectionPlus.forward (image_encoder.py:193-228): Perceiver resampler, learnable queries attend to the patches."""
    b = hidden.shape[0]
    if "proj_in.weight" in w:
        hidden = _lin(hidden, w, "proj_in")
    lat = w["latents"].expand(b, -1, -1)
    dim = lat.shape[-1]
    d = dim // heads
    layer = 0
    while f"layers.{layer}.norm1.weight" in w:
        q = f"layers.{layer}."
        wi, bi = w[q + "cross_attn.in_proj_weight"], w[q + "cross_attn.in_proj_bias"]
        x = _ln(lat, w, q + "norm1")
        qh = F.linear(x, wi[:dim], bi[:dim]).view(b, -1, heads, d).transpose(1, 2)
        kh = F.linear(hidden, wi[dim:2 * dim], bi[dim:2 * dim]).view(b, -1, heads, d).transpose(1, 2)
        vh = F.linear(hidden, wi[2 * dim:], bi[2 * dim:]).view(b, -1, heads, d).transpose(1, 2)
        a = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(d), dim=-1) @ vh
        lat = lat + _lin(a.transpose(1, 2).reshape(b, -1, dim), w, q + "cross_attn.out_proj")
        x = _ln(lat, w, q + "norm2")
        lat = lat + _lin(F.gelu(_lin(x, w, q + "ff.0")), w, q + "ff.2")
        layer += 1
    return _ln(lat, w, "norm_out")


def projection_basic(w: W, image_embeds: torch.Tensor, num_tokens: int, dim: int) -> torch.Tensor:
    """ImageProjection.forward (image_encoder.py:119-133)."""
    return _ln(_lin(image_embeds, w, "projection").reshape(-1, num_tokens, dim), w, "norm")
