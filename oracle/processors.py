"""Oracle restatement of the two DADD cross-attention processors and of attn1 (TEST INFRASTRUCTURE).

Pinned against the verbatim reference classes by tests/golden/make_golden.py ->
tests/golden/processor_*.npz (see tests/test_oracle_golden.py).
All functions are pure: weights come in as a dict ``w`` holding the keys of one diffusers
``Attention`` module (``to_q.weight`` ... ``to_out.0.bias``) plus, for the routing processor,
``processor.{to_k_dis.weight,to_v_dis.weight,anat_gate,dis_gate}``.
"""

from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

W = Dict[str, torch.Tensor]


def _heads(x: torch.Tensor, heads: int) -> torch.Tensor:
    b, n, c = x.shape
    return x.view(b, n, heads, c // heads).transpose(1, 2)


def _softmax_attend(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    # explicit matmul / sqrt(d) -> softmax -> matmul, attention_processor_routing_gates.py:148-158
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    return torch.matmul(F.softmax(s, dim=-1), v)


def split_injection_attention(
    w: W,
    hidden_states: torch.Tensor,
    encoder_hidden_states: torch.Tensor,
    delta_scale: float,
    heads: int = 8,
    n_aoe: int = 16,
    n_img: int = 16,
    n_delta: int = 16,
) -> torch.Tensor:
    """``SplitInjectionAttentionProcessor.__call__`` (attention_processor_routing_gates.py:84-196)
    for the SD-1.x case (3-D input, no mask, spatial_norm/group_norm None, no residual, rescale 1)."""
    b = hidden_states.shape[0]
    q = _heads(F.linear(hidden_states, w["to_q.weight"]), heads)
    dis = encoder_hidden_states[:, :n_aoe, :]                      # :129
    anat = encoder_hidden_states[:, n_aoe:n_aoe + n_img, :]        # :130
    delta = encoder_hidden_states[:, -n_delta:, :]                 # :131
    k_a = _heads(F.linear(anat, w["to_k.weight"]), heads)          # anatomy uses the pretrained text K/V, :133-134
    v_a = _heads(F.linear(anat, w["to_v.weight"]), heads)
    k_d = _heads(F.linear(dis, w["processor.to_k_dis.weight"]), heads)   # :136-137
    v_d = _heads(F.linear(dis, w["processor.to_v_dis.weight"]), heads)
    z_a = _softmax_attend(q, k_a, v_a)
    z_d = _softmax_attend(q, k_d, v_d)
    g_a = w["processor.anat_gate"]
    g_d = w["processor.dis_gate"]
    if delta_scale != 0.0:                                         # :160 (I2: pathway skipped when 0)
        k_x = _heads(F.linear(delta, w["processor.to_k_dis.weight"]), heads)
        v_x = _heads(F.linear(delta, w["processor.to_v_dis.weight"]), heads)
        z_x = _softmax_attend(q, k_x, v_x)
        z = g_a * z_a + g_d * z_d + delta_scale * z_x              # :172-176
    else:
        z = g_a * z_a + g_d * z_d                                  # :178
    z = z.transpose(1, 2).reshape(b, -1, q.shape[1] * q.shape[-1])
    return F.linear(z, w["to_out.0.weight"], w["to_out.0.bias"])  # to_out[1] is Dropout(0)


def ordinal_ip_attention(
    w: W,
    hidden_states: torch.Tensor,
    encoder_hidden_states: torch.Tensor,
    frequency_mode: str = "both",
    heads: int = 8,
    n_aoe: int = 16,
    n_img: int = 16,
) -> torch.Tensor:
    """``OrdinalIPAttnProcessor2_0.__call__`` (attention_processor_base.py:39-138): one softmax over
    the concatenated [AOE | image] tokens; for frequency_mode != "both" the probabilities are multiplied by
    an all-ones scale vector (scale_aoe = scale_ip = 1, :29-37) and renormalised (:103-116)."""
    b = hidden_states.shape[0]
    q = _heads(F.linear(hidden_states, w["to_q.weight"]), heads)
    k = _heads(F.linear(encoder_hidden_states, w["to_k.weight"]), heads)
    v = _heads(F.linear(encoder_hidden_states, w["to_v.weight"]), heads)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(q.shape[-1])
    p = F.softmax(s, dim=-1)
    if frequency_mode != "both":
        scale = torch.ones((1, 1, 1, p.shape[-1]), dtype=p.dtype)
        if p.shape[-1] >= n_aoe + n_img:
            scale[..., :n_aoe] *= 1
            scale[..., -n_img:] *= 1
        p = p * scale
        p = p / p.sum(dim=-1, keepdim=True)
    z = torch.matmul(p, v).transpose(1, 2).reshape(b, -1, q.shape[1] * q.shape[-1])
    return F.linear(z, w["to_out.0.weight"], w["to_out.0.bias"])


def frequency_mode_of(block_name: str) -> str:
    """``get_frequency_mode_for_block`` (attention_processor_base.py:141-167)."""
    if "mid_block" in block_name:
        return "aoe_dominant"
    if "down_blocks" in block_name:
        idx = int(block_name.split("down_blocks.")[1].split(".")[0])
        return "image_dominant" if idx <= 1 else "aoe_dominant"
    if "up_blocks" in block_name:
        idx = int(block_name.split("up_blocks.")[1].split(".")[0])
        return "aoe_dominant" if idx <= 1 else "image_dominant"
    return "both"


def self_attention(w: W, hidden_states: torch.Tensor, heads: int = 8, use_sdpa: bool = True) -> torch.Tensor:
    """attn1 = diffusers ``AttnProcessor2_0`` (un-vendored): bias-free q/k/v,
    ``F.scaled_dot_product_attention`` (scale d^-1/2, no mask), merge heads, to_out[0] (+bias)."""
    b = hidden_states.shape[0]
    q = _heads(F.linear(hidden_states, w["to_q.weight"]), heads)
    k = _heads(F.linear(hidden_states, w["to_k.weight"]), heads)
    v = _heads(F.linear(hidden_states, w["to_v.weight"]), heads)
    z = F.scaled_dot_product_attention(q, k, v) if use_sdpa else _softmax_attend(q, k, v)
    z = z.transpose(1, 2).reshape(b, -1, q.shape[1] * q.shape[-1])
    return F.linear(z, w["to_out.0.weight"], w["to_out.0.bias"])
