"""Self-attention processor for ``attn1`` and the host-side pieces shared by the DADD cross-attention processors.

``AttnProcessor2_0`` has the name and call protocol of diffusers' class that the reference installs on every ``attn1``
(``/root/reference/src/models/attention_processor_routing_gates.py:284-286``, ``attention_processor_base.py:196-197``):
``processor(attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None) -> Tensor``.
The softmax(QK^T/sqrt d)V core runs in ``dadd_self_attn_fwd``; Q, K and V come from ONE fused (3C x C) projection whose
output the kernel reads in place through row strides.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import ops, wcache

# Element type of activations / GEMM operands on the device path.  The default is fp16 - the reference's own mixed-precision
# dtype (training "16-mixed", evaluation_pipeline.py:943 autocast float16): same tensor-core rate and bytes as bf16, 3 more
# mantissa bits, and the only 16-bit operand type that meets ALL parity gates of BASELINE.md section 4 (eps <= 2e-2 AND 50-step
# PSNR >= 40 dB: 42-45 dB; any bf16-operand implementation, stock PyTorch autocast included, ends at 26-30 dB on the
# non-contractive random-init weights, profiles/r01_precision_experiment.txt).  bf16 stays selectable.
DEFAULT_COMPUTE_DTYPE = torch.float16


class _Compute:
    dtype = DEFAULT_COMPUTE_DTYPE


wcache.set_namespace(str(DEFAULT_COMPUTE_DTYPE))


def set_compute_dtype(dtype: torch.dtype) -> None:
    if dtype not in (torch.bfloat16, torch.float16):
        raise ValueError("compute dtype must be torch.bfloat16 or torch.float16")
    wcache.set_namespace(str(dtype))   # derived weights (fused QKV, cached K/V ...) are stored per compute dtype
    _Compute.dtype = dtype


def compute_dtype() -> torch.dtype:
    return _Compute.dtype



def _as_tokens(attn, hidden_states: torch.Tensor, temb):
    """The reference processors' prologue (routing_gates.py:94-104): optional spatial_norm, 4-D -> (B, HW, C)."""
    if getattr(attn, "spatial_norm", None) is not None:
        hidden_states = attn.spatial_norm(hidden_states, temb)
    shape4 = None
    if hidden_states.ndim == 4:
        shape4 = hidden_states.shape
        b, c, h, w = shape4
        hidden_states = hidden_states.reshape(b, c, h * w).transpose(1, 2)
    if getattr(attn, "group_norm", None) is not None:
        hidden_states = attn.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
    return hidden_states, shape4


def _finish(attn, out: torch.Tensor, residual: torch.Tensor, shape4, out_dtype: torch.dtype) -> torch.Tensor:
    """to_out[0] (+bias), to_out[1] = Dropout(0), optional reshape / residual / rescale (routing_gates.py:183-196)."""
    w = wcache.cast(attn.to_out[0], "w", attn.to_out[0].weight, compute_dtype())
    bias = attn.to_out[0].bias
    b32 = None if bias is None else wcache.cast(attn.to_out[0], "b32", bias, torch.float32)
    b_lp = None if bias is None else wcache.cast(attn.to_out[0], "b", bias, compute_dtype())
    out = ops.linear(out, w, b32, bias_lp=b_lp)
    if shape4 is not None:
        b, c, h, w_ = shape4
        out = out.transpose(-1, -2).reshape(b, c, h, w_)
    if getattr(attn, "residual_connection", False):
        out = out + residual.to(out.dtype)
    factor = getattr(attn, "rescale_output_factor", 1.0)
    if factor != 1.0:
        out = out / factor
    return out.to(out_dtype)


def _reject_mask(attention_mask) -> None:
    if attention_mask is not None:
        raise NotImplementedError(
            "the B200 attention kernels take no attention_mask (it is always None on DADD's UNet path, SURVEY.md 8a)")


class AttnProcessor2_0:
    """Drop-in for diffusers' ``AttnProcessor2_0`` on self-attention sites (no parameters, like the original)."""

    def __call__(self, attn, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                 attention_mask: Optional[torch.Tensor] = None, temb: Optional[torch.Tensor] = None, *args, **kwargs):
        _reject_mask(attention_mask)
        if encoder_hidden_states is not None and encoder_hidden_states is not hidden_states:
            raise NotImplementedError("AttnProcessor2_0 (B200) implements self-attention only; cross-attention sites use "
                                      "SplitInjectionAttentionProcessor / OrdinalIPAttnProcessor2_0")
        residual = hidden_states
        out_dtype = hidden_states.dtype
        x, shape4 = _as_tokens(attn, hidden_states, temb)
        x = x.to(compute_dtype())
        if not x.is_contiguous():
            x = x.contiguous()
        wqkv = wcache.get(attn, "wqkv", (attn.to_q.weight, attn.to_k.weight, attn.to_v.weight),
                          lambda: torch.cat([attn.to_q.weight, attn.to_k.weight, attn.to_v.weight], 0)
                          .detach().to(compute_dtype()).contiguous())
        c = attn.to_q.weight.shape[0]
        bqkv = bqkv_lp = None
        if attn.to_q.bias is not None:
            bqkv = wcache.get(attn, "bqkv32", (attn.to_q.bias, attn.to_k.bias, attn.to_v.bias),
                              lambda: torch.cat([attn.to_q.bias, attn.to_k.bias, attn.to_v.bias], 0)
                              .detach().float().contiguous())
            bqkv_lp = wcache.get(attn, "bqkv", (attn.to_q.bias, attn.to_k.bias, attn.to_v.bias), lambda: bqkv.to(compute_dtype()))
        qkv = ops.linear(x, wqkv, bqkv, bias_lp=bqkv_lp)           # (B, N, 3C): one GEMM
        o = ops.self_attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], attn.heads)
        return _finish(attn, o, residual, shape4, out_dtype)
