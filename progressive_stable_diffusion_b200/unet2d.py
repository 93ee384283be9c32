"""SD-1.x ``UNet2DConditionModel`` for B200, with the surface the reference uses from diffusers.

The reference gets this graph from the un-vendored ``diffusers`` package
(``/root/reference/src/models/unet/unet.py:70-75,140-146``); what it touches is: ``unet(sample=, timestep=,
encoder_hidden_states=).sample``, ``unet.config.{in_channels,out_channels,cross_attention_dim,block_out_channels}``,
``unet.attn_processors``, ``unet.set_attn_processor(dict)``, ``named_modules()`` with ``Attention`` objects exposing
``to_q/to_k/to_v/to_out/heads/processor`` - and the diffusers parameter names in checkpoints.  All of that
is kept.  Execution is B200-first:

* activations are bf16 **channels-last** end to end: convolutions hit cuDNN's NHWC tensor-core kernels (off-path by the
  north-star), the Transformer2D entry/exit permutes are free views and the 1x1 ``proj_in/out`` are plain GEMMs;
* every GroupNorm(+SiLU) is one ``dadd_groupnorm_fwd`` call, with the resnet time-embedding add folded into ``norm2``; the
  skip concatenations of the up path are never written (``dadd_groupnorm_cat_fwd`` + a shortcut GEMM split over its halves);
* no convolution adds its own bias (PyTorch would launch a broadcasting add per conv): conv1's bias rides on the
  time-embedding row, conv2's on the fused residual add (``dadd_bias_residual_fwd``) or the shortcut GEMM;
* LayerNorm is fused with the residual add that feeds it (``dadd_add_layernorm_fwd``); the feed-forward projection and its
  GEGLU gate are one tcgen05 GEMM (``dadd_ff_geglu_fwd``); the biases of ``ff.net[2]`` and ``proj_out`` ride on the
  residuals they are added to, so both projections are single accumulating GEMMs; self/cross attention are the
  processors' fused kernels;
* the 22 ``time_emb_proj`` matrices are one stacked fp32 GEMM per forward (or one per sampling call, see
  ``precompute_time_terms``), and nothing in ``forward`` synchronises, so a whole step can be captured in a CUDA graph.
"""

from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, wcache
from .attention_processor import AttnProcessor2_0, compute_dtype

CL = torch.channels_last


# ----------------------------------------------------------------------------------------------- leaf helpers
def _conv_nobias(mod: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    """cuDNN convolution without its bias: PyTorch adds a conv bias as a separate (slow, broadcasting) elementwise
    kernel, so callers fold it into the kernel that consumes the result instead."""
    return F.conv2d(x, wcache.conv_filter(mod, "w", mod.weight, compute_dtype()), None, mod.stride, mod.padding)


def _bias32(mod, tag: str = "b32") -> Optional[torch.Tensor]:
    return None if mod.bias is None else wcache.cast(mod, tag, mod.bias, torch.float32)


def _conv(mod: nn.Conv2d, x: torch.Tensor) -> torch.Tensor:
    y = _conv_nobias(mod, x)
    if mod.bias is None:
        return y
    if mod.out_channels % 8 == 0:
        return ops.bias_residual(y, None, _bias32(mod), out=y)      # one vectorised in-place pass
    return y + wcache.cast(mod, "b", mod.bias, compute_dtype()).view(1, -1, 1, 1)


def _linear(mod: nn.Linear, x: torch.Tensor, residual: Optional[torch.Tensor] = None, bias: bool = True) -> torch.Tensor:
    """``mod(x) (+ residual)`` through ``ops.linear``; ``bias=False`` leaves the bias out (it already rides on ``residual``)."""
    if not bias or mod.bias is None:
        return ops.linear(x, wcache.cast(mod, "w", mod.weight, compute_dtype()), None, residual)
    return ops.linear(x, wcache.cast(mod, "w", mod.weight, compute_dtype()), _bias32(mod), residual,
                      bias_lp=wcache.cast(mod, "b", mod.bias, compute_dtype()))


def _bias_sum32(*mods) -> torch.Tensor:
    """fp32 sum of the biases of ``mods`` (``None`` entries skipped), cached on the first module."""
    mods = [m for m in mods if m is not None]
    if len(mods) == 1:
        return _bias32(mods[0])
    srcs = tuple(m.bias for m in mods)
    return wcache.get(mods[0], "b32+" + ",".join(str(id(m)) for m in mods[1:]), srcs,
                      lambda: sum(p.detach().float() for p in srcs).contiguous())


def _conv1x1_weight(mod: nn.Conv2d, cols: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """(C_out, C_in) 16-bit GEMM weight of a 1x1 convolution, or the contiguous copy of its input columns [lo, hi)."""
    if cols is None:
        return wcache.get(mod, "w2d", (mod.weight,), lambda: mod.weight.detach().to(compute_dtype()).reshape(mod.out_channels, -1).contiguous())
    lo, hi = cols
    return wcache.get(mod, f"w2d[{lo}:{hi}]", (mod.weight,),
                      lambda: mod.weight.detach().to(compute_dtype()).reshape(mod.out_channels, -1)[:, lo:hi].contiguous())


def _conv1x1_as_linear(mod: nn.Conv2d, tokens: torch.Tensor, extra_bias: Optional[nn.Module] = None,
                       cols: Optional[Tuple[int, int]] = None, residual: Optional[torch.Tensor] = None,
                       bias: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """1x1 convolution as a GEMM on channels-last tokens, bias and ``residual`` (tokens) in the GEMM epilogue.
    ``extra_bias``: a module whose bias is added on top (a resnet's conv2 bias rides on its shortcut GEMM).  ``cols``: use
    only these input channels.  ``bias=False`` leaves the bias out (second half of a split GEMM)."""
    w = _conv1x1_weight(mod, cols)
    if not bias:
        b = b_lp = None
    elif extra_bias is None:
        b, b_lp = _bias32(mod), wcache.cast(mod, "b", mod.bias, compute_dtype())
    else:
        b = _bias_sum32(mod, *([extra_bias] if isinstance(extra_bias, nn.Module) else extra_bias))
        b_lp = wcache.get(mod, "b+lp", (b,), lambda: b.to(compute_dtype()))
    return ops.linear(tokens, w, b, residual, out=out, bias_lp=b_lp)


def _tokens(x: torch.Tensor) -> torch.Tensor:
    """(B, C, H, W) channels-last -> (B, H*W, C) view."""
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, h * w, c)


def _image(t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """(B, H*W, C) -> (B, C, H, W) channels-last view."""
    b, _, c = t.shape
    return t.view(b, h, w, c).permute(0, 3, 1, 2)


def _gn(mod: nn.GroupNorm, x: torch.Tensor, silu: bool, chan_add: Optional[torch.Tensor] = None) -> torch.Tensor:
    return ops.group_norm(x, wcache.cast(mod, "w", mod.weight, torch.float32), wcache.cast(mod, "b", mod.bias, torch.float32),
                          mod.num_groups, mod.eps, silu, chan_add)


def _ln(mod: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    return ops.layer_norm(x, wcache.cast(mod, "w", mod.weight, torch.float32), wcache.cast(mod, "b", mod.bias, torch.float32),
                          mod.eps)


def _add_ln(mod: nn.LayerNorm, x: torch.Tensor, r: torch.Tensor, sum_bias: Optional[torch.Tensor] = None):
    """(x + r (+ sum_bias), LayerNorm(x + r)) in one kernel."""
    return ops.add_layer_norm(x, r, wcache.cast(mod, "w", mod.weight, torch.float32),
                              wcache.cast(mod, "b", mod.bias, torch.float32), mod.eps, sum_bias=sum_bias)


# ----------------------------------------------------------------------------------------------- modules
class Attention(nn.Module):
    """Parameter container + processor hook with the attribute surface of diffusers' ``Attention``."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int] = None, heads: int = 8, dim_head: int = 64,
                 bias: bool = False, out_bias: bool = True) -> None:
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.is_cross_attention = cross_attention_dim is not None
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def set_processor(self, processor) -> None:
        if "processor" in self._modules and not isinstance(processor, nn.Module):
            self._modules.pop("processor")
        self.processor = processor

    def get_processor(self):
        return self.processor

    def prepare_attention_mask(self, *args, **kwargs):
        raise NotImplementedError("attention masks are not part of DADD's UNet path")

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kwargs)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int) -> None:
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """``ff.net = [GEGLU(proj), Dropout, Linear]`` (keys ``ff.net.0.proj``, ``ff.net.2``)."""

    def __init__(self, dim: int, mult: int = 4) -> None:
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None, out_bias: bool = True) -> torch.Tensor:
        # ``net(x) (+ residual)``: the residual add of the transformer block rides on the output GEMM (``out_bias=False``: the
        # caller has already put net[2].bias on the residual, so the GEMM just accumulates onto it)
        # (slicing the batch so that proj -> GEGLU -> out stays inside L2 was measured slower than one pass: the smaller
        # GEMMs lose more than the L2-resident intermediate gains; profiles/r01_ff_slice_ab.txt)
        proj = self.net[0].proj
        if proj.out_features % 256 == 0 and x.is_contiguous():       # (M, 8C) projection + GEGLU in one tcgen05 GEMM
            g = ops.ff_geglu(x, wcache.cast(proj, "w", proj.weight, compute_dtype()), _bias32(proj))
        else:
            g = ops.geglu(_linear(proj, x))
        return _linear(self.net[2], g, residual, bias=out_bias)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int) -> None:
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x: torch.Tensor, ehs: torch.Tensor) -> torch.Tensor:
        # hidden = attn(norm(hidden)) + hidden, three times; each residual add is fused with the
        # LayerNorm that follows it, the last one with the feed-forward's output GEMM: the stored residual already carries
        # ff.net[2].bias, so ``ff(n) + hidden`` is one GEMM accumulating onto it
        x, n = _add_ln(self.norm2, x, self.attn1(_ln(self.norm1, x)))
        x, n = _add_ln(self.norm3, x, self.attn2(n, encoder_hidden_states=ehs), sum_bias=_bias32(self.ff.net[2]))
        return self.ff(n, residual=x, out_bias=False)


class Transformer2DModel(nn.Module):
    """GN(eps 1e-6) -> 1x1 -> tokens -> BasicTransformerBlock -> 1x1 -> + residual."""

    def __init__(self, channels: int, heads: int, cross_attention_dim: int, groups: int = 32) -> None:
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(channels, channels, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(channels, heads, channels // heads, cross_attention_dim)])
        self.proj_out = nn.Conv2d(channels, channels, 1)

    def forward(self, x: torch.Tensor, ehs: torch.Tensor, x_carries_out_bias: bool = False) -> torch.Tensor:
        """``x_carries_out_bias``: the producer of ``x`` (the resnet before this block, whose only consumers are the norm and
        the residual add below) has already added ``proj_out.bias`` to it.  The norm then takes ``-proj_out.bias`` as its
        per-channel additive term and ``proj_out(t) + x`` is one GEMM accumulating onto ``x`` - no bias / residual pass."""
        b, c, h, w = x.shape
        shift = None
        if x_carries_out_bias:
            shift = wcache.get(self.proj_out, "-b32", (self.proj_out.bias,), lambda: -self.proj_out.bias.detach().float())
            shift = shift.view(1, -1).expand(b, -1)
        t = _tokens(_gn(self.norm, x, silu=False, chan_add=shift))   # free view (channels-last)
        t = _conv1x1_as_linear(self.proj_in, t)
        for blk in self.transformer_blocks:
            t = blk(t, ehs)
        return _image(_conv1x1_as_linear(self.proj_out, t, residual=_tokens(x), bias=not x_carries_out_bias), h, w)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb_channels: Optional[int] = 1280, groups: int = 32, eps: float = 1e-5) -> None:
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, cout) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x: torch.Tensor, temb_term: Optional[torch.Tensor], skip: Optional[torch.Tensor] = None,
                carry_bias: Optional[nn.Module] = None) -> torch.Tensor:
        """``temb_term`` = time_emb_proj(silu(emb)) + conv1.bias as fp32 (B, cout) (``UNet2DConditionModel.time_terms``):
        folded into norm2's input by the GN kernel.  No convolution adds its own bias: conv1's rides on ``temb_term``,
        conv2's on the residual add (or on the shortcut GEMM's epilogue).

        ``skip``: the block input is ``cat([x, skip], dim=1)`` (up path).  The concatenation is never written: norm1 reads
        both tensors (``dadd_groupnorm_cat_fwd``) and the 1x1 shortcut is two accumulating GEMMs over the two halves of
        its weight.  ``carry_bias``: a module whose bias is added to the output on top (``Transformer2DModel.forward``)."""
        if skip is not None and not ops.group_norm_cat_supported(x, skip, self.norm1.num_groups):
            x, skip = torch.cat([x, skip], dim=1), None
        if skip is not None:
            n1 = ops.group_norm_cat(x, skip, wcache.cast(self.norm1, "w", self.norm1.weight, torch.float32),
                                    wcache.cast(self.norm1, "b", self.norm1.bias, torch.float32), self.norm1.num_groups,
                                    self.norm1.eps, True)
        else:
            n1 = _gn(self.norm1, x, silu=True)
        h = _conv_nobias(self.conv1, n1)
        if temb_term is None:                                       # VAE resnets: only conv1's bias, one row for all samples
            temb_term = _bias32(self.conv1).view(1, -1).expand(x.shape[0], -1)
        h = _conv_nobias(self.conv2, _gn(self.norm2, h, silu=True, chan_add=temb_term))
        if self.conv_shortcut is None:
            return ops.bias_residual(h, x, _bias_sum32(self.conv2, carry_bias), out=h)
        # out = shortcut(x) + conv2(h) + both biases: the 1x1 shortcut GEMM takes h as the residual of its epilogue; with a
        # skip tensor it is two accumulating GEMMs over the two halves of its weight (the second adds onto the first in place)
        b, c1, hh, ww = x.shape
        if skip is None:
            out = _conv1x1_as_linear(self.conv_shortcut, _tokens(x), extra_bias=(self.conv2, carry_bias), residual=_tokens(h))
        else:
            out = _conv1x1_as_linear(self.conv_shortcut, _tokens(x), extra_bias=(self.conv2, carry_bias), cols=(0, c1),
                                     residual=_tokens(h))
            out = _conv1x1_as_linear(self.conv_shortcut, _tokens(skip), cols=(c1, c1 + skip.shape[1]), residual=out, bias=False,
                                     out=out)
        return _image(out, hh, ww)


class Downsample2D(nn.Module):
    def __init__(self, c: int) -> None:
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return _conv(self.conv, x)


class Upsample2D(nn.Module):
    def __init__(self, c: int) -> None:
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return _conv(self.conv, ops.upsample_nearest2x(x))


class _Block(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.resnets = nn.ModuleList()
        self.attentions = nn.ModuleList()


class TimestepEmbedding(nn.Module):
    def __init__(self, cin: int, dim: int) -> None:
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)


class UNet2DConditionModel(nn.Module):
    def __init__(self, in_channels: int = 4, out_channels: int = 4, block_out_channels=(320, 640, 1280, 1280),
                 layers_per_block: int = 2, cross_attention_dim: int = 768, attention_head_dim: int = 8,
                 norm_num_groups: int = 32) -> None:
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      cross_attention_dim=cross_attention_dim, attention_head_dim=attention_head_dim,
                                      norm_num_groups=norm_num_groups, norm_eps=1e-5, act_fn="silu",
                                      flip_sin_to_cos=True, freq_shift=0)
        heads = attention_head_dim          # SD-1.x: ``attention_head_dim`` is the number of heads
        ch = list(block_out_channels)
        tdim = ch[0] * 4
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], tdim)
        self.down_blocks = nn.ModuleList()
        cin = ch[0]
        skip_ch = [ch[0]]
        for i, cout in enumerate(ch):
            blk = _Block()
            for j in range(layers_per_block):
                blk.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, tdim, norm_num_groups))
                if i < len(ch) - 1:
                    blk.attentions.append(Transformer2DModel(cout, heads, cross_attention_dim, norm_num_groups))
                skip_ch.append(cout)
            if i < len(ch) - 1:
                blk.downsamplers = nn.ModuleList([Downsample2D(cout)])
                skip_ch.append(cout)
            else:
                blk.downsamplers = None
            self.down_blocks.append(blk)
            cin = cout
        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(ch[-1], ch[-1], tdim, norm_num_groups))
        self.mid_block.attentions.append(Transformer2DModel(ch[-1], heads, cross_attention_dim, norm_num_groups))
        self.mid_block.resnets.append(ResnetBlock2D(ch[-1], ch[-1], tdim, norm_num_groups))
        self.up_blocks = nn.ModuleList()
        rev = ch[::-1]
        prev = ch[-1]
        for i, cout in enumerate(rev):
            blk = _Block()
            for j in range(layers_per_block + 1):
                skip = skip_ch.pop()
                blk.resnets.append(ResnetBlock2D(prev + skip, cout, tdim, norm_num_groups))
                if i > 0:
                    blk.attentions.append(Transformer2DModel(cout, heads, cross_attention_dim, norm_num_groups))
                prev = cout
            blk.upsamplers = nn.ModuleList([Upsample2D(cout)]) if i < len(ch) - 1 else None
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, ch[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)
        self._time_terms: Optional[Tuple[torch.Tensor, int]] = None   # (current-step row, row length) set by samplers

    # ------------------------------------------------------------------ diffusers processor plumbing
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": m.processor for name, m in self.named_modules() if isinstance(m, Attention)}

    def set_attn_processor(self, processor) -> None:
        sites = {f"{name}.processor": m for name, m in self.named_modules() if isinstance(m, Attention)}
        if isinstance(processor, dict):
            if len(processor) != len(sites):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not "
                                 f"match the number of attention layers: {len(sites)}.")
            for key, m in sites.items():
                m.set_processor(processor[key])
        else:
            for m in sites.values():
                m.set_processor(processor)

    # ------------------------------------------------------------------ time embedding
    def _resnets(self) -> List[ResnetBlock2D]:
        cached = self.__dict__.get("_resnet_list")
        if cached is None:
            cached = [m for m in self.modules() if isinstance(m, ResnetBlock2D)]
            self.__dict__["_resnet_list"] = cached
        return cached

    def time_terms(self, timesteps: torch.Tensor) -> torch.Tensor:
        """(T, sum C_out) fp32: sinusoid -> time_embedding MLP -> SiLU -> all 22 ``time_emb_proj`` as one stacked GEMM
        (SURVEY.md A.2 step 1, A.3), plus each resnet's ``conv1.bias``.  Row t is what every resnet adds to the bias-free
        conv1 output before norm2 at that timestep."""
        half = self.config.block_out_channels[0] // 2
        dev = timesteps.device
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=dev) / half)
        args = timesteps.to(torch.float32)[:, None] * freqs[None, :]
        t_emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        te = self.time_embedding
        f32 = torch.float32
        emb = F.linear(t_emb, wcache.cast(te.linear_1, "w", te.linear_1.weight, f32), wcache.cast(te.linear_1, "b", te.linear_1.bias, f32))
        emb = F.linear(F.silu(emb), wcache.cast(te.linear_2, "w", te.linear_2.weight, f32), wcache.cast(te.linear_2, "b", te.linear_2.bias, f32))
        res = self._resnets()
        wsrc = tuple(r.time_emb_proj.weight for r in res)
        bsrc = tuple(r.time_emb_proj.bias for r in res) + tuple(r.conv1.bias for r in res)
        w = wcache.get(self, "temb_w", wsrc, lambda: torch.cat([p.detach().to(f32) for p in wsrc], 0).contiguous())
        # conv1's bias is a per-channel constant added right before the time term: it rides on the same row
        b = wcache.get(self, "temb_b", bsrc, lambda: torch.cat([r.time_emb_proj.bias.detach().to(f32) + r.conv1.bias.detach().to(f32)
                                                                for r in res], 0).contiguous())
        return F.linear(F.silu(emb), w, b)

    def _split_terms(self, terms: torch.Tensor) -> Dict[int, torch.Tensor]:
        out, off = {}, 0
        for r in self._resnets():
            c = r.time_emb_proj.out_features
            out[id(r)] = terms[:, off:off + c]
            off += c
        return out

    # ------------------------------------------------------------------ forward
    def forward(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor,
                time_terms: Optional[torch.Tensor] = None, return_dict: bool = True):
        """``time_terms``: optional precomputed (B or 1, sum C_out) fp32 row(s) of ``self.time_terms`` for this step."""
        bsz = sample.shape[0]
        if time_terms is None:
            if not torch.is_tensor(timestep):
                timestep = torch.tensor([timestep], dtype=torch.long, device=sample.device)
            elif timestep.ndim == 0:
                timestep = timestep[None]
            time_terms = self.time_terms(timestep.to(sample.device).expand(bsz))
        if time_terms.shape[0] != bsz:
            time_terms = time_terms.expand(bsz, -1)      # stride-0 rows: the GN kernel takes the row stride
        terms = self._split_terms(time_terms)
        ehs = encoder_hidden_states

        x = sample.to(compute_dtype()).contiguous(memory_format=CL)
        x = _conv(self.conv_in, x)
        skips = [x]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                # a resnet followed by a Transformer2D hands it proj_out's bias in advance (see Transformer2DModel.forward)
                attn = blk.attentions[j] if len(blk.attentions) > 0 else None
                x = res(x, terms[id(res)], carry_bias=attn.proj_out if attn is not None else None)
                if attn is not None:
                    x = attn(x, ehs, x_carries_out_bias=True)
                skips.append(x)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0](x)
                skips.append(x)
        mb = self.mid_block
        x = mb.resnets[0](x, terms[id(mb.resnets[0])], carry_bias=mb.attentions[0].proj_out)
        x = mb.attentions[0](x, ehs, x_carries_out_bias=True)
        x = mb.resnets[1](x, terms[id(mb.resnets[1])])
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                attn = blk.attentions[j] if len(blk.attentions) > 0 else None
                x = res(x, terms[id(res)], skip=skips.pop(), carry_bias=attn.proj_out if attn is not None else None)
                if attn is not None:
                    x = attn(x, ehs, x_carries_out_bias=True)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0](x)
        x = _gn(self.conv_norm_out, x, silu=True)
        eps = _conv(self.conv_out, x).to(sample.dtype).contiguous()
        return SimpleNamespace(sample=eps) if return_dict else (eps,)
