"""``SDVAE`` with the interface of ``/root/reference/src/models/vae/vae.py`` (:32-112) over a B200 AutoencoderKL decoder.

Only ``decode`` is on the progression path (``_latents_to_images``, inference_pipeline_ip.py:473-486).  The decoder keeps
diffusers' parameter names (``vae.vae.decoder.*``, ``post_quant_conv``) and runs channels-last: convolutions on cuDNN
(off-path), every GroupNorm(+SiLU) through ``dadd_groupnorm_fwd``; the single-head d=512 mid-block attention goes to
``F.scaled_dot_product_attention`` (library; SURVEY.md 8f row f2 lists the VAE as the next tier).  ``encode`` (the frozen first
step of ``training_step``, diffusion_module_ip.py:419-420) needs the encoder half, built on ``build_encoder=True`` with the
same blocks; a module without it raises.
"""

from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .unet2d import CL, ResnetBlock2D, Upsample2D, _Block, _conv, _gn, _linear, Attention


class _Decoder(nn.Module):
    def __init__(self, latent_channels: int = 4, out_channels: int = 3, chans=(128, 256, 512, 512), layers: int = 2) -> None:
        super().__init__()
        top = chans[-1]
        self.conv_in = nn.Conv2d(latent_channels, top, 3, padding=1)
        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(top, top, None, eps=1e-6))
        attn = Attention(top, None, heads=1, dim_head=top, bias=True)
        attn.group_norm = nn.GroupNorm(32, top, eps=1e-6, affine=True)
        attn.residual_connection = True
        self.mid_block.attentions.append(attn)
        self.mid_block.resnets.append(ResnetBlock2D(top, top, None, eps=1e-6))
        self.up_blocks = nn.ModuleList()
        cin = top
        for i, cout in enumerate(chans[::-1]):
            blk = _Block()
            for j in range(layers + 1):
                blk.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, None, eps=1e-6))
            blk.upsamplers = nn.ModuleList([Upsample2D(cout)]) if i < len(chans) - 1 else None
            self.up_blocks.append(blk)
            cin = cout
        self.conv_norm_out = nn.GroupNorm(32, chans[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(chans[0], out_channels, 3, padding=1)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        x = _conv(self.conv_in, z)
        x = self.mid_block.resnets[0](x, None)
        x = _mid_attention(self.mid_block.attentions[0], x)
        x = self.mid_block.resnets[1](x, None)
        for blk in self.up_blocks:
            for res in blk.resnets:
                x = res(x, None)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0](x)
        return _conv(self.conv_out, _gn(self.conv_norm_out, x, silu=True))


class _Downsample(nn.Module):
    """diffusers ``Downsample2D(padding=0)``: pad right / bottom by one, then a stride-2 3x3 convolution."""

    def __init__(self, c: int) -> None:
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return _conv(self.conv, F.pad(x, (0, 1, 0, 1)).contiguous(memory_format=CL))


def _mid_attention(a: Attention, x: torch.Tensor) -> torch.Tensor:
    b, c, h, w = x.shape
    t = _gn(a.group_norm, x, silu=False).permute(0, 2, 3, 1).reshape(b, h * w, c)
    q, k, v = _linear(a.to_q, t), _linear(a.to_k, t), _linear(a.to_v, t)
    o = ops.self_attention(q, k, v, heads=1)          # one 512-wide head: dadd_self_attn_fwd's wide-head kernel
    return x + _linear(a.to_out[0], o).view(b, h, w, c).permute(0, 3, 1, 2)


class _Encoder(nn.Module):
    """SD AutoencoderKL encoder half (diffusers ``Encoder``; keys ``encoder.*``): 3 -> 128 -> (128, 256, 512, 512) -> 8 moments."""

    def __init__(self, in_channels: int = 3, latent_channels: int = 4, chans=(128, 256, 512, 512), layers: int = 2) -> None:
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, chans[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cin = chans[0]
        for i, cout in enumerate(chans):
            blk = _Block()
            for j in range(layers):
                blk.resnets.append(ResnetBlock2D(cin if j == 0 else cout, cout, None, eps=1e-6))
            blk.downsamplers = nn.ModuleList([_Downsample(cout)]) if i < len(chans) - 1 else None
            self.down_blocks.append(blk)
            cin = cout
        top = chans[-1]
        self.mid_block = _Block()
        self.mid_block.resnets.append(ResnetBlock2D(top, top, None, eps=1e-6))
        attn = Attention(top, None, heads=1, dim_head=top, bias=True)
        attn.group_norm = nn.GroupNorm(32, top, eps=1e-6, affine=True)
        attn.residual_connection = True
        self.mid_block.attentions.append(attn)
        self.mid_block.resnets.append(ResnetBlock2D(top, top, None, eps=1e-6))
        self.conv_norm_out = nn.GroupNorm(32, top, eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(top, 2 * latent_channels, 3, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _conv(self.conv_in, x)
        for blk in self.down_blocks:
            for res in blk.resnets:
                x = res(x, None)
            if blk.downsamplers is not None:
                x = blk.downsamplers[0](x)
        x = self.mid_block.resnets[0](x, None)
        x = _mid_attention(self.mid_block.attentions[0], x)
        x = self.mid_block.resnets[1](x, None)
        return _conv(self.conv_out, _gn(self.conv_norm_out, x, silu=True))


class DiagonalGaussianDistribution:
    """diffusers' posterior object: ``moments`` (B, 8, h, w) = [mean | logvar], logvar clamped to [-30, 20]."""

    def __init__(self, moments: torch.Tensor) -> None:
        self.mean, logvar = torch.chunk(moments.float(), 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        return self.mean + self.std * torch.randn(self.mean.shape, device=self.mean.device, dtype=self.mean.dtype, generator=generator)

    def mode(self) -> torch.Tensor:
        return self.mean


class AutoencoderKL(nn.Module):
    def __init__(self, build_encoder: bool = False) -> None:
        super().__init__()
        self.decoder = _Decoder()
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        if build_encoder:
            self.encoder = _Encoder()
            self.quant_conv = nn.Conv2d(8, 8, 1)
        self.config = SimpleNamespace(latent_channels=4, scaling_factor=0.18215)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        from .attention_processor import compute_dtype
        x = z.to(compute_dtype()).contiguous(memory_format=CL)
        img = self.decoder(_conv(self.post_quant_conv, x))
        return SimpleNamespace(sample=img) if return_dict else (img,)

    def encode(self, images: torch.Tensor, return_dict: bool = True):
        """images (B, 3, H, W) in [-1, 1] -> ``.latent_dist`` (reference use: ``vae.encode(images).latent_dist.sample()``)."""
        if not hasattr(self, "encoder"):
            raise NotImplementedError("this AutoencoderKL was built without its encoder half (build_encoder=True builds it)")
        from .attention_processor import compute_dtype
        x = images.to(compute_dtype()).contiguous(memory_format=CL)
        post = DiagonalGaussianDistribution(_conv(self.quant_conv, self.encoder(x)))
        return SimpleNamespace(latent_dist=post) if return_dict else (post,)


class SDVAE(nn.Module):
    def __init__(self, pretrained_path=None, *, torch_dtype: Optional[torch.dtype] = None, local_files_only: bool = False,
                 build_encoder: bool = False) -> None:
        super().__init__()
        self.vae = AutoencoderKL(build_encoder=build_encoder)
        if torch_dtype is not None:
            self.vae.to(torch_dtype)
        self.vae.eval()
        self.vae.requires_grad_(False)

    @torch.no_grad()
    def encode(self, images: torch.Tensor, *, return_dict: bool = True):
        return self.vae.encode(images, return_dict=return_dict)

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, *, return_dict: bool = True):
        return self.vae.decode(latents, return_dict=return_dict)
