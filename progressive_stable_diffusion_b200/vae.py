"""``SDVAE`` with the interface of ``/root/reference/src/models/vae/vae.py`` (:32-112) over a B200 AutoencoderKL decoder.

Only ``decode`` is on the progression path (``_latents_to_images``, inference_pipeline_ip.py:473-486).  The decoder keeps
diffusers' parameter names (``vae.vae.decoder.*``, ``post_quant_conv``) and runs channels-last: convolutions on cuDNN
(off-path), every GroupNorm(+SiLU) through ``dadd_groupnorm_fwd``; the single-head d=512 mid-block attention goes to
``F.scaled_dot_product_attention`` (library; SURVEY.md 8f row f2 lists the VAE as the next tier).  ``encode`` is training-only
(next tier) and raises.
"""

from __future__ import annotations

from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

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
        a = self.mid_block.attentions[0]
        b, c, h, w = x.shape
        t = _gn(a.group_norm, x, silu=False).permute(0, 2, 3, 1).reshape(b, h * w, c)
        q, k, v = _linear(a.to_q, t), _linear(a.to_k, t), _linear(a.to_v, t)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        x = x + _linear(a.to_out[0], o).view(b, h, w, c).permute(0, 3, 1, 2)
        x = self.mid_block.resnets[1](x, None)
        for blk in self.up_blocks:
            for res in blk.resnets:
                x = res(x, None)
            if blk.upsamplers is not None:
                x = blk.upsamplers[0](x)
        return _conv(self.conv_out, _gn(self.conv_norm_out, x, silu=True))


class AutoencoderKL(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.decoder = _Decoder()
        self.post_quant_conv = nn.Conv2d(4, 4, 1)
        self.config = SimpleNamespace(latent_channels=4, scaling_factor=0.18215)

    def decode(self, z: torch.Tensor, return_dict: bool = True):
        from .attention_processor import compute_dtype
        x = z.to(compute_dtype()).contiguous(memory_format=CL)
        img = self.decoder(_conv(self.post_quant_conv, x))
        return SimpleNamespace(sample=img) if return_dict else (img,)

    def encode(self, *args, **kwargs):
        raise NotImplementedError("VAE encode is training-only (SURVEY.md 8f, next tier); the B200 build covers decode")


class SDVAE(nn.Module):
    def __init__(self, pretrained_path=None, *, torch_dtype: Optional[torch.dtype] = None, local_files_only: bool = False) -> None:
        super().__init__()
        self.vae = AutoencoderKL()
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
