"""``OrdinalUNet`` / ``UNetConfig`` with the interface of ``/root/reference/src/models/unet/unet.py`` (:21-48, :51-146).

The reference loads diffusers' pretrained ``UNet2DConditionModel``; there is no network (and no diffusers) here, so the
wrapped ``.unet`` is this package's SD-1.x-shaped B200 model (``unet2d.UNet2DConditionModel``), random-initialised, whose
parameter names are diffusers' - a real ``CompVis/stable-diffusion-v1-4`` UNet state dict loads with ``load_state_dict``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor, nn

from .unet2d import UNet2DConditionModel


@dataclass
class UNetConfig:
    pretrained_unet_path: str = "CompVis/stable-diffusion-v1-4"
    conditioning_dim: int = 768
    in_channels: int = 4
    out_channels: int = 4
    torch_dtype: Optional[torch.dtype] = None
    local_files_only: bool = False


class OrdinalUNet(nn.Module):
    def __init__(self, config: UNetConfig) -> None:
        super().__init__()
        self.config = config
        self.unet = UNet2DConditionModel()
        if config.torch_dtype is not None:
            self.unet.to(config.torch_dtype)
        # the same three consistency checks as the reference (:78-94)
        for field, have, want in (("in_channels", self.unet.config.in_channels, config.in_channels),
                                  ("out_channels", self.unet.config.out_channels, config.out_channels),
                                  ("cross_attention_dim", self.unet.config.cross_attention_dim, config.conditioning_dim)):
            if have != want:
                raise ValueError(f"UNet {field} mismatch: {have} (from weights) vs {want} (config).")

    def forward(self, latents: Tensor, timesteps: Tensor, cond_embed: Tensor, time_terms: Optional[Tensor] = None) -> Tensor:
        if cond_embed.ndim == 2:
            encoder_hidden_states = cond_embed.unsqueeze(1)
        elif cond_embed.ndim == 3:
            encoder_hidden_states = cond_embed
        else:
            raise ValueError(f"cond_embed must have shape (B, D) or (B, seq_len, D), got {cond_embed.shape}")
        if timesteps is not None:
            if timesteps.ndim == 0:
                timesteps = timesteps[None]
            elif timesteps.ndim > 1:
                timesteps = timesteps.view(-1)
            timesteps = timesteps.to(latents.device)
        return self.unet(sample=latents, timestep=timesteps, encoder_hidden_states=encoder_hidden_states,
                         time_terms=time_terms).sample
