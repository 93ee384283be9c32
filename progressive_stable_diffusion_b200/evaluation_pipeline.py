"""Batched evaluation sampling for B200: the sampling half of
``/root/reference/src/pipelines/evaluation/evaluation_pipeline.py`` (``_prepare_conditioning`` :406-461,
``_set_delta_scale`` :464-468, ``_ddim_sample_batched`` :471-564, ``_decode_latents`` :567-574, job batching of
``generate_all`` :867-975).  The FID / CMMD / P&R metrics of that file use third-party backbones and are not on the
denoising path.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import parallel
from .inference_pipeline_ip import (_latents_to_images, _prepare_conditioning as _prep, _sample,
                                    _set_delta_scale_on_processors)


def _prepare_conditioning(module, target_labels: Tensor, source_labels: Tensor, structure_images: Tensor,
                          image_scale: float = 1.0, zero_aoe: bool = False) -> Tensor:
    return _prep(module, target_labels, source_labels, structure_images, image_scale=image_scale, zero_aoe=zero_aoe)


def _set_delta_scale(module, scale: float) -> None:
    _set_delta_scale_on_processors(module, scale)


@torch.no_grad()
def _ddim_sample_batched(module, target_labels: Tensor, source_labels: Tensor, structure_images: Tensor,
                         sampling_steps: int, device: torch.device, eta: float = 0.0, image_scale: float = 1.0,
                         steer_scale: float = 0.0, guidance_scale: float = 1.0, *, init_latents: Optional[Tensor] = None,
                         use_graph: bool = True) -> Tensor:
    """Like ``_ddim_sample_ip`` but every sample draws its own initial noise (reference :506)."""
    device = torch.device(device)
    routing = getattr(module.diff_cfg, "use_routing_gates", True)
    do_cfg = (not routing) and (guidance_scale != 1.0)
    b = target_labels.shape[0]
    h = module.cfg.dataset.image_size // 8
    if sampling_steps > module.diff_cfg.num_train_timesteps:
        raise ValueError(f"sampling_steps={sampling_steps} must be <= num_train_timesteps={module.diff_cfg.num_train_timesteps}")
    if init_latents is None:
        init_latents = torch.randn(b, module.cfg.model.latent_channels, h, h, device=device, dtype=torch.float32)
    return _sample(module, target_labels, source_labels, structure_images, init_latents.to(device), sampling_steps, device,
                   eta, image_scale, None, steer_scale, guidance_scale, do_cfg, use_graph)


@torch.no_grad()
def _decode_latents(module, latents: Tensor) -> Tensor:
    """latents -> [0,1] RGB (B,3,H,W) on the CPU (reference :567-574)."""
    return _latents_to_images(module, latents).float().cpu()


@torch.no_grad()
def generate_all(module, jobs: Sequence[Tuple[int, float, float]], image_tokens: Tensor, device: torch.device,
                 batch_size: int = 12, sampling_steps: int = 50, image_scale: float = 1.0, steer_scale: float = 0.0,
                 guidance_scale: float = 1.0, eta: float = 0.0, seed: int = 42, rank: int = 0, world_size: int = 1,
                 decode: bool = True) -> Dict[int, Tensor]:
    """Run this rank's share of an evaluation sweep.  ``jobs`` = (source index into ``image_tokens``, source label,
    target label), already in the reference's sorted order (:897-903).  The reference seeds once and draws noise batch by
    batch (:909,506); to keep a sharded run identical to the single-process one, ALL initial noise is drawn up front in
    job order from one seeded generator and each rank takes its slice."""
    device = torch.device(device)
    h = module.cfg.dataset.image_size // 8
    c = module.cfg.model.latent_channels
    gen = torch.Generator(device="cpu").manual_seed(seed)
    noise = torch.randn(len(jobs), c, h, h, generator=gen, dtype=torch.float32)
    mine = parallel.shard_indices(len(jobs), rank, world_size)
    out: Dict[int, Tensor] = {}
    for s in range(0, len(mine), batch_size):
        idx = mine[s:s + batch_size]
        if len(idx) < batch_size:                       # keep the captured batch shape: pad with repeats, drop after
            idx_run = idx + [idx[-1]] * (batch_size - len(idx))
        else:
            idx_run = idx
        src = torch.tensor([jobs[i][1] for i in idx_run], dtype=torch.float32, device=device)
        tgt = torch.tensor([jobs[i][2] for i in idx_run], dtype=torch.float32, device=device)
        tok = image_tokens[[jobs[i][0] for i in idx_run]].to(device)
        lat = _ddim_sample_batched(module, tgt, src, tok, sampling_steps, device, eta, image_scale, steer_scale,
                                   guidance_scale, init_latents=noise[idx_run].to(device))
        res = _decode_latents(module, lat) if decode else lat.cpu()
        for k, i in enumerate(idx):
            out[i] = res[k]
    return out
