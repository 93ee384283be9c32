"""Oracle restatement of the DDIM progression sampler (TEST INFRASTRUCTURE).

Follows ``_ddim_sample_ip`` (src/pipelines/inference/inference_pipeline_ip.py:321-470) and its batched copy
``_ddim_sample_batched`` (src/pipelines/evaluation/evaluation_pipeline.py:471-564).  The reference file is not
importable here (omegaconf / lightning / diffusers missing); constants are pinned by SURVEY.md Appendix B
(tests/test_oracle_invariants.py).  fp32 throughout, scalar coefficients computed with torch fp32 ops exactly as the
reference does (0-dim tensor arithmetic), so the product's coefficient table can be compared bit for bit.
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from .conditioning import prepare_conditioning
from .unet import CrossCfg, unet_forward


def build_noise_schedule(beta_start: float = 0.00085, beta_end: float = 0.012, T: int = 1000):
    """``_build_noise_schedule`` (diffusion_module_ip.py:274-287): linear in beta (not SD's scaled-linear)."""
    betas = torch.linspace(beta_start, beta_end, T, dtype=torch.float32)
    return betas, torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(T: int = 1000, sampling_steps: int = 50) -> torch.Tensor:
    """inference_pipeline_ip.py:390-396."""
    return torch.linspace(T - 1, 0, steps=sampling_steps, dtype=torch.long)


def build_labels(num_steps: int, start: float = 0.0, end: float = 3.0) -> torch.Tensor:
    """``_build_labels`` (inference_pipeline_ip.py:184-195)."""
    if num_steps <= 0:
        raise ValueError("`mes_steps` must be a positive integer.")
    return torch.linspace(start, end, steps=num_steps, dtype=torch.float32)


def ddim_coefficients(alphas_cumprod: torch.Tensor, timesteps: torch.Tensor, eta: float = 0.0) -> List[Dict[str, float]]:
    """Per-step scalars of inference_pipeline_ip.py:434-468 as Python floats of the fp32 tensors the reference forms."""
    out = []
    n = len(timesteps)
    for i in range(n):
        ab = alphas_cumprod[int(timesteps[i])]
        row = {"sqrt_ab": torch.sqrt(ab).item(), "sqrt_1mab": torch.sqrt(1.0 - ab).item(), "last": i == n - 1}
        if i < n - 1:
            abp = alphas_cumprod[int(timesteps[i + 1])]
            row["sqrt_abp"] = torch.sqrt(abp).item()
            if eta == 0.0:
                row["eps_coef"] = torch.sqrt(1.0 - abp).item()
                row["sigma"] = 0.0
            else:
                sigma = eta * torch.sqrt((1 - abp) / (1 - ab) * (1 - ab / abp))
                row["eps_coef"] = torch.sqrt(1 - abp - sigma ** 2).item()
                row["sigma"] = sigma.item()
        out.append(row)
    return out


def ddim_update(x: torch.Tensor, eps: torch.Tensor, alphas_cumprod: torch.Tensor, t: int, t_prev: Optional[int],
                eta: float = 0.0, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One iteration body of inference_pipeline_ip.py:434-468 (t_prev None = last iteration -> clamped x0)."""
    ab = alphas_cumprod[t].to(dtype=x.dtype)
    x0 = ((x - torch.sqrt(1.0 - ab) * eps) / torch.sqrt(ab)).clamp(-4.0, 4.0)
    if t_prev is None:
        return x0
    abp = alphas_cumprod[t_prev].to(dtype=x.dtype)
    if eta == 0.0:
        return torch.sqrt(abp) * x0 + torch.sqrt(1.0 - abp) * eps
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab) * (1 - ab / abp))
    return torch.sqrt(abp) * x0 + torch.sqrt(1 - abp - sigma ** 2) * eps + sigma * noise


def cfg_combine(eps_cond: torch.Tensor, eps_uncond: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """inference_pipeline_ip.py:427-430."""
    return eps_uncond + guidance_scale * (eps_cond - eps_uncond)


def ddim_sample(
    unet_w: Dict[str, torch.Tensor],
    aoe_w: Dict[str, torch.Tensor],
    purifier_w: Optional[Dict[str, torch.Tensor]],
    target_labels: torch.Tensor,
    source_labels: torch.Tensor,
    image_embeds: torch.Tensor,
    init_latents: torch.Tensor,
    sampling_steps: int = 50,
    eta: float = 0.0,
    image_scale: float = 1.0,
    steer_scale: float = 0.0,
    guidance_scale: float = 1.0,
    use_routing_gates: bool = True,
    T: int = 1000,
    step_noise: Optional[Callable[[int], torch.Tensor]] = None,
    eps_trace: Optional[list] = None,
    max_steps: Optional[int] = None,
) -> torch.Tensor:
    """The sampling loop.  ``init_latents`` is injected (the reference draws it on ``device``,
    inference_pipeline_ip.py:377-385: one (1,4,h,w) tensor repeated over the levels; evaluation_pipeline.py:506 draws
    (B,4,h,w)).  ``image_embeds`` stands for ``module._get_image_embeds(...)`` (CLIP + resampler, off-path)."""
    if sampling_steps > T:
        raise ValueError(f"sampling_steps={sampling_steps} must be <= num_train_timesteps={T}")
    do_cfg = (not use_routing_gates) and (guidance_scale != 1.0)
    n = target_labels.shape[0]
    x = init_latents.to(torch.float32)
    if x.shape[0] == 1 and n > 1:
        x = x.repeat(n, 1, 1, 1)
    _, ac = build_noise_schedule(T=T)
    ts = ddim_timesteps(T, sampling_steps)
    cond = prepare_conditioning(aoe_w, purifier_w, target_labels, source_labels, image_embeds, use_routing_gates, image_scale)
    uncond = None
    if do_cfg:
        uncond = prepare_conditioning(aoe_w, purifier_w, target_labels, source_labels, image_embeds, use_routing_gates,
                                      image_scale, zero_aoe=True)
    cfg = CrossCfg(use_routing_gates=use_routing_gates, delta_scale=steer_scale)
    steps = sampling_steps if max_steps is None else min(max_steps, sampling_steps)
    for i in range(steps):
        t_int = int(ts[i].item())
        t = torch.full((n,), t_int, dtype=torch.long)
        if do_cfg:
            eps = cfg_combine(unet_forward(unet_w, x, t, cond, cfg), unet_forward(unet_w, x, t, uncond, cfg), guidance_scale)
        else:
            eps = unet_forward(unet_w, x, t, cond, cfg)
        if eps_trace is not None:
            eps_trace.append(eps.clone())
        t_prev = None if i == sampling_steps - 1 else int(ts[i + 1].item())
        noise = step_noise(i) if (eta != 0.0 and t_prev is not None and step_noise is not None) else (
            torch.zeros_like(x) if eta != 0.0 else None)
        x = ddim_update(x, eps, ac, t_int, t_prev, eta, noise)
    return x
