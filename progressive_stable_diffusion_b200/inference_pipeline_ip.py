"""DDIM progression sampler for B200 with the call surface of
``/root/reference/src/pipelines/inference/inference_pipeline_ip.py``: ``_build_labels`` (:184-195),
``_prepare_conditioning`` (:232-308), ``_set_delta_scale_on_processors`` (:311-318), ``_ddim_sample_ip`` (:321-470),
``_latents_to_images`` (:473-486).

What is different underneath:

* one denoising step (time-term select -> UNet -> [second UNet pass + CFG] -> DDIM update) is captured ONCE as a CUDA
  graph and replayed ``sampling_steps`` times; the per-step scalars (sqrt(abar_t) ...) and the hoisted time-embedding
  projections sit in device tables indexed by a device-side step counter, so the loop has no ``.item()`` host syncs;
* the condition-token K/V projections are computed once per call, not once per step;
* the x0/clamp/update arithmetic is one kernel, bit-exact with the reference's fp32 expression order.
"""

from __future__ import annotations

import os
from collections import OrderedDict

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import ops, wcache
from .attention_processor import compute_dtype
from .attention_processor_base import OrdinalIPAttnProcessor2_0
from .attention_processor_routing_gates import SplitInjectionAttentionProcessor


def _build_labels(num_steps: int, start: float = 0.0, end: float = 3.0, device: Optional[torch.device] = None) -> Tensor:
    device = device or torch.device("cpu")
    if num_steps <= 0:
        raise ValueError("`mes_steps` must be a positive integer.")
    return torch.linspace(start, end, steps=num_steps, device=device, dtype=torch.float32)


def _tokens(x: Tensor) -> Tensor:
    return x.unsqueeze(1) if x.dim() == 2 else x


def _prepare_conditioning(module, target_labels: Tensor, source_labels: Tensor, structure_image: Tensor,
                          image_scale: float = 1.0, leace: Optional[dict] = None, zero_aoe: bool = False) -> Tensor:
    """[Source_AOE | E_clean | Delta_AOE] (routing gates) or [AOE | Image] (baseline); reference :232-308."""
    if leace is not None:
        raise NotImplementedError("the legacy LEACE projection is outside the B200 hot path (SURVEY.md section 2, row 19)")
    batch = target_labels.shape[0]
    routing = getattr(module.diff_cfg, "use_routing_gates", True)
    emb = module.ordinal_embedder
    source_aoe = _tokens(emb(source_labels, is_training=False))
    # (the reference expands the one structure image to the batch BEFORE the CLIP tower, :282-283, and encodes 13 identical
    # copies; the tower is per-sample, so encoding once and expanding the tokens gives the same result)
    image_embeds = module._get_image_embeds(structure_image)
    if image_embeds.shape[0] != batch:
        image_embeds = image_embeds.expand(batch, *image_embeds.shape[1:])
    if module.feature_purifier is not None:
        image_embeds = module.feature_purifier(image_embeds, source_aoe)
    if image_scale != 1.0:
        image_embeds = image_embeds * image_scale
    if routing:
        delta = _tokens(emb.get_ordinal_delta_embedding(source_labels, target_labels))
        return torch.cat([source_aoe, image_embeds, delta], dim=1)
    if zero_aoe:
        target_aoe = _tokens(emb.get_negative_embedding(target_labels, is_training=False))
    else:
        target_aoe = _tokens(emb(target_labels, is_training=False))
    return torch.cat([target_aoe, image_embeds], dim=1)


def _set_delta_scale_on_processors(module, delta_scale: float) -> None:
    for _name, mod in module.unet.unet.named_modules():
        if hasattr(mod, "processor") and hasattr(mod.processor, "delta_scale"):
            mod.processor.delta_scale = delta_scale


def ddim_schedule(alphas_cumprod: Tensor, T: int, sampling_steps: int, eta: float) -> Tuple[Tensor, Tensor]:
    """(timesteps int64 [S], coefficient table fp32 [S][8]) on the CPU.  Each entry is formed with the same fp32 tensor
    ops, in the same order, as the reference's per-step scalars (:434-460), so the table is bit-identical to them."""
    ac = alphas_cumprod.detach().to("cpu", torch.float32)
    ts = torch.linspace(T - 1, 0, steps=sampling_steps, dtype=torch.long)
    table = torch.zeros(sampling_steps, 8, dtype=torch.float32)
    for i in range(sampling_steps):
        ab = ac[int(ts[i])]
        table[i, 0] = torch.sqrt(ab)
        table[i, 1] = torch.sqrt(1.0 - ab)
        if i == sampling_steps - 1:
            table[i, 5] = 1.0
            continue
        abp = ac[int(ts[i + 1])]
        table[i, 2] = torch.sqrt(abp)
        if eta == 0.0:
            table[i, 3] = torch.sqrt(1.0 - abp)
        else:
            sigma = eta * torch.sqrt((1 - abp) / (1 - ab) * (1 - ab / abp))
            table[i, 3] = torch.sqrt(1 - abp - sigma ** 2)
            table[i, 4] = sigma
    return ts, table


# cuDNN picks each convolution's algorithm by timing the candidates once per shape (the off-path convolutions are 38 % of a
# step); the flag is scoped to this package's calls.  DADD_CUDNN_BENCHMARK=0 restores cuDNN's heuristics.
CUDNN_BENCHMARK = os.environ.get("DADD_CUDNN_BENCHMARK", "1") != "0"


def _cudnn_flags():
    return torch.backends.cudnn.flags(enabled=True, benchmark=CUDNN_BENCHMARK)


MAX_ENGINES = int(os.environ.get("DADD_MAX_ENGINES", "4"))     # least-recently-used engines beyond this are freed


class ProgressionEngine:
    """Static buffers + one captured step graph for a fixed (batch, latent size, steps, eta, cfg, pathways) signature."""

    def __init__(self, module, batch: int, sampling_steps: int, eta: float, do_cfg: bool, guidance_scale: float,
                 tokens: int, device: torch.device, use_graph: bool = True) -> None:
        self.module = module
        self.batch, self.steps, self.eta, self.do_cfg, self.guidance = batch, sampling_steps, eta, do_cfg, guidance_scale
        self.device = device
        T = module.diff_cfg.num_train_timesteps
        h = module.cfg.dataset.image_size // 8
        c = module.cfg.model.latent_channels
        dim = module.cfg.model.conditioning_dim
        self.x = torch.zeros(batch, c, h, h, device=device, dtype=torch.float32)
        self.ehs = torch.zeros(batch, tokens, dim, device=device, dtype=torch.float32)
        self.ehs_u = torch.zeros_like(self.ehs) if do_cfg else None
        wcache.pin(self.ehs)                     # the K/V projected from these buffers are read by the captured graph:
        if self.ehs_u is not None:               # their cache entries live as long as the buffers
            wcache.pin(self.ehs_u)
        self.state = torch.zeros(2, device=device, dtype=torch.int32)
        self.noise = (torch.zeros(sampling_steps, batch * c * h * h, device=device, dtype=torch.float32)
                      if eta != 0.0 else None)
        self.coef = self.terms_table = None
        self._build_tables()
        self.terms_row = torch.zeros(1, self.terms_table.shape[1], device=device, dtype=torch.float32)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph
        self.launches_per_step = 0
        self.weight_stamp = self._weights_stamp()

    def _build_tables(self) -> None:
        """DDIM coefficient table and the hoisted time-embedding rows (S, sum C_out), refreshed IN PLACE when they exist."""
        m = self.module
        self.timesteps, table = ddim_schedule(m.alphas_cumprod, m.diff_cfg.num_train_timesteps, self.steps, self.eta)
        terms = m.unet.unet.time_terms(self.timesteps.to(self.device)).contiguous()
        if self.coef is None:
            self.coef, self.terms_table = table.to(self.device), terms
        else:
            self.coef.copy_(table)
            self.terms_table.copy_(terms)

    def _weights_stamp(self) -> Tuple:
        """(address, version) of every parameter and buffer the step reads: UNet + processors and the noise schedule."""
        m = self.module
        ts = list(m.unet.parameters()) + list(m.unet.buffers()) + [m.alphas_cumprod]
        return tuple((t.data_ptr(), t._version, t.dtype) for t in ts)

    def _revalidate_weights(self) -> None:
        """``load_state_dict`` / an EMA swap / ``module.to(...)`` after the first capture: the derived 16-bit / fused weights
        (``wcache``) are refreshed only when the Python forward runs and the tables above were computed from the old
        weights, so a replay would mix old and new.  Rebuild the tables in place and drop the graph; the next ``run``
        re-captures after an eager warm-up that refreshes every derived tensor."""
        stamp = self._weights_stamp()
        if stamp != self.weight_stamp:
            self._build_tables()
            self.graph = None
            self.launches_per_step = 0
            self.weight_stamp = stamp

    # one denoising step on the static buffers
    def _step(self) -> None:
        ops.step_begin_(self.state, self.terms_table, self.terms_row, self.steps)
        m = self.module
        with _cudnn_flags():
            eps_c = m(self.x, None, self.ehs, time_terms=self.terms_row)
            eps_u = m(self.x, None, self.ehs_u, time_terms=self.terms_row) if self.do_cfg else None
        ops.ddim_step_table_(self.x, eps_c, eps_u, self.guidance, self.coef, self.state, self.noise, 4.0)

    def _refresh_kv(self) -> None:
        """Everything a replay reads that the Python forward would otherwise refresh: the condition tokens re-projected into
        the (address-stable) per-site K/V caches, and each routing processor's device gate vector
        (dis_gate, anat_gate, delta_scale) - a steer-scale sweep on one engine changes only that vector."""
        for mod in self.module.unet.unet.modules():
            proc = getattr(mod, "processor", None)
            if isinstance(proc, (SplitInjectionAttentionProcessor, OrdinalIPAttnProcessor2_0)):
                proc.project_kv(mod, self.ehs)
                if self.do_cfg:
                    proc.project_kv(mod, self.ehs_u)
                if isinstance(proc, SplitInjectionAttentionProcessor):
                    proc.gate_vector()

    def _capture(self) -> None:
        from . import _lib
        x0 = self.x.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):            # warm-up: cuDNN/cuBLAS plans, weight caches, allocator
            for _ in range(2):
                self.state.zero_()
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            self.state.zero_()
            before = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self._step()
            self.launches_per_step = _lib.launch_count() - before
        self.x.copy_(x0)

    @torch.no_grad()
    def run(self, init_latents: Tensor, embed_cond: Tensor, embed_uncond: Optional[Tensor] = None,
            step_noise: Optional[Tensor] = None) -> Tensor:
        self.x.copy_(init_latents.to(torch.float32).expand_as(self.x))
        self.ehs.copy_(embed_cond)
        if self.do_cfg:
            self.ehs_u.copy_(embed_uncond)
        if self.noise is not None:
            self.noise.zero_()
            if step_noise is not None:
                self.noise[: step_noise.shape[0]].copy_(step_noise.reshape(step_noise.shape[0], -1))
        self._revalidate_weights()
        self._refresh_kv()
        if self.graph is None and (self.use_graph or self.launches_per_step == 0):
            self._capture()
            self._refresh_kv()
        self.state.zero_()
        if self.graph is not None:
            for _ in range(self.steps):
                self.graph.replay()
        else:
            from . import _lib
            before = _lib.launch_count()
            for _ in range(self.steps):
                self._step()
            self.launches_per_step = (_lib.launch_count() - before) // self.steps
        return self.x.clone()


def _engine_for(module, batch: int, sampling_steps: int, eta: float, do_cfg: bool, guidance_scale: float, tokens: int,
                device: torch.device, steer_scale: float, use_graph: bool = True) -> ProgressionEngine:
    cache: "OrderedDict" = module.__dict__.setdefault("_b200_engines", OrderedDict())
    key = (batch, sampling_steps, float(eta), do_cfg, float(guidance_scale) if do_cfg else 0.0, tokens, str(device),
           steer_scale != 0.0, use_graph, module.cfg.dataset.image_size, str(compute_dtype()))
    eng = cache.get(key)
    if eng is None:
        eng = ProgressionEngine(module, batch, sampling_steps, eta, do_cfg, guidance_scale, tokens, device, use_graph)
        cache[key] = eng
        while len(cache) > MAX_ENGINES:          # each engine owns a captured graph and its private memory pool
            cache.popitem(last=False)
    cache.move_to_end(key)
    return eng


@torch.no_grad()
def _ddim_sample_ip(module, target_labels: Tensor, source_labels: Tensor, structure_image: Tensor, sampling_steps: int,
                    device: torch.device, eta: float = 0.0, image_scale: float = 1.0, leace: Optional[dict] = None,
                    steer_scale: float = 0.0, guidance_scale: float = 1.0, *, init_latents: Optional[Tensor] = None,
                    use_graph: bool = True) -> Tensor:
    """Progression sampling (reference :321-470): every MES level starts from the SAME noise tensor.
    ``init_latents`` (1 or B, 4, h, w) optionally injects that noise (tests; the reference draws it on ``device``)."""
    device = torch.device(device)
    routing = getattr(module.diff_cfg, "use_routing_gates", True)
    do_cfg = (not routing) and (guidance_scale != 1.0)
    n = target_labels.shape[0]
    height = module.cfg.dataset.image_size
    T = module.diff_cfg.num_train_timesteps
    if sampling_steps > T:
        raise ValueError(f"sampling_steps={sampling_steps} must be <= num_train_timesteps={T}")
    if init_latents is None:
        init_latents = torch.randn(1, module.cfg.model.latent_channels, height // 8, height // 8, device=device,
                                   dtype=torch.float32)
    latents = init_latents.to(device).repeat(n, 1, 1, 1) if init_latents.shape[0] == 1 else init_latents.to(device)
    return _sample(module, target_labels, source_labels, structure_image, latents, sampling_steps, device, eta,
                   image_scale, leace, steer_scale, guidance_scale, do_cfg, use_graph)


def _sample(module, target_labels, source_labels, structure_images, latents, sampling_steps, device, eta, image_scale,
            leace, steer_scale, guidance_scale, do_cfg, use_graph) -> Tensor:
    target_labels = target_labels.to(device)
    source_labels = source_labels.to(device)
    structure_images = structure_images.to(device)
    n = target_labels.shape[0]
    embed_cond = _prepare_conditioning(module, target_labels, source_labels, structure_images, image_scale=image_scale, leace=leace)
    embed_uncond = None
    if do_cfg:
        embed_uncond = _prepare_conditioning(module, target_labels, source_labels, structure_images,
                                             image_scale=image_scale, leace=leace, zero_aoe=True)
    _set_delta_scale_on_processors(module, steer_scale)
    step_noise = None
    if eta != 0.0 and sampling_steps > 1:      # same RNG order as the reference's per-step randn_like (:463)
        step_noise = torch.stack([torch.randn_like(latents) for _ in range(sampling_steps - 1)])
    eng = _engine_for(module, n, sampling_steps, eta, do_cfg, guidance_scale, embed_cond.shape[1], device, steer_scale, use_graph)
    return eng.run(latents, embed_cond, embed_uncond, step_noise)


@torch.no_grad()
def sample_progressions(module, image_tokens: Tensor, source_labels: Tensor, mes_steps: int = 13, sampling_steps: int = 50,
                        device: Optional[torch.device] = None, eta: float = 0.0, image_scale: float = 1.0,
                        steer_scale: float = 3.0, guidance_scale: float = 1.0, init_latents: Optional[Tensor] = None,
                        decode: bool = True, use_graph: bool = True) -> Tensor:
    """P patient progressions in one batch: what ``main()`` of the reference does for one patient (:596-640: labels
    ``linspace(0, K-1, mes_steps)``, one shared noise tensor per patient, ``_ddim_sample_ip``, ``_latents_to_images``),
    for ``P = image_tokens.shape[0]`` patients at once.  ``image_tokens``: CLIP-preprocessed structure images (P,3,224,224) -
    encoded ONCE per patient by the module's CLIP + resampler front end - or already projected tokens (P,16,768);
    ``source_labels`` (P,) / ``init_latents`` (P,4,h,w); all may live on the host (pinned or not): the copies are part of the call.
    Returns images (P*mes_steps, 3, H, W) fp32 in [0,1] on ``device`` (or the final latents when ``decode=False``)."""
    device = torch.device(device if device is not None else image_tokens.device)
    p = image_tokens.shape[0]
    routing = getattr(module.diff_cfg, "use_routing_gates", True)
    do_cfg = (not routing) and (guidance_scale != 1.0)
    k = module.cfg.dataset.num_classes if hasattr(module.cfg.dataset, "num_classes") else 4
    h = module.cfg.dataset.image_size // 8
    tokens = image_tokens.to(device, non_blocking=True)
    if tokens.dim() == 4:
        tokens = module._get_image_embeds(tokens)
    tokens = tokens.repeat_interleave(mes_steps, dim=0)
    source = source_labels.to(device, non_blocking=True).to(torch.float32).repeat_interleave(mes_steps)
    target = _build_labels(mes_steps, 0.0, float(k - 1), device).repeat(p)
    if init_latents is None:
        init_latents = torch.randn(p, module.cfg.model.latent_channels, h, h, device=device, dtype=torch.float32)
    latents = init_latents.to(device, non_blocking=True).repeat_interleave(mes_steps, dim=0)
    latents = _sample(module, target, source, tokens, latents, sampling_steps, device, eta, image_scale, None, steer_scale,
                      guidance_scale, do_cfg, use_graph)
    return _latents_to_images(module, latents) if decode else latents


@torch.no_grad()
def _latents_to_images(module, latents: Tensor) -> Tensor:
    """vae.decode(latents / latent_scale) -> clamp(-1,1) -> [0,1] (reference :473-486); fp32 (B,3,H,W) on the device."""
    with _cudnn_flags():
        decoded = module.vae.decode(latents / module.diff_cfg.latent_scale)
    images = decoded.sample if hasattr(decoded, "sample") else decoded
    return ops.image_post(images).contiguous()
