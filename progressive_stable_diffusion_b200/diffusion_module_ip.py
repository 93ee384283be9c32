"""``DiffusionModuleWithIP`` for B200 - the module surface the reference pipelines use
(``/root/reference/src/models/diffusion_module_ip.py``: ``DiffusionIPConfig`` :33-64, ``__init__`` :81-201,
``_setup_attention_processors`` :203-233, ``_build_noise_schedule`` :274-287, ``_get_image_embeds`` :315-332,
``forward`` :383-390), as a plain ``nn.Module`` (no Lightning).  State-dict keys equal a Lightning checkpoint's
(``unet.unet.*``, ``vae.vae.*``, ``ordinal_embedder.*``, ``feature_purifier.*``), schedule buffers are non-persistent.

The CLIP + resampler front end (``image_encoder.py``) is built on request (``build_image_encoder=True``
or a checkpoint that carries its weights): it is 304 M frozen parameters that the synthetic benchmark - which feeds
``(B, 16, 768)`` image tokens straight into ``_get_image_embeds``, as SURVEY.md 8d allows - does not need.
Out of this tier: ``training_step`` / optimisers (f1).
"""

from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .attention_processor_base import set_ordinal_ip_attention_processors
from .attention_processor_routing_gates import SplitInjectionAttentionProcessor, set_split_injection_processors
from .feature_purifier import FeaturePurifier
from .image_encoder import ImageEncoder, ImageProjection, ImageProjectionPlus
from .ordinal_embedder import AdditiveOrdinalEmbedder
from .unet import OrdinalUNet, UNetConfig
from .vae import SDVAE


@dataclass
class DiffusionIPConfig:
    num_train_timesteps: int
    beta_start: float
    beta_end: float
    noise_schedule: str = "linear"
    sampling_steps: int = 50
    guidance_scale: float = 2.0
    min_snr_gamma: float = 1.0
    ema_update_interval: int = 10
    latent_scale: float = 0.18215
    input_perturbation: float = 0.0
    image_encoder_path: str = "openai/clip-vit-base-patch16"
    num_image_tokens: int = 16
    num_aoe_tokens: int = 16
    use_frequency_strategy: bool = True
    use_image_projection_plus: bool = False
    use_feature_purifier: bool = True
    purifier_num_heads: int = 8
    purifier_ff_mult: int = 2
    delta_scale: float = 0.0
    use_routing_gates: bool = True
    gate_init_anatomy: tuple = (0.5, 0.5)
    gate_init_disease: tuple = (0.5, 0.5)


class AttrDict(dict):
    """Minimal stand-in for an OmegaConf ``DictConfig`` (omegaconf is not in this image): attribute + item access."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    __setattr__ = dict.__setitem__

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return [AttrDict.wrap(v) for v in obj]
        return obj


def load_config(path) -> AttrDict:
    """Read one of the reference's Hydra/OmegaConf YAML files (configs/train_ip.yaml, configs/evaluation_configs/*.yaml)."""
    import yaml

    with open(path) as f:
        return AttrDict.wrap(yaml.safe_load(f))


def default_config(**model_overrides) -> AttrDict:
    """The keys of the reference's ``configs/train_ip.yaml`` that the hot path reads (:7-44, :100-108), with the gate
    preset of ``configs/evaluation_configs/uqqx9kg9_all.yaml`` (anatomy (0.1, 0.9), disease (0.9, 0.1))."""
    cfg = {
        "model": {
            "embedding_dim": 768, "conditioning_dim": 768, "latent_channels": 4,
            "pretrained_vae_path": "CompVis/stable-diffusion-v1-4", "pretrained_unet_path": "CompVis/stable-diffusion-v1-4",
            "image_encoder_path": "openai/clip-vit-large-patch14", "num_image_tokens": 16, "num_aoe_tokens": 16,
            "use_image_projection_plus": True, "use_frequency_strategy": True, "use_routing_gates": True,
            "use_feature_purifier": True, "gate_init_anatomy": [0.1, 0.9], "gate_init_disease": [0.9, 0.1],
            "purifier_num_heads": 8, "purifier_ff_mult": 2, "delta_scale": 0.0, "cfg_drop_prob": 0.0,
            "ordinal_embedder": {"type": "aoe", "num_classes": 4, "aoe": {"delta_scale": 0.05}},
        },
        "dataset": {"image_size": 256, "num_classes": 4},
        "training": {"precision": "16-mixed", "use_min_snr_weighting": True},
        "diffusion": {"noise_schedule": "linear", "beta_start": 0.00085, "beta_end": 0.012, "num_train_timesteps": 1000,
                      "sampling_steps": 50, "guidance_scale": 1.0, "min_snr_gamma": 1.0, "ema_update_interval": 1},
    }
    cfg["model"].update(model_overrides)
    return AttrDict.wrap(cfg)


class DiffusionModuleWithIP(nn.Module):
    def __init__(self, cfg: Any, build_vae: bool = True, build_image_encoder: bool = False, build_vae_encoder: bool = False) -> None:
        super().__init__()
        self.cfg = cfg
        m, d = cfg.model, cfg.diffusion
        self.diff_cfg = DiffusionIPConfig(
            num_train_timesteps=d.num_train_timesteps, beta_start=d.beta_start, beta_end=d.beta_end,
            noise_schedule=d.noise_schedule, sampling_steps=d.sampling_steps, guidance_scale=d.guidance_scale,
            min_snr_gamma=d.min_snr_gamma, ema_update_interval=d.ema_update_interval,
            latent_scale=getattr(d, "latent_scale", 0.18215),
            input_perturbation=getattr(getattr(cfg, "training", SimpleNamespace()), "input_perturbation", 0.0),
            image_encoder_path=getattr(m, "image_encoder_path", "openai/clip-vit-base-patch16"),
            num_image_tokens=getattr(m, "num_image_tokens", 16),
            use_image_projection_plus=getattr(m, "use_image_projection_plus", False),
            num_aoe_tokens=getattr(m, "num_aoe_tokens", 16),
            use_frequency_strategy=getattr(m, "use_frequency_strategy", True),
            use_feature_purifier=getattr(m, "use_feature_purifier", True),
            purifier_num_heads=getattr(m, "purifier_num_heads", 8),
            purifier_ff_mult=getattr(m, "purifier_ff_mult", 2),
            delta_scale=getattr(m, "delta_scale", 0.0),
            use_routing_gates=getattr(m, "use_routing_gates", True),
            gate_init_anatomy=tuple(getattr(m, "gate_init_anatomy", [0.5, 0.5])),
            gate_init_disease=tuple(getattr(m, "gate_init_disease", [0.5, 0.5])),
        )
        self.vae = SDVAE(getattr(m, "pretrained_vae_path", None), build_encoder=build_vae_encoder) if build_vae else None
        self.image_encoder = None        # frozen CLIP tower + trainable projection (reference :130-149)
        self.image_projection = None
        if build_image_encoder:
            self.image_encoder = ImageEncoder(pretrained_path=self.diff_cfg.image_encoder_path, torch_dtype=torch.float32)
            if self.diff_cfg.use_image_projection_plus:
                self.image_projection = ImageProjectionPlus(clip_hidden_dim=self.image_encoder.hidden_size,
                                                            cross_attention_dim=m.conditioning_dim,
                                                            num_tokens=self.diff_cfg.num_image_tokens)
            else:
                self.image_projection = ImageProjection(clip_embedding_dim=self.image_encoder.projection_dim,
                                                        cross_attention_dim=m.conditioning_dim,
                                                        num_tokens=self.diff_cfg.num_image_tokens)
        emb = m.ordinal_embedder
        self.ordinal_embedder = AdditiveOrdinalEmbedder(
            num_classes=emb.num_classes, embedding_dim=m.embedding_dim,
            delta_scale=getattr(emb.aoe, "delta_scale", 0.1), num_tokens=getattr(m, "num_aoe_tokens", 16))
        self.unet = OrdinalUNet(UNetConfig(pretrained_unet_path=m.pretrained_unet_path, conditioning_dim=m.conditioning_dim,
                                           in_channels=m.latent_channels, out_channels=m.latent_channels))
        self.feature_purifier = (FeaturePurifier(dim=m.conditioning_dim, num_heads=self.diff_cfg.purifier_num_heads,
                                                 ff_mult=self.diff_cfg.purifier_ff_mult)
                                 if self.diff_cfg.use_feature_purifier else None)
        self._setup_attention_processors()
        betas, alphas_cumprod = self._build_noise_schedule()
        self.register_buffer("betas", betas, persistent=False)
        self.register_buffer("alphas_cumprod", alphas_cumprod, persistent=False)
        self.register_buffer("alphas_cumprod_prev",
                             torch.cat([torch.ones(1, dtype=alphas_cumprod.dtype), alphas_cumprod[:-1]], dim=0), persistent=False)
        self.register_buffer("snr_values", alphas_cumprod / (1.0 - alphas_cumprod + 1e-8), persistent=False)

    # ------------------------------------------------------------------ reference :203-233
    def _setup_attention_processors(self) -> None:
        unet = self.unet.unet
        c = self.diff_cfg
        if c.use_routing_gates:
            set_split_injection_processors(
                unet=unet, num_image_tokens=c.num_image_tokens, num_aoe_tokens=c.num_aoe_tokens,
                num_delta_tokens=c.num_aoe_tokens, use_frequency_strategy=c.use_frequency_strategy,
                delta_scale=c.delta_scale,
                gate_inits={"anatomy": c.gate_init_anatomy, "disease": c.gate_init_disease, "both": (0.5, 0.5)})
        else:
            set_ordinal_ip_attention_processors(unet=unet, num_image_tokens=c.num_image_tokens,
                                                num_aoe_tokens=c.num_aoe_tokens,
                                                use_frequency_strategy=c.use_frequency_strategy)

    # ------------------------------------------------------------------ reference :274-287
    def _build_noise_schedule(self) -> Tuple[Tensor, Tensor]:
        if self.diff_cfg.noise_schedule != "linear":
            raise NotImplementedError("Only linear noise schedule is supported.")
        betas = torch.linspace(self.diff_cfg.beta_start, self.diff_cfg.beta_end, self.diff_cfg.num_train_timesteps,
                               dtype=torch.float32)
        return betas, torch.cumprod(1.0 - betas, dim=0)

    # ------------------------------------------------------------------ reference :315-332
    def _get_image_embeds(self, structure_images: Tensor) -> Tensor:
        """(B, num_image_tokens, conditioning_dim) anatomy tokens from CLIP-preprocessed pixels (B, 3, 224, 224); tokens that
        are already projected ``(B, num_image_tokens, conditioning_dim)`` pass through (synthetic benchmarks, cached tokens)."""
        if structure_images.dim() == 3 and structure_images.shape[-1] == self.cfg.model.conditioning_dim:
            return structure_images
        if self.image_encoder is None:
            raise RuntimeError("this module was built without the CLIP front end: construct it with build_image_encoder=True "
                               "(or load a checkpoint that carries image_encoder.* weights), or pass projected (B, 16, 768) tokens")
        if self.diff_cfg.use_image_projection_plus:
            image_embeds = self.image_encoder.get_hidden_states(structure_images)
        else:
            image_embeds = self.image_encoder(structure_images)
        return self.image_projection(image_embeds)

    def forward(self, latents: Tensor, timesteps: Tensor, cond_embed: Tensor, time_terms: Optional[Tensor] = None) -> Tensor:
        return self.unet(latents, timesteps, cond_embed, time_terms=time_terms)

    # ------------------------------------------------------------------ reference :464-478
    def _collect_gate_values(self) -> Dict[str, Dict[str, float]]:
        out: Dict[str, Dict[str, float]] = {}
        for name, mod in self.unet.unet.named_modules():
            proc = getattr(mod, "processor", None)
            if isinstance(proc, SplitInjectionAttentionProcessor):
                out[name] = {"anat_gate": proc.anat_gate.item(), "dis_gate": proc.dis_gate.item()}
        return out

    # ------------------------------------------------------------------ reference :289-313 (host-side pieces of the training path)
    @property
    def device(self) -> torch.device:
        return self.alphas_cumprod.device

    def _sample_timesteps(self, batch_size: int) -> Tensor:
        return torch.randint(0, self.diff_cfg.num_train_timesteps, (batch_size,), device=self.device, dtype=torch.long)

    def _q_sample(self, x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
        """Forward diffusion x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) noise (reference :299-303)."""
        sqrt_ab = torch.sqrt(self.alphas_cumprod[t]).view(-1, 1, 1, 1)
        sqrt_1m = torch.sqrt(1.0 - self.alphas_cumprod[t]).view(-1, 1, 1, 1)
        return sqrt_ab * x0 + sqrt_1m * noise

    def _min_snr_weight(self, t: Tensor) -> Tensor:
        """Min-SNR-gamma loss weight min(snr, gamma) / (snr + 1e-8) (reference :305-313)."""
        if not getattr(getattr(self.cfg, "training", SimpleNamespace()), "use_min_snr_weighting", True):
            return torch.ones_like(t, dtype=torch.float32, device=self.device)
        snr = self.snr_values[t]
        return torch.minimum(snr, torch.tensor(self.diff_cfg.min_snr_gamma, device=snr.device)) / (snr + 1e-8)

    def training_step(self, *args, **kwargs):
        raise NotImplementedError("training (backward kernels + DDP) is the next tier (SURVEY.md 8f, row f1)")

    @classmethod
    def load_from_checkpoint(cls, path, cfg: Any = None, weights_only: bool = False, strict: bool = False,
                             map_location="cpu", **kwargs) -> "DiffusionModuleWithIP":
        """Lightning-compatible loader: reads ``ckpt['state_dict']`` (EMA weights when the EMA callback saved them,
        ema_callback.py:316-329) into a module built from ``cfg``."""
        ckpt = torch.load(path, map_location=map_location, weights_only=weights_only)
        state = ckpt.get("state_dict", ckpt)
        kwargs.setdefault("build_image_encoder", any(k.startswith("image_encoder.") for k in state))
        module = cls(cfg if cfg is not None else default_config(), **kwargs)
        module.load_state_dict(state, strict=strict)
        return module
