"""Seeded random-init state dicts with the reference's key layout (TEST INFRASTRUCTURE).

Keys follow a Lightning checkpoint of ``DiffusionModuleWithIP``
(``/root/reference/src/models/diffusion_module_ip.py:124-176``): ``unet.unet.*`` is the
diffusers SD-1.x ``UNet2DConditionModel`` layout plus the DADD
processor tensors ``...attn2.processor.{to_k_dis.weight,to_v_dis.weight,anat_gate,dis_gate}``
(``attention_processor_routing_gates.py:73-82``); ``vae.vae.*`` the AutoencoderKL decoder;
``ordinal_embedder.*`` (``ordinal_embedder.py:70-88``); ``feature_purifier.*``
(``feature_purifier.py:46-62``).

Initialisation mirrors PyTorch defaults (``kaiming_uniform_(a=sqrt(5))`` -> U(-1/sqrt(fan_in),
1/sqrt(fan_in)) for weights and biases) drawn from one ``torch.Generator`` so the same
seed gives the same tensors in the build container and on the GPU box (same torch build).
``gain`` multiplies every weight bound (gain=sqrt(3) keeps activations at unit variance, which
makes whole-network parity tests sensitive to the attention blocks); ``affine_jitter`` perturbs
norm weights/biases away from (1, 0) so channel indexing errors cannot hide.
"""

from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Tuple

import torch

State = Dict[str, torch.Tensor]

BLOCK_OUT = (320, 640, 1280, 1280)
CROSS_DIM = 768
TIME_DIM = 1280
HEADS = 8


class _Init:
    def __init__(self, seed: int, gain: float, affine_jitter: float) -> None:
        self.g = torch.Generator().manual_seed(seed)
        self.gain = gain
        self.jit = affine_jitter
        self.sd: State = {}

    def _u(self, shape, bound):
        return (torch.rand(shape, generator=self.g, dtype=torch.float32) * 2.0 - 1.0) * bound

    def linear(self, name: str, cin: int, cout: int, bias: bool = True) -> None:
        b = 1.0 / math.sqrt(cin)
        self.sd[name + ".weight"] = self._u((cout, cin), b * self.gain)
        if bias:
            self.sd[name + ".bias"] = self._u((cout,), b)

    def conv(self, name: str, cin: int, cout: int, k: int) -> None:
        b = 1.0 / math.sqrt(cin * k * k)
        self.sd[name + ".weight"] = self._u((cout, cin, k, k), b * self.gain)
        self.sd[name + ".bias"] = self._u((cout,), b)

    def norm(self, name: str, c: int) -> None:
        w = torch.ones(c)
        b = torch.zeros(c)
        if self.jit > 0:
            w = w + self.jit * torch.randn(c, generator=self.g)
            b = b + self.jit * torch.randn(c, generator=self.g)
        self.sd[name + ".weight"] = w
        self.sd[name + ".bias"] = b


def role_of(block_name: str) -> str:
    """Restates ``get_block_type`` (attention_processor_routing_gates.py:199-230)."""
    if "mid_block" in block_name:
        return "disease"
    if "down_blocks" in block_name:
        idx = int(block_name.split("down_blocks.")[1].split(".")[0])
        return "disease" if idx >= 2 else "anatomy"
    if "up_blocks" in block_name:
        idx = int(block_name.split("up_blocks.")[1].split(".")[0])
        return "disease" if idx <= 1 else "anatomy"
    return "both"


def attention_sites() -> Iterable[Tuple[str, int]]:
    """(prefix, C) of the 16 Transformer2DModel sites in diffusers' registration order."""
    for i in range(3):
        for j in range(2):
            yield f"down_blocks.{i}.attentions.{j}", BLOCK_OUT[i]
    yield "mid_block.attentions.0", 1280
    rev = list(reversed(BLOCK_OUT))
    for i in range(1, 4):
        for j in range(3):
            yield f"up_blocks.{i}.attentions.{j}", rev[i]


def _resnet(it: _Init, p: str, cin: int, cout: int, temb: Optional[int]) -> None:
    it.norm(p + ".norm1", cin)
    it.conv(p + ".conv1", cin, cout, 3)
    if temb is not None:
        it.linear(p + ".time_emb_proj", temb, cout)
    it.norm(p + ".norm2", cout)
    it.conv(p + ".conv2", cout, cout, 3)
    if cin != cout:
        it.conv(p + ".conv_shortcut", cin, cout, 1)


def _transformer(it: _Init, p: str, c: int, gates: Tuple[float, float], routing: bool) -> None:
    it.norm(p + ".norm", c)
    it.conv(p + ".proj_in", c, c, 1)
    t = p + ".transformer_blocks.0"
    it.norm(t + ".norm1", c)
    for n in ("to_q", "to_k", "to_v"):
        it.linear(f"{t}.attn1.{n}", c, c, bias=False)
    it.linear(t + ".attn1.to_out.0", c, c)
    it.norm(t + ".norm2", c)
    it.linear(t + ".attn2.to_q", c, c, bias=False)
    it.linear(t + ".attn2.to_k", CROSS_DIM, c, bias=False)
    it.linear(t + ".attn2.to_v", CROSS_DIM, c, bias=False)
    it.linear(t + ".attn2.to_out.0", c, c)
    if routing:
        # independent draws (a trained checkpoint has to_k_dis != to_k); the warm-start copy
        # (attention_processor_routing_gates.py:308-314) is tested separately (invariant I5)
        it.linear(t + ".attn2.processor.to_k_dis", CROSS_DIM, c, bias=False)
        it.linear(t + ".attn2.processor.to_v_dis", CROSS_DIM, c, bias=False)
        it.sd[t + ".attn2.processor.anat_gate"] = torch.tensor(float(gates[0]))
        it.sd[t + ".attn2.processor.dis_gate"] = torch.tensor(float(gates[1]))
    it.norm(t + ".norm3", c)
    it.linear(t + ".ff.net.0.proj", c, 8 * c)
    it.linear(t + ".ff.net.2", 4 * c, c)
    it.conv(p + ".proj_out", c, c, 1)


def make_unet_state(
    seed: int = 0,
    gain: float = 1.0,
    affine_jitter: float = 0.1,
    routing: bool = True,
    gate_inits: Optional[Dict[str, Tuple[float, float]]] = None,
    prefix: str = "",
) -> State:
    """SD-1.x UNet + DADD processor tensors.  Default gates = configs/evaluation_configs/uqqx9kg9_all.yaml:31-32."""
    if gate_inits is None:
        gate_inits = {"anatomy": (0.1, 0.9), "disease": (0.9, 0.1), "both": (0.5, 0.5)}
    it = _Init(seed, gain, affine_jitter)
    it.conv("conv_in", 4, 320, 3)
    it.linear("time_embedding.linear_1", 320, TIME_DIM)
    it.linear("time_embedding.linear_2", TIME_DIM, TIME_DIM)
    cin = 320
    for i, cout in enumerate(BLOCK_OUT):
        for j in range(2):
            _resnet(it, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, TIME_DIM)
            if i < 3:
                p = f"down_blocks.{i}.attentions.{j}"
                _transformer(it, p, cout, gate_inits[role_of(p)], routing)
        if i < 3:
            it.conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        cin = cout
    _resnet(it, "mid_block.resnets.0", 1280, 1280, TIME_DIM)
    _transformer(it, "mid_block.attentions.0", 1280, gate_inits[role_of("mid_block")], routing)
    _resnet(it, "mid_block.resnets.1", 1280, 1280, TIME_DIM)
    up_in = {0: (2560, 2560, 2560), 1: (2560, 2560, 1920), 2: (1920, 1280, 960), 3: (960, 640, 640)}
    rev = list(reversed(BLOCK_OUT))
    for i in range(4):
        for j in range(3):
            _resnet(it, f"up_blocks.{i}.resnets.{j}", up_in[i][j], rev[i], TIME_DIM)
            if i > 0:
                p = f"up_blocks.{i}.attentions.{j}"
                _transformer(it, p, rev[i], gate_inits[role_of(p)], routing)
        if i < 3:
            it.conv(f"up_blocks.{i}.upsamplers.0.conv", rev[i], rev[i], 3)
    it.norm("conv_norm_out", 320)
    it.conv("conv_out", 320, 4, 3)
    return {prefix + k: v for k, v in it.sd.items()}


def make_vae_decoder_state(seed: int = 1, gain: float = 1.0, affine_jitter: float = 0.1, prefix: str = "") -> State:
    """SD AutoencoderKL decoder half (+ post_quant_conv); SURVEY.md Appendix A.6."""
    it = _Init(seed, gain, affine_jitter)
    it.conv("post_quant_conv", 4, 4, 1)
    it.conv("decoder.conv_in", 4, 512, 3)
    _resnet(it, "decoder.mid_block.resnets.0", 512, 512, None)
    a = "decoder.mid_block.attentions.0"
    it.norm(a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        it.linear(f"{a}.{n}", 512, 512)
    _resnet(it, "decoder.mid_block.resnets.1", 512, 512, None)
    chans = (512, 512, 256, 128)
    cin = 512
    for i, cout in enumerate(chans):
        for j in range(3):
            _resnet(it, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i < 3:
            it.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
        cin = cout
    it.norm("decoder.conv_norm_out", 128)
    it.conv("decoder.conv_out", 128, 3, 3)
    return {prefix + k: v for k, v in it.sd.items()}


def make_vae_encoder_state(seed: int = 7, gain: float = 1.0, affine_jitter: float = 0.1, prefix: str = "") -> State:
    """SD AutoencoderKL encoder half (+ quant_conv): the frozen first step of the reference's training_step
    (diffusion_module_ip.py:419-420 ``vae.encode(images).latent_dist.sample()``)."""
    it = _Init(seed, gain, affine_jitter)
    it.conv("encoder.conv_in", 3, 128, 3)
    chans = (128, 256, 512, 512)
    cin = 128
    for i, cout in enumerate(chans):
        for j in range(2):
            _resnet(it, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout, None)
        if i < 3:
            it.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
        cin = cout
    _resnet(it, "encoder.mid_block.resnets.0", 512, 512, None)
    a = "encoder.mid_block.attentions.0"
    it.norm(a + ".group_norm", 512)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        it.linear(f"{a}.{n}", 512, 512)
    _resnet(it, "encoder.mid_block.resnets.1", 512, 512, None)
    it.norm("encoder.conv_norm_out", 512)
    it.conv("encoder.conv_out", 512, 8, 3)
    it.conv("quant_conv", 8, 8, 1)
    return {prefix + k: v for k, v in it.sd.items()}


def make_purifier_state(seed: int = 2, dim: int = 768, ff_mult: int = 2, affine_jitter: float = 0.1,
                        prefix: str = "") -> State:
    """Keys of ``FeaturePurifier`` (feature_purifier.py:46-62): nn.MultiheadAttention packs q/k/v."""
    it = _Init(seed, 1.0, affine_jitter)
    it.norm("norm_img", dim)
    it.norm("norm_aoe", dim)
    it.sd["cross_attn.in_proj_weight"] = it._u((3 * dim, dim), math.sqrt(6.0 / (4 * dim)))  # xavier_uniform
    it.sd["cross_attn.in_proj_bias"] = 0.02 * torch.randn(3 * dim, generator=it.g)
    it.linear("cross_attn.out_proj", dim, dim)
    it.linear("gate.0", 2 * dim, dim * ff_mult)
    it.linear("gate.2", dim * ff_mult, dim)
    it.norm("norm_out", dim)
    return {prefix + k: v for k, v in it.sd.items()}


def make_aoe_state(seed: int = 3, num_classes: int = 4, dim: int = 768, num_tokens: int = 16,
                   init_std: float = 0.02, delta_scale: float = 0.05, prefix: str = "") -> State:
    """Keys of ``AdditiveOrdinalEmbedder`` (ordinal_embedder.py:70-88,90-105)."""
    it = _Init(seed, 1.0, 0.0)
    it.sd["base"] = init_std * torch.randn(dim, generator=it.g)
    d = delta_scale + init_std * torch.randn(num_classes - 1, dim, generator=it.g)
    for i in range(num_classes - 1):
        d[i] *= 1.0 + 0.1 * i
    it.sd["deltas"] = d
    it.linear("projector.0", dim, 2 * dim)
    it.linear("projector.2", 2 * dim, dim * num_tokens)
    it.norm("norm", dim * num_tokens)          # constructed at :85, never applied in any forward
    it.sd["null_embedding"] = torch.zeros(1, dim)
    return {prefix + k: v for k, v in it.sd.items()}


def make_module_state(seed: int = 0, gain: float = 1.0, routing: bool = True,
                      gate_inits: Optional[Dict[str, Tuple[float, float]]] = None, with_vae: bool = True) -> State:
    """Whole-module state dict with the Lightning checkpoint prefixes."""
    sd: State = {}
    sd.update(make_unet_state(seed, gain, routing=routing, gate_inits=gate_inits, prefix="unet.unet."))
    if with_vae:
        sd.update(make_vae_decoder_state(seed + 1, prefix="vae.vae."))
    sd.update(make_purifier_state(seed + 2, prefix="feature_purifier."))
    sd.update(make_aoe_state(seed + 3, prefix="ordinal_embedder."))
    return sd


def make_clip_vision_state(seed: int = 5, hidden: int = 1024, inter: int = 4096, layers: int = 24, heads: int = 16,
                           image: int = 224, patch: int = 14, proj: int = 768, affine_jitter: float = 0.1) -> State:
    """Keys of transformers' ``CLIPVisionModelWithProjection.state_dict()`` (the tower the reference loads at
    image_encoder.py:34-38).  PyTorch-default-style bounds with gain sqrt(3) so 24 residual layers keep the signal alive."""
    it = _Init(seed, math.sqrt(3.0), affine_jitter)
    p = "vision_model."
    it.sd[p + "embeddings.class_embedding"] = torch.randn(hidden, generator=it.g)
    b = 1.0 / math.sqrt(3 * patch * patch)
    it.sd[p + "embeddings.patch_embedding.weight"] = it._u((hidden, 3, patch, patch), b * it.gain)
    it.sd[p + "embeddings.position_embedding.weight"] = 0.5 * torch.randn((image // patch) ** 2 + 1, hidden, generator=it.g)
    it.norm(p + "pre_layrnorm", hidden)
    for i in range(layers):
        q = f"{p}encoder.layers.{i}."
        for t in ("k_proj", "v_proj", "q_proj", "out_proj"):
            it.linear(q + "self_attn." + t, hidden, hidden)
        it.norm(q + "layer_norm1", hidden)
        it.linear(q + "mlp.fc1", hidden, inter)
        it.linear(q + "mlp.fc2", inter, hidden)
        it.norm(q + "layer_norm2", hidden)
    it.norm(p + "post_layernorm", hidden)
    it.linear("visual_projection", hidden, proj, bias=False)
    return it.sd


def make_projection_plus_state(seed: int = 6, clip_hidden: int = 1024, dim: int = 768, num_tokens: int = 16, depth: int = 2,
                               affine_jitter: float = 0.1) -> State:
    """Keys of the reference's ``ImageProjectionPlus`` (image_encoder.py:143-191)."""
    it = _Init(seed, math.sqrt(3.0), affine_jitter)
    it.sd["latents"] = torch.randn(1, num_tokens, dim, generator=it.g) * 0.5
    if clip_hidden != dim:
        it.linear("proj_in", clip_hidden, dim)
    for i in range(depth):
        q = f"layers.{i}."
        b = 1.0 / math.sqrt(dim)
        it.sd[q + "cross_attn.in_proj_weight"] = it._u((3 * dim, dim), b * it.gain)
        it.sd[q + "cross_attn.in_proj_bias"] = it._u((3 * dim,), b)
        it.linear(q + "cross_attn.out_proj", dim, dim)
        it.linear(q + "ff.0", dim, 4 * dim)
        it.linear(q + "ff.2", 4 * dim, dim)
        it.norm(q + "norm1", dim)
        it.norm(q + "norm2", dim)
    it.norm("norm_out", dim)
    return it.sd


def sub_state(sd: State, prefix: str) -> State:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}
