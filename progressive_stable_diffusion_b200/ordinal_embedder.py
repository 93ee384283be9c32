"""Additive Ordinal Embedder on B200: the module surface of ``/root/reference/src/models/ordinal_embedder.py``
(``AdditiveOrdinalEmbedder`` :43-309: parameters ``base``, ``deltas``, ``projector.{0,2}``, ``norm``, ``null_embedding``;
methods ``forward``, ``get_negative_embedding``, ``get_disease_delta_embedding``, ``get_ordinal_delta_embedding``).

The class table E[k] = base + cumsum(deltas), the clamp / floor / ceil gather and the linear interpolation are one kernel
(``dadd_aoe_interp_fwd``, integer indices exact); the 768 -> 1536 -> 16*768 projector runs on cuBLAS in fp32.
Runs once per sampling call.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, wcache


class AdditiveOrdinalEmbedder(nn.Module):
    def __init__(self, num_classes: int, embedding_dim: int, init_std: float = 0.02, delta_scale: float = 0.1,
                 learnable_null: bool = True, num_tokens: int = 16) -> None:
        super().__init__()
        if num_classes < 2:
            raise ValueError("num_classes must be ≥ 2 for ordinal modeling.")
        self.num_classes = num_classes
        self.embedding_dim = embedding_dim
        self.num_tokens = num_tokens
        self.base = nn.Parameter(torch.zeros(embedding_dim))
        nn.init.normal_(self.base, mean=0.0, std=init_std)
        self.deltas = nn.Parameter(torch.empty(num_classes - 1, embedding_dim))
        with torch.no_grad():                      # monotone init: positive-mean deltas, growing with the class index
            for i in range(num_classes - 1):
                nn.init.normal_(self.deltas[i], mean=delta_scale, std=init_std)
                self.deltas[i] *= 1.0 + 0.1 * i
        self.projector = nn.Sequential(nn.Linear(embedding_dim, embedding_dim * 2), nn.GELU(),
                                       nn.Linear(embedding_dim * 2, embedding_dim * num_tokens))
        self.norm = nn.LayerNorm(embedding_dim * num_tokens)   # present in checkpoints; the reference never applies it (:85)
        if learnable_null:
            self.null_embedding = nn.Parameter(torch.zeros(1, embedding_dim))
        else:
            self.register_buffer("null_embedding", torch.zeros(1, embedding_dim))

    # ------------------------------------------------------------------
    def _interp(self, labels: torch.Tensor) -> torch.Tensor:
        base = wcache.cast(self, "base", self.base, torch.float32)
        deltas = wcache.cast(self, "deltas", self.deltas, torch.float32)
        return ops.aoe_interp(base, deltas, labels.to(device=base.device, dtype=torch.float32).reshape(-1).contiguous())

    def _project(self, emb: torch.Tensor) -> torch.Tensor:
        p0, p2 = self.projector[0], self.projector[2]
        f32 = torch.float32
        h = F.gelu(F.linear(emb, wcache.cast(p0, "w", p0.weight, f32), wcache.cast(p0, "b", p0.bias, f32)))
        out = F.linear(h, wcache.cast(p2, "w", p2.weight, f32), wcache.cast(p2, "b", p2.bias, f32))
        return out.view(-1, self.num_tokens, self.embedding_dim)

    def forward(self, labels: torch.Tensor, is_training: bool = False, unconditional: bool = False,
                noise_std: float = 0.005) -> torch.Tensor:
        if unconditional:
            return self.null_embedding.expand(labels.shape[0] if labels.dim() > 0 else 1, -1)
        scalar = labels.dim() == 0
        emb = self._interp(labels)
        if is_training and noise_std > 0:
            emb = emb + torch.randn_like(emb) * noise_std
        out = self._project(emb)
        return out.squeeze(0) if scalar else out

    def get_negative_embedding(self, labels: torch.Tensor, is_training: bool = False, noise_std: float = 0.005) -> torch.Tensor:
        """CFG negative conditioning: label y contrasts against clamp(1 - y, 0, 1) (reference :182-221)."""
        if labels.dim() == 0:
            labels = labels.unsqueeze(0)
        return self.forward(torch.clamp(1.0 - labels, min=0.0, max=1.0), is_training=is_training, unconditional=False,
                            noise_std=noise_std)

    def get_ordinal_delta_embedding(self, source_labels: torch.Tensor, target_labels: torch.Tensor) -> torch.Tensor:
        """proj(E[target]) - proj(E[source]); subtraction after projection so biases cancel and equal labels give
        exactly zero (reference :246-294, invariant I1)."""
        scalar = source_labels.dim() == 0
        delta = self._project(self._interp(target_labels)) - self._project(self._interp(source_labels))
        return delta.squeeze(0) if scalar else delta

    def get_disease_delta_embedding(self, source_labels: torch.Tensor) -> torch.Tensor:
        return self.get_ordinal_delta_embedding(source_labels=source_labels, target_labels=torch.zeros_like(source_labels))
