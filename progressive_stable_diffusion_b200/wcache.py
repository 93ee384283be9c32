"""Inference-time weight cache.

Modules keep their parameters exactly as a checkpoint / ``module.to(...)`` left them (the reference runs
``module.to(device).to(torch.float32)``, inference_pipeline_ip.py:593-595); the kernels want bf16 activations,
channels-last convolution filters, fused QKV matrices and fp32 norm affines.  ``get`` returns a derived tensor for a set
of source parameters and rebuilds it only when one of them was modified in place, re-allocated or moved, so derived
tensors have stable addresses across the replays of a captured CUDA graph.
"""

from __future__ import annotations

import weakref
from typing import Callable, Dict, Sequence, Tuple

import torch

_CACHE: Dict[Tuple, Tuple[Tuple, torch.Tensor]] = {}
_FINALIZERS: Dict[int, weakref.finalize] = {}
_MAX_ENTRIES = 8192


def _stamp(ts: Sequence[torch.Tensor]) -> Tuple:
    return tuple((t.data_ptr(), t._version, t.dtype, t.device) for t in ts)


def _drop(owner_id: int) -> None:
    for k in [k for k in _CACHE if k[0] == owner_id]:
        del _CACHE[k]
    _FINALIZERS.pop(owner_id, None)


_NAMESPACE = [""]


def set_namespace(ns: str) -> None:
    """Entries built under different compute dtypes coexist (captured graphs of either keep their tensors alive)."""
    _NAMESPACE[0] = ns


def get(owner: object, tag: str, sources: Sequence[torch.Tensor], build: Callable[[], torch.Tensor]) -> torch.Tensor:
    key = (id(owner), tag, _NAMESPACE[0])
    stamp = _stamp(sources)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    with torch.no_grad():
        val = build()
    if hit is not None and hit[1].shape == val.shape and hit[1].dtype == val.dtype and hit[1].device == val.device:
        hit[1].copy_(val)          # keep the address: captured graphs keep working after a weight update
        val = hit[1]
    _CACHE[key] = (stamp, val)
    if len(_CACHE) > _MAX_ENTRIES:                 # conditioning tensors come and go: bound the K/V entries
        for k in [k for k in _CACHE if k[1].startswith(("kkv", "vkv"))][: len(_CACHE) - _MAX_ENTRIES]:
            if k != key:
                del _CACHE[k]
    if id(owner) not in _FINALIZERS:
        try:
            _FINALIZERS[id(owner)] = weakref.finalize(owner, _drop, id(owner))
        except TypeError:
            pass
    return val


def cast(owner: object, tag: str, p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if p.dtype == dtype and p.is_contiguous():
        return p.detach()
    return get(owner, tag + str(dtype), (p,), lambda: p.detach().to(dtype).contiguous())


def conv_filter(owner: object, tag: str, p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return get(owner, tag + "cl" + str(dtype), (p,),
               lambda: p.detach().to(dtype).contiguous(memory_format=torch.channels_last))


def clear() -> None:
    _CACHE.clear()
