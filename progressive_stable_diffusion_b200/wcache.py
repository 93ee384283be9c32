"""Inference-time weight cache.

Modules keep their parameters exactly as a checkpoint / ``module.to(...)`` left them (the reference runs
``module.to(device).to(torch.float32)``, inference_pipeline_ip.py:593-595); the kernels want bf16 activations,
channels-last convolution filters, fused QKV matrices and fp32 norm affines.  ``get`` returns a derived tensor for a set
of source parameters and rebuilds it only when one of them was modified in place, re-allocated or moved, so derived
tensors have stable addresses across the replays of a captured CUDA graph.
"""

from __future__ import annotations

import weakref
from collections import OrderedDict
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
        hit[1].copy_(val)          # keep the address: a graph that is re-captured or replayed after ITS OWN refresh of
        val = hit[1]               # this entry keeps reading valid memory (ProgressionEngine re-validates weights per run)
    _CACHE[key] = (stamp, val)
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


# ------------------------------------------------------------------------------------------------ conditioning-derived tensors
_PINNED: Dict[int, "weakref.ref"] = {}


def pin(cond: torch.Tensor) -> None:
    """Mark ``cond`` as a long-lived conditioning buffer (a ProgressionEngine's static ``ehs``): ``CondCache`` entries built
    from it are never evicted while the tensor is alive - a captured graph reads them on every replay."""
    i = id(cond)
    _PINNED[i] = weakref.ref(cond, lambda _r, i=i: _PINNED.pop(i, None))


def _is_pinned(cond: torch.Tensor) -> bool:
    r = _PINNED.get(id(cond))
    return r is not None and r() is cond


class CondCache:
    """Per-processor cache of tensors derived from a conditioning tensor (the projected K/V of the condition tokens).

    An entry belongs to the conditioning tensor OBJECT through a weak reference - never to a raw address, which the caching
    allocator hands to the next ``torch.cat`` of the same shape - and goes away with it.  An in-place update of the tensor
    (version bump), its re-allocation, or a change of any weight in ``sources`` rebuilds the values into the SAME storage, so
    the addresses a captured graph baked in stay valid.  Transient conditionings (eager ``module(x, t, cond)`` calls) share a
    small LRU; entries of pinned tensors (see ``pin``) are only dropped when the tensor dies."""

    def __init__(self, max_transient: int = 2) -> None:
        self.max_transient = max_transient
        self.entries: "OrderedDict[Tuple, dict]" = OrderedDict()

    def _forget(self, key: Tuple) -> None:
        self.entries.pop(key, None)

    def get(self, cond: torch.Tensor, extra: Tuple, sources: Sequence[torch.Tensor], build: Callable[[], Tuple[torch.Tensor, ...]]):
        key = (id(cond), extra, _NAMESPACE[0])
        stamp = _stamp((cond, *sources))
        e = self.entries.get(key)
        if e is not None and e["ref"]() is not cond:          # the id was recycled by a different tensor object
            self._forget(key)
            e = None
        if e is not None and e["stamp"] == stamp:
            self.entries.move_to_end(key)
            return e["vals"]
        with torch.no_grad():
            vals = tuple(build())
        if e is not None and all(o.shape == v.shape and o.dtype == v.dtype and o.device == v.device for o, v in zip(e["vals"], vals)):
            for o, v in zip(e["vals"], vals):
                o.copy_(v)
            e["stamp"] = stamp
            self.entries.move_to_end(key)
            return e["vals"]
        self.entries[key] = {"ref": weakref.ref(cond, lambda _r, k=key: self._forget(k)), "stamp": stamp, "vals": vals}
        self.entries.move_to_end(key)
        transient = [k for k, v in self.entries.items() if not _is_pinned_ref(v["ref"])]     # oldest first
        for k in transient[: max(0, len(transient) - self.max_transient)]:
            if k != key:
                self._forget(k)
        return vals


def _is_pinned_ref(ref) -> bool:
    t = ref()
    return t is not None and _is_pinned(t)
