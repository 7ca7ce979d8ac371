"""Utterance sharding across GPUs.

The reference is single-device (reference gsv_tts/Config.py:59-70) and treats the segments of an
``infer_batched`` call as a queue feeding a fixed number of slots (GPT/t2s_model.py:696-722), already
length-balancing them for the vocoder (TTS.py:705-720).  Across GPUs the same units are independent:
one process per GPU holds a full replica, segments are dealt to ranks by predicted length, every rank runs
its own continuous batch + vocoder, and the host gathers the results in the caller's order.  The only
collective is the weight broadcast at load; nothing is exchanged per step.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.distributed as dist


def shard_by_length(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Deal request indices to ``world`` ranks: longest first, serpentine (0..w-1, w-1..0, ...), so
    every rank gets the same count (+-1) and near-equal total length.  Deterministic; ties by index."""
    if world < 1:
        raise ValueError("world must be >= 1")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        rnd, k = divmod(pos, world)
        shards[k if rnd % 2 == 0 else world - 1 - k].append(idx)
    return shards


def broadcast_state_dict(sd: Dict[str, torch.Tensor], src: int = 0, device=None) -> Dict[str, torch.Tensor]:
    """Weight blobs from rank ``src`` to every rank (NCCL over NVLink on GPUs, gloo on CPU).  Every rank
    passes a state dict with the same keys/shapes/dtypes (non-source ranks may pass uninitialised tensors).
    With ``device`` the tensors are broadcast GPU to GPU and STAY on the device (no host bounce: the native
    loaders take device tensors); without it they travel and return as CPU tensors.  This is the only
    collective of the whole path."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sd
    out = {}
    for k in sorted(sd):
        t = sd[k]
        buf = t.to(device) if device is not None else t.clone()
        dist.broadcast(buf, src)
        out[k] = buf
    return out


def broadcast_checkpoint(read_fn, src: int = 0, device=None):
    """Rank ``src`` calls ``read_fn() -> (meta, state_dict)`` (the only rank that touches the file system); ``meta`` (config /
    hps: small python objects) and the tensors' names, shapes and dtypes go out as objects, the tensors themselves with
    ``broadcast_state_dict``.  Returns ``(meta, state_dict)`` on every rank.  Single process: just ``read_fn()``."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return read_fn()
    rank = dist.get_rank()
    if rank == src:
        meta, sd = read_fn()
        head = [meta, [(k, tuple(v.shape), v.dtype) for k, v in sorted(sd.items())]]
    else:
        sd, head = None, [None, None]
    dist.broadcast_object_list(head, src)
    meta, layout = head
    if rank != src:
        sd = {k: torch.empty(shape, dtype=dtype) for k, shape, dtype in layout}
    return meta, broadcast_state_dict(sd, src, device)


def gather_in_order(local_results: Dict[int, object], n_total: int) -> List[object]:
    """Host-side gather of per-request results (tokens / waveforms as CPU objects), returned in request
    order on every rank.  ``local_results`` maps request index -> result for this rank's shard."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        merged = dict(local_results)
    else:
        parts: List[Dict[int, object]] = [None] * dist.get_world_size()   # type: ignore[list-item]
        dist.all_gather_object(parts, local_results)
        merged = {}
        for p in parts:
            merged.update(p)
    missing = [i for i in range(n_total) if i not in merged]
    if missing:
        raise RuntimeError(f"results missing for requests {missing[:8]}")
    return [merged[i] for i in range(n_total)]
