"""One process per GPU, utterances sharded `utts[rank::world]` — the pattern of the reference's only
multi-GPU inference tool (tools/vqgan/extract_vq.py:43-44, 161-194, 207). There is NO collective on the
decode path: every rank holds a full replica. torch.distributed (NCCL over NVLink on GPUs, gloo in the
CPU tests) is used only to replicate the weights at start-up and to collect the small results."""
from __future__ import annotations

from typing import Any, Sequence

import torch
import torch.distributed as dist


def rank_world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(items: Sequence[Any], rank: int, world: int) -> list:
    """files[RANK::WORLD_SIZE] (extract_vq.py:207)."""
    return list(items[rank::world])


def unshard(per_rank: Sequence[Sequence[Any]]) -> list:
    """Inverse of `shard`: per_rank[r] = results of items[r::world] -> results in the original order."""
    world = len(per_rank)
    n = sum(len(p) for p in per_rank)
    out: list = [None] * n
    for r, part in enumerate(per_rank):
        for k, v in enumerate(part):
            out[r + k * world] = v
    return out


def broadcast_state_dict(sd: dict | None, src: int = 0, device=None) -> dict:
    """Rank `src` holds the checkpoint; every other rank receives it tensor by tensor (ncclBroadcast over
    NVLink when the tensors live on GPUs). Key order and shapes travel as one small object broadcast."""
    rank, world = rank_world()
    if world == 1:
        return sd
    meta = [[(k, tuple(v.shape), v.dtype) for k, v in sd.items()]] if rank == src else [None]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for k, shape, dtype in meta[0]:
        t = sd[k].to(device) if rank == src else torch.empty(shape, dtype=dtype, device=device)
        if rank == src and device is None:
            t = sd[k]
        dist.broadcast(t, src=src)
        out[k] = t
    return out


def gather_objects(local: list) -> list:
    """All ranks' result lists, re-interleaved into the original utterance order (KBs of codes)."""
    rank, world = rank_world()
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, local)
    return unshard(parts)
