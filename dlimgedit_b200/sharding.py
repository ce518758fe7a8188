"""Multi-GPU sharding helpers (one process per GPU, torch.distributed for the plumbing).

The segmentation path partitions by image: image i is encoded by rank i mod N and that rank answers all of its
prompts; weights are replicated.  The only exchange is the gather of per-prompt results (IoU scores; masks stay
where they were produced unless the caller asks for them)."""
from __future__ import annotations

from typing import List

import numpy as np


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_items, world))


def gather_scores(local: np.ndarray, n_images: int, rank: int, world: int) -> np.ndarray:
    """local: (len(shard), P) float32 scores of this rank's images -> (n_images, P) on every rank.
    Ragged shards are padded to the longest shard for the all_gather (NCCL and gloo need equal sizes)."""
    import torch
    import torch.distributed as dist
    per_rank = -(-n_images // world)
    P = local.shape[1] if local.ndim == 2 else 0
    on_gpu = world > 1 and dist.get_backend() == "nccl"  # NCCL moves device tensors only
    if world > 1:
        P_t = torch.tensor([P], device="cuda" if on_gpu else "cpu")
        dist.all_reduce(P_t, op=dist.ReduceOp.MAX)
        P = int(P_t.item())
    buf = torch.zeros(per_rank, P)
    if local.size:
        buf[: local.shape[0]] = torch.from_numpy(local)
    if world == 1:
        return buf[:n_images].numpy()
    if on_gpu:
        buf = buf.cuda()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.zeros((n_images, P), np.float32)
    for r in range(world):
        idx = shard_indices(n_images, r, world)
        full[idx] = out[r][: len(idx)].cpu().numpy()
    return full
