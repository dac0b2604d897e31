"""Multi-GPU sharding of grasp candidates: contiguous blocks, labels gathered at the end.

Candidates are independent (the reference resets the simulation at the top of every loop
iteration, /root/reference/mgs/env/gravityless_object_grasping.py:158), so ranks exchange nothing
during the rollout; the only collective is an all_gather of the uint8 labels.
Works with any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of ceil(n/world) candidates for `rank` (keeps enough_stable's prefix
    semantics cheap: a rank's block is a contiguous piece of the sequential order)."""
    per = -(-n // world) if world > 0 else n
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def gather_labels(local: np.ndarray, n: int, group=None, device=None) -> np.ndarray:
    """all_gather of per-rank label blocks -> bool[n] on every rank (padding trimmed)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return np.asarray(local, dtype=bool)
    world = dist.get_world_size(group)
    per = -(-n // world)
    buf = torch.zeros(per, dtype=torch.uint8, device=device)
    buf[: len(local)] = torch.as_tensor(np.asarray(local, dtype=np.uint8), device=device)
    out = torch.empty(world * per, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.cpu().numpy().astype(bool)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n, r, world)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return np.concatenate(pieces) if pieces else np.zeros(0, dtype=bool)


def apply_enough_stable(labels: np.ndarray, enough_stable) -> np.ndarray:
    """Sequential `enough_stable` semantics of the reference loop (:151-156, :278-279): once that many
    successes have been seen, every later candidate is labelled False."""
    labels = np.asarray(labels, dtype=bool)
    if enough_stable is None:
        return labels
    before = np.concatenate([[0], np.cumsum(labels)[:-1]])
    return labels & (before < enough_stable)
