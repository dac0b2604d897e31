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


def dist_state():
    """(world, rank, device for collectives) of the initialised default process group, else (1, 0, None)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            import torch
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else None
            return dist.get_world_size(), dist.get_rank(), dev
    except Exception:
        pass
    return 1, 0, None


def evaluate_sharded(n: int, run_range, enough_stable=None, chunk: int = 0) -> np.ndarray:
    """Labels bool[n] of candidates 0..n-1, evaluated by `run_range(lo, hi) -> bool[hi-lo]` on this rank's share.

    Without `enough_stable`: one contiguous block per rank, one label gather.  With it, the reference's early stop
    (gravityless_object_grasping.py:151-156: once `enough_stable` successes were seen the remaining candidates are
    labelled False WITHOUT being simulated) is kept: candidates are evaluated in sequential rounds of world x `chunk`
    (one chunk = what keeps one GPU full, so a smaller round would not finish sooner), the round's labels are gathered,
    and the loop stops as soon as the running success count reaches `enough_stable`.  The returned array equals the
    reference's sequential result."""
    world, rank, dev = dist_state()
    if enough_stable is None:
        if world == 1:
            return np.asarray(run_range(0, n), dtype=bool)
        lo, hi = shard_range(n, rank, world)
        return gather_labels(run_range(lo, hi), n, device=dev)
    chunk = max(1, int(chunk) if chunk else n)
    labels = np.zeros(n, dtype=bool)
    done, count = 0, 0
    while done < n and count < enough_stable:
        rn = min(n - done, world * chunk)
        lo, hi = shard_range(rn, rank, world)
        local = np.asarray(run_range(done + lo, done + hi), dtype=bool) if hi > lo else np.zeros(0, dtype=bool)
        got = gather_labels(local, rn, device=dev) if world > 1 else local
        labels[done:done + rn] = got
        count += int(got.sum())
        done += rn
    return apply_enough_stable(labels, enough_stable)
