"""Mixed-model batches (BASELINE.json configs[2]: ViperX vx300 + Panda, 65,536 candidates over 2/4/8 B200).

The reference evaluates one (gripper, object) pair per process (/root/reference/mgs/cli/filter_to_stable.py:22-68): a
"mixed batch" is many such jobs.  Here they form ONE job: the candidates are bucketed by model, every
bucket is cut into chunks that each fill one GPU, and the chunks are handed out to the ranks

  * dynamically (`run_mixed`, the default with torch.distributed initialised): ranks pull the next chunk index from an atomic
    counter in the process group's store (a host-side work queue: no data-path collective, a few bytes per chunk), so a rank
    whose chunks finish early - candidates that fail after the close phase stop 5000 steps sooner - simply takes more; or
  * statically (`plan_static`): longest-processing-time assignment on an estimated cost, for launchers without a store.

Only the labels are exchanged at the end (one all_gather of uint8, like shard.gather_labels).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Bucket:
    """All candidates of one compiled model.  `run(lo, hi) -> (labels bool[hi-lo], env_steps int, kernel_seconds float)`
    evaluates candidates lo..hi-1 of this bucket on the calling rank's GPU."""
    name: str
    n: int
    run: object
    cost_per_candidate: float = 1.0  # relative (static planning only): e.g. nv^2-ish or a measured ms per candidate
    extra: dict = field(default_factory=dict)


def make_chunks(buckets, chunk: int):
    """[(bucket index, lo, hi)] in bucket order; a bucket's tail shorter than chunk/2 is merged into its last chunk."""
    out = []
    for b, bk in enumerate(buckets):
        lo = 0
        while lo < bk.n:
            hi = min(bk.n, lo + chunk)
            if bk.n - hi < chunk // 2:
                hi = bk.n
            out.append((b, lo, hi))
            lo = hi
    return out


def plan_static(buckets, chunk: int, world: int):
    """Longest-processing-time-first assignment of the chunks to `world` ranks.  Returns per-rank lists of chunk indices
    (into make_chunks(buckets, chunk)), each list in bucket order so that a rank's launches of one model are adjacent."""
    chunks = make_chunks(buckets, chunk)
    cost = [buckets[b].cost_per_candidate * (hi - lo) for b, lo, hi in chunks]
    load = [0.0] * world
    plan = [[] for _ in range(world)]
    for i in sorted(range(len(chunks)), key=lambda i: (-cost[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        plan[r].append(i)
        load[r] += cost[i]
    return [sorted(p) for p in plan], chunks


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


_QUEUE_SERIAL = [0]


def run_mixed(buckets, chunk: int, dynamic: bool = True, device=None):
    """Evaluate every bucket; returns (labels: list of bool[n_b] on every rank, stats dict of THIS rank).

    stats: chunks (how many this rank ran), candidates, env_steps, kernel_s (sum of the per-launch device times the
    buckets' `run` reported), wall_s."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    chunks = make_chunks(buckets, chunk)
    mine = []
    t0 = time.perf_counter()
    stats = dict(chunks=0, candidates=0, env_steps=0, kernel_s=0.0)
    local = [np.zeros(b.n, dtype=np.uint8) for b in buckets]

    def do(i):
        b, lo, hi = chunks[i]
        lab, steps, ks = buckets[b].run(lo, hi)
        local[b][lo:hi] = np.asarray(lab, dtype=np.uint8)
        stats["chunks"] += 1; stats["candidates"] += hi - lo; stats["env_steps"] += int(steps); stats["kernel_s"] += float(ks)
        mine.append(i)

    if world == 1:
        for i in range(len(chunks)):
            do(i)
    elif dynamic:
        # the work queue: an atomic counter in the rendezvous store.  Every call uses a fresh key (all ranks count calls alike).
        from torch.distributed import distributed_c10d as c10d
        store = c10d._get_default_store()
        key = f"mgs_mixed_queue_{_QUEUE_SERIAL[0]}"
        _QUEUE_SERIAL[0] += 1
        while True:
            i = store.add(key, 1) - 1
            if i >= len(chunks):
                break
            do(i)
    else:
        plan, _ = plan_static(buckets, chunk, world)
        for i in plan[rank]:
            do(i)
    stats["wall_s"] = time.perf_counter() - t0
    if world == 1:
        return [l.astype(bool) for l in local], stats
    # label exchange: every chunk was written by exactly one rank, all others hold zeros there -> a MAX all-reduce of the
    # concatenated uint8 labels is the gather (<= 64 KB for cfg3)
    import torch
    flat = torch.as_tensor(np.concatenate(local), device=device if dist.get_backend() == "nccl" else None)
    dist.all_reduce(flat, op=dist.ReduceOp.MAX)
    flat = flat.cpu().numpy().astype(bool)
    out, o = [], 0
    for b in buckets:
        out.append(flat[o:o + b.n]); o += b.n
    return out, stats
