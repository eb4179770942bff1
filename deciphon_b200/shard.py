"""Profile sharding across the GPUs of one box and the host-side merge.

Every (window, profile) pair is independent (c-core/thread.c:59-72), so the database is cut
into contiguous profile ranges -- like the reference's per-thread partitions
(c-core/protein_reader.c:112-128) but balanced by total core size (DP cells) instead of by
profile count -- every rank scans every read, and rank 0 concatenates the per-rank results in
rank order, which IS profile order (what product_close does with the per-thread files,
c-core/product.c:63-80).  No collective touches the data path; torch.distributed is used only
to gather the small per-rank result objects and to take the max of the timers.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(core_sizes, world: int):
    """cuts[r]..cuts[r+1] = profiles of rank r; contiguous, covering, balanced by sum(K)."""
    sizes = np.asarray(core_sizes, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        c = int(np.searchsorted(csum, total * r / world))
        cuts.append(min(max(c, cuts[-1]), len(sizes)))
    cuts.append(len(sizes))
    return cuts


def merge_rank_results(per_rank):
    """per_rank[r] = list of (global_profile, seq, payload) from rank r, each already ordered by
    (profile, seq).  Returns the scan-order list: (profile index, batch order) ascending, the row
    order of the reference's products.tsv (c-core/thread.c:59-72, product.c:63-80)."""
    out = []
    last = (-1, -1)
    for rows in per_rank:
        for row in rows:
            key = (int(row[0]), int(row[1]))
            if key < last:
                raise ValueError("rank results are not in profile order; shards must be contiguous")
            last = key
            out.append(row)
    return out


def gather_to_rank0(obj, rank: int, world: int):
    """Gather small python objects on rank 0 (gloo or nccl process group); no-op at world 1."""
    if world == 1:
        return [obj]
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(obj, bucket, dst=0)
    return bucket
