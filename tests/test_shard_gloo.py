"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding + merge path.
Each rank "scans" its contiguous profile shard (with the oracle standing in for the GPU, which
this container does not have) and rank 0 merges; the merged rows must equal the 1-rank scan."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deciphon_b200 import shard, synth  # noqa: E402


def _workload():
    rng = np.random.default_rng(11)
    pool = synth.NodePool()
    sizes = [40, 7, 90, 33, 150, 12, 64]
    profs = [synth.synth_profile(rng, K, pool, name=f"P{i}") for i, K in enumerate(sizes)]
    reads = [synth.random_read(rng, int(n)) for n in (120, 33, 250, 61)]
    cons = np.argmax(profs[4].emission[:150, 20:84], axis=1)
    reads.append(np.stack([cons // 16, (cons // 4) % 4, cons % 4], axis=1).reshape(-1).astype(np.uint8))
    return sizes, profs, reads


def _scan(profs, first, reads):
    from oracle.oracle import Oracle
    orc = Oracle()
    rows = []
    for pi, p in enumerate(profs):
        costs = p.costs()
        for si, x in enumerate(reads):
            L = min(len(x), 50 * p.core_size)
            w = np.ascontiguousarray(x[:L])
            xt = orc.xtrans(L, True, False)
            rows.append((first + pi, si, (orc.null(costs[0], xt, w).tobytes(), orc.alt(costs, xt, w).tobytes())))
    return rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes, profs, reads = _workload()
    cuts = shard.shard_bounds(sizes, world)
    mine = _scan(profs[cuts[rank]:cuts[rank + 1]], cuts[rank], reads)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the timer reduction of bench.py
    parts = shard.gather_to_rank0(mine, rank, world)
    if rank == 0:
        q.put((cuts, shard.merge_rank_results(parts), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_properties():
    rng = np.random.default_rng(3)
    sizes = synth.core_sizes(rng, 5000)
    for world in (1, 2, 3, 4, 8):
        cuts = shard.shard_bounds(sizes, world)
        assert cuts[0] == 0 and cuts[-1] == len(sizes) and len(cuts) == world + 1
        assert all(b >= a for a, b in zip(cuts, cuts[1:]))
        loads = [int(sizes[a:b].sum()) for a, b in zip(cuts, cuts[1:])]
        assert max(loads) - min(loads) <= 2 * int(sizes.max())  # balanced by cells, not by count
    assert shard.shard_bounds([5], 4)[-1] == 1  # more ranks than profiles: empty shards allowed


def test_merge_rejects_out_of_order():
    with pytest.raises(ValueError):
        shard.merge_rank_results([[(3, 0, None)], [(1, 0, None)]])


@pytest.mark.timeout(300)
def test_two_rank_gloo_scan_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cuts, merged, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes, profs, reads = _workload()
    single = _scan(profs, 0, reads)
    assert tmax == 2.0
    assert 0 < cuts[1] < len(sizes)
    assert merged == single
