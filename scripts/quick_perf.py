"""Quick device-side throughput probe of the score kernels per kernel class (dev tool)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deciphon_b200 import synth  # noqa: E402
from deciphon_b200.device import Device  # noqa: E402


def main():
    Ks = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 64, 128, 200, 224, 256, 512, 1000]
    nprof = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    nreads = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    L = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
    rng = np.random.default_rng(1)
    pool = synth.NodePool()
    dev = Device(0)
    first = dev.pool_add(pool.emission, pool.trans)
    reads = [synth.random_read(rng, L) for _ in range(nreads)]
    dev.set_reads(reads)
    for K in Ks:
        p0 = dev.num_profiles
        for _ in range(nprof):
            ids, bmk = synth.synth_profile_nodes(rng, K, pool)
            dev.profile_add(K, bmk, pool.null_emission, pool.bg_emission, ids + first)
        p1 = dev.num_profiles
        best = 1e30
        for rep in range(3):
            dev.score_grid(p0, p1, 0, nreads)
            dev.sync()
            best = min(best, dev.last_kernel_ms())
        cells = dev.last_cells()
        print(f"K={K:5d} profiles={nprof} reads={nreads} L={L} cells={cells:.3e} ms={best:9.3f} "
              f"GCUPS={cells / best / 1e6:9.2f} hits={len(dev.hits_fetch())} redo={dev.last_redo()}", flush=True)
    dev.close()


if __name__ == "__main__":
    main()
