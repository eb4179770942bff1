"""Quick device-side timing of the trace pass (dev tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deciphon_b200 import synth
from deciphon_b200.device import Device, PAIR_DTYPE

def main():
    rng = np.random.default_rng(3)
    pool = synth.NodePool()
    dev = Device(0)
    first = dev.pool_add(pool.emission, pool.trans)
    L = 2000
    reads = [synth.random_read(rng, L) for _ in range(64)]
    dev.set_reads(reads)
    for K in (100, 250, 600, 1500):
        p0 = dev.num_profiles
        for _ in range(16):
            ids, bmk = synth.synth_profile_nodes(rng, K, pool)
            dev.profile_add(K, bmk, pool.null_emission, pool.bg_emission, ids + first)
        pairs = np.zeros(16 * 64, dtype=PAIR_DTYPE)
        pairs["profile"] = p0 + np.repeat(np.arange(16), 64)
        pairs["seq"] = np.tile(np.arange(64), 16)
        pairs["len"] = min(L, 50 * K)
        best = 1e9
        for rep in range(3):
            t = time.perf_counter()
            alt, off, ids_, sz = dev.trace_pairs_flat(pairs)
            best = min(best, time.perf_counter() - t)
        cells = float(pairs["len"].sum()) * K
        print(f"trace K={K:5d} pairs={len(pairs)} cells={cells:.3e} wall_ms={best*1e3:9.2f} GCUPS={cells/best/1e9:8.2f} steps={int(off[-1])}", flush=True)
    dev.close()

if __name__ == "__main__":
    main()
