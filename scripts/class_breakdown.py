"""Per-kernel-class share of one bench step's score pass (dev tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from deciphon_b200 import synth
from deciphon_b200.device import Device, PAIR_DTYPE


def shape(K):
    """(W, Q, G) as csrc/layout.cuh layout_shape assigns them."""
    if os.environ.get("DCPGPU_SUBWARP", "1") != "0" and K <= 128:
        vl = 4 if K <= 32 else 8 if K <= 64 else 16
        q = max(5, -(-K // vl))
        return 1, 8 if q == 7 else q, 32 // vl
    W = 1
    while 32 * W * 8 < K:
        W *= 2
    Q = -(-K // (32 * W))
    return W, 8 if (W == 1 and Q == 7) else Q, 1


def main():
    seed = 20261018
    nprof = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    R, L = 48, 2000
    pool = synth.NodePool()
    sizes = synth.core_sizes(np.random.default_rng(seed), nprof)
    dev = Device(0)
    first = dev.pool_add(pool.emission, pool.trans)
    for p in range(nprof):
        ids, bmk = bench.profile_nodes(seed, p, sizes[p], pool)
        dev.profile_add(int(sizes[p]), bmk, pool.null_emission, pool.bg_emission, ids + first)
    reads = bench.make_reads(seed, 0, R, L, sizes, pool)
    dev.set_reads(reads)
    win = np.minimum(np.minimum(sizes * 50, 100000), L).astype(np.int32)
    WQ = np.array([shape(int(k)) for k in sizes])
    dev.score_grid(0, nprof, 0, R); dev.sync()
    dev.score_grid(0, nprof, 0, R); dev.sync()
    whole = dev.last_kernel_ms()
    tot = 0.0
    rows = []
    for W, G in ((1, 8), (1, 4), (1, 2), (1, 1), (2, 1), (4, 1), (8, 1), (16, 1)):
        for Q in range(1, 9):
            sel = np.nonzero((WQ[:, 0] == W) & (WQ[:, 1] == Q) & (WQ[:, 2] == G))[0]
            if W == 16:
                sel = np.nonzero(WQ[:, 0] >= 16)[0]
                if Q > 1:
                    continue
            if not len(sel):
                continue
            pr = np.zeros(len(sel) * R, dtype=PAIR_DTYPE)
            pr["profile"] = np.repeat(sel, R)
            pr["seq"] = np.tile(np.arange(R), len(sel))
            pr["len"] = np.repeat(win[sel], R)
            dev.score_pairs(pr, multi_hits=True)
            dev.score_pairs(pr, multi_hits=True)
            ms = dev.last_kernel_ms()
            cells = float((sizes[pr["profile"]].astype(np.float64) * pr["len"]).sum())
            rows.append((W, Q, G, len(sel), cells, ms, dev.last_redo()))
            tot += ms
    for W, Q, G, n, cells, ms, redo in rows:
        print(f"W={W:2d} G={G} Q={Q} profiles={n:6d} cells={cells:.3e} ms={ms:8.2f} share={ms/tot:6.1%} GCUPS={cells/ms/1e6:7.1f} redo={redo}")
    print(f"sum of classes {tot:.1f} ms; whole grid pass {whole:.1f} ms")
    dev.close()


main()
