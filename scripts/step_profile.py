"""Where a bench step's time goes beyond the score kernels (dev tool): the C ABI leg piece by piece,
then dcp_scan_run with DCP_TIMING=1.   python scripts/step_profile.py [profiles] [reads]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deciphon_b200 import synth, waves  # noqa: E402
from deciphon_b200.device import PAIR_DTYPE, Device  # noqa: E402


def main():
    seed = 20261018
    nprof = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 96
    L = 2000
    pool = synth.NodePool()
    sizes = synth.core_sizes(np.random.default_rng(seed), 20000)[:nprof]
    nodes = [bench.profile_nodes(seed, p, sizes[p], pool) for p in range(nprof)]
    reads = bench.make_reads(seed, 0, 3 * R, L, synth.core_sizes(np.random.default_rng(seed), 20000), pool)
    if os.environ.get("SKIP_CABI"):
        return plugin(nprof, R, sizes, pool, nodes, reads)
    dev = Device(0)
    first = dev.pool_add(pool.emission, pool.trans)
    for p in range(nprof):
        dev.profile_add(int(sizes[p]), nodes[p][1], pool.null_emission, pool.bg_emission, nodes[p][0] + first)
    dev.set_reads(reads)
    win = np.minimum(np.minimum(sizes * 50, 100000), L).astype(np.int32)
    lens = np.full(R, L, dtype=np.int64)
    for i in range(3):
        t = [time.perf_counter()]
        dev.score_grid(0, nprof, i * R, (i + 1) * R, True, False)
        idx = dev.hits_fetch()
        t.append(time.perf_counter())
        kms = dev.last_kernel_ms()
        pr = np.zeros(len(idx), dtype=PAIR_DTYPE)
        pr["profile"], pr["seq"], pr["len"] = idx // R, i * R + idx % R, win[idx // R]
        _, off, _ids, _sz = dev.trace_pairs_flat(pr, True, False)
        t.append(time.perf_counter())
        hit, hs, he = dev.match_build()
        t.append(time.perf_counter())
        w = waves.later_waves(dev, sizes, i * R, lens, pr, hit, he, True, False)
        t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        print(f"step {i}: grid+hits {d[0]:.1f} ms (kernels {kms:.1f})  trace+fetch {d[1]:.1f}  extents {d[2]:.1f}  "
              f"later waves {d[3]:.1f} ({w})  hits {len(idx)} path steps {int(off[-1])}", flush=True)
    dev.close()
    if not os.environ.get("SKIP_PLUGIN"):
        plugin(nprof, R, sizes, pool, nodes, reads)


def plugin(nprof, R, sizes, pool, nodes, reads):
    from deciphon_b200.scan import Batch, Scan, Sequence
    root = tempfile.mkdtemp(prefix="dcpprof_", dir="/tmp")
    db = os.path.join(root, "p.dcp")
    synth.write_synth_dcp(db, sizes, pool, lambda p: nodes[p])
    os.environ["DCP_TIMING"] = "1"
    cfgs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["1:1e11", "2:1e11", "3:1e11"]
    for cfg in cfgs:
        spg, chunk = cfg.split(":")
        os.environ["DCP_SHARDS_PER_GPU"] = spg
        os.environ["DCP_CHUNK_CELLS"] = chunk
        print("== config", cfg, flush=True)
        with Scan(db, 0, 1, True, False, False) as scan:
            print(f"== {scan.num_shards} shard(s) on {scan.num_gpus} GPU", flush=True)
            for i in (0, 1, 0, 1):
                batch = Batch()
                for j in range(i * R, (i + 1) * R):
                    batch.add(Sequence(j, "r%d" % j, bench.to_text(reads[j])))
                t0 = time.perf_counter()
                scan.run(os.path.join(root, "o"), batch)
                print(f"dcp_scan_run step {i}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
    import shutil
    shutil.rmtree(root, ignore_errors=True)


main()
