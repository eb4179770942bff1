"""Host-side wall-clock breakdown of one bench step (dev tool)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from deciphon_b200 import synth
from deciphon_b200.device import Device, PAIR_DTYPE
import ctypes as C
from deciphon_b200._lib import lib

class A: pass
def main():
    args = A(); args.seed = 20261018; args.profiles = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    args.reads_per_step = 48; args.read_len = 2000
    pool = synth.NodePool()
    sizes = synth.core_sizes(np.random.default_rng(args.seed), args.profiles)
    dev = Device(0)
    first = dev.pool_add(pool.emission, pool.trans)
    for p in range(args.profiles):
        ids, bmk = bench.profile_nodes(args.seed, p, sizes[p], pool)
        dev.profile_add(int(sizes[p]), bmk, pool.null_emission, pool.bg_emission, ids + first)
    R, L = 48, 2000
    reads = bench.make_reads(args.seed, 0, 3 * R, L, sizes, pool)
    dev.set_reads(reads)
    win = np.minimum(np.minimum(sizes * 50, 100000), L).astype(np.int32)
    for i in range(3):
        t0 = time.perf_counter()
        dev.score_grid(0, args.profiles, i * R, (i + 1) * R); dev.sync()
        t1 = time.perf_counter()
        idx = dev.hits_fetch()
        t2 = time.perf_counter()
        pr = np.zeros(len(idx), dtype=PAIR_DTYPE)
        pr["profile"] = idx // R; pr["seq"] = i * R + idx % R; pr["len"] = win[idx // R]
        n = len(pr)
        alt = np.empty(n, dtype=np.float32); nsteps = np.zeros(n, dtype=np.int32)
        t3 = time.perf_counter()
        dev._check(lib.dcpgpu_trace_pairs(dev._h, n, pr.ctypes.data_as(C.c_void_p), 1, alt.ctypes.data_as(C.c_void_p), nsteps.ctypes.data_as(C.c_void_p)))
        t4 = time.perf_counter()
        off = np.zeros(n + 1, dtype=np.int64); off[1:] = np.cumsum(nsteps)
        ids_ = np.zeros(int(off[-1]) + 1, dtype=np.uint16); sz = np.zeros(int(off[-1]) + 1, dtype=np.uint8)
        dev._check(lib.dcpgpu_trace_fetch(dev._h, off.ctypes.data_as(C.c_void_p), ids_.ctypes.data_as(C.c_void_p), sz.ctypes.data_as(C.c_void_p)))
        t5 = time.perf_counter()
        K = sizes[pr["profile"]]
        print(f"step {i}: score {1e3*(t1-t0):7.1f} ms (kernel {dev.last_kernel_ms():7.1f}) hits_fetch {1e3*(t2-t1):6.1f} prep {1e3*(t3-t2):5.1f} "
              f"trace_pairs {1e3*(t4-t3):7.1f} trace_fetch {1e3*(t5-t4):6.1f} | hits {n} Kmax {K.max() if n else 0} "
              f"trace cells {float((K*pr['len']).sum()):.3e} by class <=256:{int((K<=256).sum())} <=512:{int(((K>256)&(K<=512)).sum())} <=1024:{int(((K>512)&(K<=1024)).sum())} >1024:{int((K>1024).sum())}", flush=True)
    dev.close()
main()
