#!/usr/bin/env python
"""bench.py -- headline benchmark of the scan hot path (BASELINE.json: GCUPS and reads/s).

Workload (BASELINE.json configs[2], SURVEY 8(d) "config 3"): a synthetic Pfam-scale profile
database (default 20,000 profiles, core sizes ~ clipped log-normal, mean 200, 20..2000; nodes
bootstrapped from the 576 real nodes of the golden minifam database) scanned with synthetic
2,000-nt reads carrying 10 % indel/substitution errors (1 % of them with an embedded
back-translated profile consensus so that hits exist).  A STEP is one pass of the hot path
over one batch of reads against every resident profile: score pass (null + alt Viterbi of
every (read, profile) pair), hit list (lrt >= 0), trace pass (trellis + on-GPU path decoding)
for the hits.  At N > 1 the database is sharded by profile (contiguous ranges balanced by
total core size), every rank sees every read, no collective on the data path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Prints ONE JSON line (rank 0).
  `value`    GCUPS of full steps through the C ABI (include/dcpgpu.h) with the reads resident in
             HBM: every window.c window of every pair (profiles of fewer than 40 nodes take two or
             three windows per 2,000-nt read), CUDA events, max over ranks.
  `e2e`      the same batches through the REFERENCE API, dcp_scan_setup / dcp_scan_run of
             libdeciphon_b200.so (include/deciphon_b200.h): the database is a real .dcp file on a
             local file system (/dev/shm when a third of the free RAM holds its 24.5 GB, else
             /tmp), reads are host strings in a dcp_batch, and every step ends with its
             products.tsv written (encode, H2D, D2H, row formatting and file writing inside the
             timed region).  At N > 1 one process drives the N GPUs as N profile
             shards (num_threads = N), exactly what a caller of the reference API gets.
  `e2e_cabi` the same batches through the C ABI with host buffers (no row formatting).
  `roofline` the score kernels against the FP32 issue ceiling (and the best measured add/min mix).
  `strong`   (N > 1) a fixed 96-read batch over the N shards, beside the weak-scaling `value`.
  `configs`  (N = 1) timed legs of BASELINE.json configs 2 and 5 through dcp_scan_run, each
             checked in the run against the reference's own CPU code (cpu_baseline leg).
  `cpu_baseline` / `--impl reference`: the reference's viterbi.c/trellis.c/xtrans.c (oracle/_ref,
             compiled from /root/reference) driven like its scan loop on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from deciphon_b200 import synth  # noqa: E402

OPS_PER_CELL = 33  # fp32 add/min per DP cell, factored recurrence


def measured_traffic():
    """DRAM bytes of the score pass of one step at the default workload, from the committed ncu
    capture (dram__bytes_read.sum + dram__bytes_write.sum; profiles/README.md).  None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic_v5.json")) as f:
            return float(json.load(f)["score_pass_dram_bytes_per_step"])
    except Exception:
        return None


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--profiles", type=int, default=20000)
    ap.add_argument("--reads-per-step", type=int, default=96,
                    help="reads per step PER GPU: a step's batch is this many reads times --gpus, so the work of a "
                         "rank (its profile shard x the batch) stays fixed as GPUs are added (weak scaling)")
    ap.add_argument("--read-len", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=20261018)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU work of the cpu_baseline leg (and cap of a reference-arm step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--contexts", type=int, default=2, help="C ABI contexts (sub-shards, host threads) per GPU in the measured legs")
    ap.add_argument("--no-plugin", action="store_true", help="skip the dcp_scan_run legs (e2e falls back to the C ABI leg)")
    ap.add_argument("--plugin-profiles", type=int, default=0,
                    help="profiles of the database the dcp_scan_run leg scans (0 = all; fewer if the disk has no room)")
    ap.add_argument("--tmp", default=os.environ.get("DCP_BENCH_TMP", ""), help="directory for the .dcp file and products")
    return ap.parse_args()


# ---- workload -------------------------------------------------------------------------------

def profile_nodes(seed: int, p: int, K: int, pool):
    rng = np.random.default_rng([seed, 1, p])
    return synth.synth_profile_nodes(rng, int(K), pool)


from deciphon_b200.shard import shard_bounds  # noqa: E402


def make_reads(seed: int, first: int, count: int, L: int, sizes: np.ndarray, pool, err=0.10, frac=0.01):
    """Reads first..first+count of the synthetic read set (deterministic per read index)."""
    out = []
    for i in range(first, first + count):
        rng = np.random.default_rng([seed, 2, i])
        if rng.random() < frac:
            q = int(rng.integers(0, len(sizes)))
            ids, _ = profile_nodes(seed, q, sizes[q], pool)
            cons = synth.consensus_dna(pool, ids)[: max(30, L - 60)]
            a = int(rng.integers(0, max(1, L - len(cons))))
            x = np.concatenate([synth.random_read(rng, a), cons, synth.random_read(rng, max(0, L - a - len(cons)))])
        else:
            x = synth.random_read(rng, L)
        out.append(synth.fixed_length(rng, synth.mutate(rng, x, err), L))
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                    "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in o.strip().split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---- reference / CPU leg --------------------------------------------------------------------

def stratified_pick(sizes, nprof):
    """nprof profiles at evenly spaced ranks of the core-size distribution, largest first."""
    order = np.argsort(sizes, kind="stable")
    return order[np.linspace(0, len(order) - 1, nprof).round().astype(int)][::-1]


def deal_partitions(pick, cost, cores, per_thread, seed):
    """Order the sample so that the reference's contiguous count partitions (scan.c:188-208,
    partition_size.c:13-16: `cores` partitions of `per_thread` consecutive profiles) carry near-equal
    cost -- what partitioning 20,000 profiles by count gives the real scan (1,250 profiles per thread
    average out to a few per cent), and what a sample of one or two profiles per thread does not.
    cost[i] = measured single-thread seconds of pick[i] (time is not linear in the core size: short
    profiles pay a per-row overhead), dealt largest first in a snake."""
    by_cost = np.argsort(-np.asarray(cost), kind="stable")
    parts = [[] for _ in range(cores)]
    for r in range(per_thread):  # round r hands one profile to every partition
        row = by_cost[r * cores:(r + 1) * cores]
        for j, i in enumerate(row if r % 2 == 0 else row[::-1]):
            parts[j].append(int(i))
    rng = np.random.default_rng(seed)
    for part in parts:
        rng.shuffle(part)  # mixed sizes inside a partition, like a real database
    return np.asarray([i for part in parts for i in part], dtype=np.int64)


def cpu_scan_sample(args, sizes, pool, target_seconds, simd=None):
    """The reference's viterbi.c/trellis.c/xtrans.c (oracle/_ref) driven like its scan loop on all
    host cores over a bounded, seed-fixed, K-stratified sample of the same workload: 32 profiles
    per thread, balanced by sum K over the count partitions, every window.c window of a pair."""
    from deciphon_b200.dcp_file import Profile
    from oracle.oracle import Reference, ref_lib_path
    if simd:
        os.environ["DCP_REF_SIMD"] = simd
    try:
        ref = Reference(ref_lib_path())
    finally:
        os.environ.pop("DCP_REF_SIMD", None)
    cores = os.cpu_count() or 1
    per_thread = 32
    pick = stratified_pick(sizes, per_thread * cores)
    profs = []
    for p in pick:
        ids, bmk = profile_nodes(args.seed, int(p), sizes[p], pool)
        tr, em = pool.trans[ids], pool.emission[ids]
        pr = Profile("s%d" % p, 1, "", int(sizes[p]), pool.null_emission, pool.bg_emission,
                     np.concatenate([tr, tr[-1:]]), np.concatenate([em, em[-1:]]), bmk)
        profs.append(ref.profile(pr.costs()))
    probe_reads = make_reads(args.seed, 0, 2, args.read_len, sizes, pool)
    # every profile's own single-thread cost on one read, then the balanced deal
    cost = [min(ref.scan([pf], probe_reads[:1], True, False, 1)["seconds"] for _ in range(2)) for pf in profs]
    deal = deal_partitions(pick, cost, cores, per_thread, args.seed)
    pick, profs = pick[deal], [profs[i] for i in deal]
    ref.scan(profs, probe_reads, True, False, cores)  # first touch of the tables
    t = ref.scan(profs, probe_reads, True, False, cores)
    rate = t["cells"] / max(t["seconds"], 1e-9)
    per_read = t["cells"] / 2
    nreads = int(max(2, min(4000, round(target_seconds * rate / per_read))))
    reads = make_reads(args.seed, 0, nreads, args.read_len, sizes, pool)
    r = ref.scan(profs, reads, True, False, cores)
    ts = r["thread_seconds"]
    gcups = r["cells"] / r["seconds"] / 1e9
    sample = (f"{len(pick)} K-stratified profiles = {per_thread} per thread x {cores} threads, count partitions balanced by "
              f"each profile's measured single-thread time (sum K = {int(sizes[pick].sum())}) x {nreads} reads of {args.read_len} nt, every window.c "
              f"window, xtrans + null + alt Viterbi per window, trellis+unzip for lrt>=0 ({r['hits']} hits), "
              f"{r['cells']:.3e} cells in {r['seconds']:.2f} s; per-thread busy {ts.min():.2f}..{ts.max():.2f} s, "
              f"parallel efficiency {r['parallel_efficiency']:.3f}, {gcups / cores:.4f} GCUPS per core; "
              f"{os.path.basename(ref.path)}")
    return {"gcups": gcups, "reads_per_s": None, "cores": cores, "sample": sample,
            "seconds": r["seconds"], "cells": r["cells"], "profs": profs, "ref": ref, "nreads": nreads,
            "sumK_sample": int(sizes[pick].sum()), "parallel_efficiency": r["parallel_efficiency"],
            "reads": reads}


def cpu_small_configs(args, pool, gpu_small):
    """Configs 2 and 5 on the reference's own CPU code (oracle/_ref), one thread per profile up to
    the core count like scan.c:105, and the in-run parity check of the GPU legs: the DP cells of a
    run (every window start depends on the hits decoded before it) and the number of windows that
    passed the lrt gate must be identical."""
    from oracle.oracle import Reference, ref_lib_path
    ref = Reference(ref_lib_path())
    out = {}
    for name, (profs, reads, _text) in small_config_workloads(args, pool).items():
        rp = [ref.profile(p.costs()) for p in profs]
        threads = min(len(rp), os.cpu_count() or 1)
        ref.scan(rp, reads[:1], True, False, threads)
        r = ref.scan(rp, reads, True, False, threads)
        rec = {"gcups": r["cells"] / r["seconds"] / 1e9, "reads_per_s": len(reads) / r["seconds"], "ms_per_run": 1e3 * r["seconds"],
               "threads": threads, "cells": r["cells"], "lrt_windows": int(r["hits"]), "kind": "reference"}
        g = (gpu_small or {}).get(name)
        if isinstance(g, dict) and "cells" in g:
            rec["parity_with_gpu_leg"] = bool(g["cells"] == r["cells"] and g["lrt_windows"] == int(r["hits"]))
        out[name] = rec
    return out


def cpu_baseline_record(args, sizes, pool, gpu_small=None):
    """cpu_baseline of the measured arm: the widest SIMD build this host runs, plus the reference
    Makefile's default level (-mavx2, c-core/Makefile:10-12) on a third of the sample, plus the
    CPU side (and parity check) of the config 2 / config 5 legs."""
    c = cpu_scan_sample(args, sizes, pool, args.cpu_seconds)
    rec = {"value": c["gcups"], "unit": "GCUPS", "cores": c["cores"], "kind": "reference", "sample": c["sample"],
           "parallel_efficiency": c["parallel_efficiency"],
           # what the same threads would deliver with no idle time at all (sum of the threads' busy-time rates):
           # an upper bound for the reference on this host, whatever the partitioning
           "value_at_perfect_balance": c["gcups"] / max(c["parallel_efficiency"], 1e-9)}
    wide = os.path.basename(c["ref"].path)
    c.clear()
    try:
        rec["configs"] = cpu_small_configs(args, pool, gpu_small)
    except Exception as e:
        rec["configs"] = {"unavailable": f"{type(e).__name__}: {e}"}
    if "avx512" in wide:
        try:
            d = cpu_scan_sample(args, sizes, pool, max(3.0, args.cpu_seconds / 3), simd="avx2")
            rec["makefile_default_avx2"] = {"value": d["gcups"], "unit": "GCUPS", "cores": d["cores"],
                                            "parallel_efficiency": d["parallel_efficiency"], "sample": d["sample"]}
        except Exception as e:
            rec["makefile_default_avx2"] = {"value": None, "sample": f"unavailable: {e}"}
    return rec


def run_reference(args, rank):
    if rank != 0:
        return
    pool = synth.NodePool()
    sizes = synth.core_sizes(np.random.default_rng(args.seed), args.profiles)
    per_step = min(max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup))), args.cpu_seconds)
    first = cpu_scan_sample(args, sizes, pool, per_step)
    ref, profs, nreads = first["ref"], first["profs"], first["nreads"]
    times, cells, effs = [], 0.0, []
    for i in range(args.warmup + args.steps):
        reads = make_reads(args.seed, i * nreads, nreads, args.read_len, sizes, pool)
        r = ref.scan(profs, reads, True, False, first["cores"])
        if i >= args.warmup:
            times.append(r["seconds"])
            cells += r["cells"]
            effs.append(r["parallel_efficiency"])
    gcups = cells / sum(times) / 1e9
    full_cells_per_read = float(np.minimum(sizes * 50, args.read_len).astype(np.float64) @ sizes)
    line = {
        "impl": "reference", "metric": "GCUPS (nt x node DP cells/s), Pfam-scale scan", "value": gcups, "unit": "GCUPS",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sizes, None),
        "reads_per_s": gcups * 1e9 / full_cells_per_read,
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": first["cores"], "kind": "reference",
                         "parallel_efficiency": float(np.mean(effs)),
                         "value_at_perfect_balance": gcups / max(float(np.mean(effs)), 1e-9),
                         "sample": first["sample"] + f"; each step = a fresh batch of {nreads} reads on the same profiles"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sizes, shard):
    return {
        "workload": (f"config3: synthetic Pfam-scale DB, {args.profiles} profiles (clipped log-normal core size, mean "
                     f"{float(sizes.mean()):.0f}, {int(sizes.min())}..{int(sizes.max())}, sum K = {int(sizes.sum())}) vs synthetic "
                     f"{args.read_len}-nt reads with 10% errors; step = score pass + hit list + trace pass of one batch of "
                     f"{args.reads_per_step * args.gpus} reads ({args.reads_per_step} per GPU) against all resident profiles"),
        "profiles": args.profiles, "reads_per_step": args.reads_per_step * args.gpus,
        "reads_per_step_per_gpu": args.reads_per_step, "read_len": args.read_len,
        "multi_hits": True, "hmmer3_compat": False, "seed": args.seed,
        "parallelism": f"profile-sharded x{args.gpus}, no collective; {args.contexts} contexts (sub-shards, host threads) per GPU",
        "l2_policy": "inputs larger than L2 (profile tables >> 126 MB); no flush",
        "shard": shard,
    }


# ---- B200 leg -------------------------------------------------------------------------------

def to_text(x: np.ndarray) -> str:
    return np.frombuffer(b"ACGT", dtype=np.uint8)[x].tobytes().decode()


def tmp_root(args, need_bytes: int):
    """A fresh directory on a local file system with room for need_bytes, else None."""
    import shutil
    import tempfile
    # memory-backed first (when a third of the available RAM holds it): the write-back of tens of GB to
    # disk would otherwise run under the timed steps that follow
    cands = [args.tmp] if args.tmp else ["/dev/shm", "/tmp", "/var/tmp"]
    for c in cands:
        try:
            os.makedirs(c, exist_ok=True)
            free = shutil.disk_usage(c).free
            if c.startswith("/dev/shm"):  # RAM: leave room for the page cache of the read-back and the host copies
                import psutil
                free = min(free, psutil.virtual_memory().available // 3)
            if free > need_bytes * 1.1 + (2 << 30):
                return tempfile.mkdtemp(prefix="dcpbench_", dir=c)
        except Exception:
            continue
    return None


def plugin_leg(args, world, sizes, pool, nodes_of, reads, R, nsteps_total):
    """`e2e`: the batches of the measured leg through dcp_scan_setup / dcp_scan_run (reference API,
    libdeciphon_b200.so) on a .dcp file of the same synthetic database.  One process, `world` GPUs."""
    import shutil

    from deciphon_b200.scan import Batch, Scan, Sequence
    nprof = min(args.plugin_profiles or len(sizes), len(sizes))
    root = None
    while True:
        need = (int(sizes[:nprof].sum()) + nprof) * 6037 + nprof * 32768
        root = tmp_root(args, need)
        if root is not None or nprof <= 250:
            break
        nprof //= 2  # no room for the whole database: scan a prefix of it and say so
    if root is None:
        return {"unavailable": "no local directory with room for the .dcp file"}
    db = os.path.join(root, "bench.dcp")
    out = os.path.join(root, "products")
    try:
        t0 = time.perf_counter()
        info = synth.write_synth_dcp(db, sizes, pool, nodes_of, count=nprof)
        write_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        scan = Scan(db, 0, world, True, False, False)
        setup_s = time.perf_counter() - t0
        texts = [to_text(r) for r in reads]

        def step(i):
            batch = Batch()
            for j in range(i * R, (i + 1) * R):
                batch.add(Sequence(j, "read%d" % j, texts[j]))
            scan.run(out, batch)

        for i in range(args.warmup):
            step(i)
        c0 = scan.counters()
        t0 = time.perf_counter()
        for i in range(args.warmup, nsteps_total):
            step(i)
        secs = time.perf_counter() - t0
        c1 = scan.counters()
        d = {k: c1[k] - c0[k] for k in c0}
        with open(os.path.join(out, "products.tsv"), "rb") as fh:
            tsv = fh.read()
        gpus, shards = scan.num_gpus, scan.num_shards
        scan.free()
        return {"value": d["cells"] / secs / 1e9, "unit": "GCUPS", "seconds": secs, "cells": d["cells"],
                "h2d_bytes_per_step": int(d["h2d_bytes"] / args.steps), "d2h_bytes_per_step": int(d["d2h_bytes"] / args.steps),
                "launches": int(d["launches"]), "windows": int(d["windows"]), "lrt_windows": int(d["lrt_windows"]),
                "replanned_windows": int(d["speculative_windows"]),
                "reads_per_s": args.steps * R / secs, "gpus": gpus, "shards": shards, "profiles": nprof,
                "dcp_bytes": info["bytes"], "dcp_write_s": write_s, "setup_s": setup_s,
                "rows_last_step": tsv.count(b"\n") - 1, "tsv_bytes_last_step": len(tsv),
                "api": "dcp_batch_add x reads + dcp_scan_run -> products.tsv (libdeciphon_b200.so), num_threads = %d" % world,
                "dir": os.path.dirname(root)}
    finally:
        shutil.rmtree(root, ignore_errors=True)


def small_config_workloads(args, pool):
    """BASELINE.json configs 2 and 5 (SURVEY 8d recipes), seed-fixed."""
    rng = np.random.default_rng([args.seed, 25])
    p3 = synth.synth_profile(rng, 3, pool, name="MASSIVE3")  # the shape of c-core/massive.hmm (LENG 3)
    r2 = [synth.fixed_length(rng, synth.mutate(rng, synth.random_read(rng, 1000), 0.10), 1000) for _ in range(1000)]
    profs5 = [synth.synth_profile(rng, K, pool, name="C5_%d" % K) for K in (50, 100, 200, 500, 1000, 2000)]
    parts = []
    for g in range(25):  # ~25 embedded genes
        p = profs5[g % len(profs5)]
        cons = np.argmax(p.emission[:p.core_size, 20:84], axis=1)
        cons = np.stack([cons // 16, (cons // 4) % 4, cons % 4], axis=1).reshape(-1).astype(np.uint8)
        a = 3 * int(rng.integers(0, max(1, (len(cons) - 600) // 3 + 1)))
        parts += [synth.random_read(rng, int(rng.integers(200, 500))), cons[a:a + 600]]
    x5 = synth.fixed_length(rng, synth.mutate(rng, np.concatenate(parts), 0.10), 24000)
    return {"config2": ([p3], r2, "K = 3 profile (massive.hmm shape, synthetic nodes) vs 1,000 reads x 1,000 nt, 10% errors; "
                                  "150-nt window.c windows"),
            "config5": (profs5, [x5], "one 24,000-nt read (25 embedded genes, 10% errors) vs profiles of 50/100/200/500/1000/2000 "
                                      "nodes; window.c windows, full traceback of every lrt >= 0 window")}


def small_config_legs(args, pool):
    """Timed dcp_scan_run legs of configs 2 and 5 on one GPU (rows written, every window)."""
    import shutil

    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    root = tmp_root(args, 64 << 20)
    res = {}
    if root is None:
        return res
    try:
        for name, (profs, reads, text) in small_config_workloads(args, pool).items():
            db = os.path.join(root, name + ".dcp")
            write_dcp(db, profs)
            texts = [to_text(r) for r in reads]
            in_run = [0.0]
            with Scan(db, 0, 1, True, False, False) as scan:
                def run():
                    batch = Batch()
                    for j, t in enumerate(texts):
                        batch.add(Sequence(j, "r%d" % j, t))
                    t1 = time.perf_counter()
                    scan.run(os.path.join(root, name), batch)
                    in_run[0] += time.perf_counter() - t1
                run()  # warm-up
                in_run[0] = 0.0
                reps = 3
                c0 = scan.counters()
                t0 = time.perf_counter()
                for _ in range(reps):
                    run()
                secs = (time.perf_counter() - t0) / reps
                c1 = scan.counters()
            rows = open(os.path.join(root, name, "products.tsv"), "rb").read().count(b"\n") - 1
            cells = (c1["cells"] - c0["cells"]) / reps
            res[name] = {"workload": text, "gcups": cells / secs / 1e9, "reads_per_s": len(reads) / secs, "ms_per_run": 1e3 * secs,
                         "ms_in_dcp_scan_run": 1e3 * in_run[0] / reps,  # the rest is dcp_batch_add from Python, one call per read
                         "cells": cells, "windows": int((c1["windows"] - c0["windows"]) / reps),
                         "lrt_windows": int((c1["lrt_windows"] - c0["lrt_windows"]) / reps), "rows": rows,
                         "replanned_windows": int((c1["speculative_windows"] - c0["speculative_windows"]) / reps),
                         "launches_per_run": int((c1["launches"] - c0["launches"]) / reps),
                         "api": "dcp_scan_run (libdeciphon_b200.so), host strings in, products.tsv out, 1 GPU"}
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return res


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from deciphon_b200 import waves
    from deciphon_b200.device import PAIR_DTYPE, Device

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (deciphon_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    # (NCCL_DEBUG is left to the caller: unset, NCCL prints nothing and stdout is the one JSON line)
    cpu_group = None
    if world > 1:
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")  # one node; the box's hostname may not resolve
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        cpu_group = dist.new_group(backend="gloo")  # host-side waits (no kernel spinning on a GPU another leg uses)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_ranks(x: float, op) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.MAX) if world > 1 else x

    def sum_over_ranks(x):
        return reduce_ranks(x, dist.ReduceOp.SUM) if world > 1 else x

    pool = synth.NodePool()
    sizes = synth.core_sizes(np.random.default_rng(args.seed), args.profiles)
    cuts = shard_bounds(sizes, world)
    p0, p1 = cuts[rank], cuts[rank + 1]
    nodes = {}

    def nodes_of(p):
        if p not in nodes:
            nodes[p] = profile_nodes(args.seed, p, sizes[p], pool)
        return nodes[p]

    # args.contexts C-ABI contexts per GPU, each holding a contiguous sub-shard (balanced by core size) on
    # its own stream and driven by its own host thread: while one context's thread collects hits,
    # traces them and prepares the next windows, the other context's score kernels keep the GPU busy
    stream = torch.cuda.current_stream()
    nctx = max(1, min(args.contexts, (p1 - p0) // 64 or 1))
    sub = [p0 + int(c) for c in shard_bounds(sizes[p0:p1], nctx)]
    devs, streams = [], []
    t_build = time.time()
    for c in range(nctx):
        dev = Device(local_rank)
        st_c = torch.cuda.Stream()
        dev.set_stream(st_c.cuda_stream)
        first = dev.pool_add(pool.emission, pool.trans)
        for p in range(sub[c], sub[c + 1]):
            ids, bmk = nodes_of(p)
            dev.profile_add(int(sizes[p]), bmk, pool.null_emission, pool.bg_emission, ids + first)
        dev.sync()
        devs.append(dev)
        streams.append(st_c)
    t_build = time.time() - t_build
    L = args.read_len
    nsteps_total = args.warmup + args.steps
    Ks_all = sizes[p0:p1]
    multi_window = int(np.count_nonzero(np.minimum(Ks_all * 50, 100000) < L))

    def run_legs(R, reads, host_buffers, kernels_only=False):
        """args.warmup + args.steps steps of R reads each.  host_buffers: the step's reads are handed over
        as host symbols (packed + copied inside the step) and every score comes back.  kernels_only:
        just the score pass of the first windows, one context at a time (the roofline's kernel time)."""
        lens = np.full(R, L, dtype=np.int64)
        offs = np.arange(R + 1, dtype=np.int64) * L
        pinned = None
        if host_buffers:
            pinned = [torch.from_numpy(np.concatenate(reads[i * R:(i + 1) * R])).pin_memory() for i in range(nsteps_total)]
        else:
            for dev in devs:
                dev.set_reads(reads)

        def step(c, i, st):
            dev, Ks = devs[c], sizes[sub[c]:sub[c + 1]]
            nprof = len(Ks)
            win = np.minimum(np.minimum(Ks * 50, 100000), L).astype(np.int32)  # first window per profile
            seq0 = 0 if host_buffers else i * R
            if host_buffers:
                dev.set_reads_packed(pinned[i].numpy(), offs)
            dev.score_grid(0, nprof, seq0, seq0 + R, True, False)
            if kernels_only:
                st["score_ms"] += dev.last_kernel_ms()
                st["grid_cells"] += dev.last_cells()
                return
            if host_buffers:
                dev.scores_fetch(nprof * R)
            idx = dev.hits_fetch()
            pr = np.zeros(len(idx), dtype=PAIR_DTYPE)
            pr["profile"], pr["seq"], pr["start"], pr["len"] = idx // R, seq0 + idx % R, 0, win[idx // R]
            hit, _hs, he, nst = waves.trace_hits(dev, pr, True, False, Ks=Ks)
            st["steps"] += nst
            st["hits"] += len(pr)
            if np.any(np.minimum(Ks * 50, 100000) < L):  # windows 2.. of the profiles of fewer than L/50 nodes
                w = waves.later_waves(dev, Ks, seq0, lens, pr, hit, he, True, False)
                st["hits"] += w["hits"]
                st["steps"] += w["steps"]
                st["later_windows"] += w["pairs"]

        zero = lambda: {"score_ms": 0.0, "grid_cells": 0.0, "hits": 0, "steps": 0, "later_windows": 0}  # noqa: E731

        def run_steps(first_step, last_step):
            """Every context runs its steps on its own thread (free-running: their phases interleave)."""
            stats = [zero() for _ in range(nctx)]
            errs = []

            def work(c):
                try:
                    torch.cuda.set_device(local_rank)
                    for i in range(first_step, last_step):
                        step(c, i, stats[c])
                except BaseException as e:  # noqa: BLE001
                    errs.append(e)

            if kernels_only or nctx == 1:
                for c in range(nctx):
                    work(c)
            else:
                ts = [threading.Thread(target=work, args=(c,)) for c in range(nctx)]
                for t in ts:
                    t.start()
                for t in ts:
                    t.join()
            if errs:
                raise errs[0]
            return {k: sum(s_[k] for s_ in stats) for k in stats[0]}

        last = nsteps_total
        if kernels_only:  # (follows the timed steps: already warm; a handful of passes is enough for an average)
            last = args.warmup + min(args.steps, 5)
        else:
            run_steps(0, args.warmup)
        sampler = ClockSampler(local_rank)
        sampler.start()
        c0 = [dev.counters() for dev in devs]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for s_ in streams:
            s_.wait_event(e0)
        st = run_steps(args.warmup, last)
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            stream.wait_event(ev)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        c1 = [dev.counters() for dev in devs]
        st["ms"] = max_over_ranks(e0.elapsed_time(e1))
        st["wall_s"] = max_over_ranks(wall)
        st["clocks"] = sampler.summary()
        for k in c0[0]:
            st[k] = sum(b_[k] - a_[k] for a_, b_ in zip(c0, c1))
        st["total_cells"] = sum_over_ranks(st["cells"])
        st["total_hits"] = sum_over_ranks(float(st["hits"]))
        return st

    # ---- phase 1: reads resident in HBM ("value", weak scaling: 96 reads per GPU per step) ----
    R = args.reads_per_step * world
    reads = make_reads(args.seed, 0, nsteps_total * R, L, sizes, pool)
    v = run_legs(R, reads, host_buffers=False)
    # ---- phase 1b: the score kernels alone, one context at a time (roofline: no concurrent work in their time) ----
    kern = run_legs(R, reads, host_buffers=False, kernels_only=True)
    # ---- phase 2: the same steps through the C ABI with host buffers ("e2e_cabi") ----
    c = run_legs(R, reads, host_buffers=True)
    # ---- phase 3 (N > 1): strong scaling, a fixed batch of 96 reads over the N shards ----
    strong = None
    if world > 1:
        sr = run_legs(args.reads_per_step, reads[: nsteps_total * args.reads_per_step], host_buffers=False)
        strong = {"value": sr["total_cells"] / (sr["ms"] * 1e-3) / 1e9, "unit": "GCUPS", "reads_per_step": args.reads_per_step,
                  "ms_per_step": sr["ms"] / args.steps, "scaling": "strong"}
    # roofline inputs before the context goes away
    peaks = {}
    if rank == 0:
        for name, mode in (("fadd_vimnmx3_2to1", 6), ("fadd_fmnmx3_2to1", 5), ("fadd_fmnmx_1to1", 0), ("fadd", 1)):
            try:
                peaks[name] = devs[0].alu_peak(mode)
            except Exception:
                peaks[name] = None
    db_bytes = sum(dev.profile_bytes for dev in devs)
    for dev in devs:
        dev.close()
    torch.cuda.empty_cache()
    score_ms_max = max_over_ranks(kern["score_ms"]) * args.steps / min(args.steps, 5)

    # ---- phase 4: end to end through the reference API ("e2e"), one process driving all N GPUs ----
    plugin = None
    small = {}
    if world > 1:
        dist.barrier(group=cpu_group)  # every rank has released its GPU memory
    if rank == 0 and not args.no_plugin:
        if world == 1:  # (before the 24-GB database file is written: its write-back would stall their small files)
            try:
                small = small_config_legs(args, pool)
            except Exception as e:
                small = {"unavailable": f"{type(e).__name__}: {e}"}
        try:
            plugin = plugin_leg(args, world, sizes, pool, nodes_of, reads, R, nsteps_total)
        except Exception as e:  # report, do not lose the measured legs
            plugin = {"unavailable": f"{type(e).__name__}: {e}"}
    if world > 1:
        dist.barrier(group=cpu_group)

    if rank == 0:
        gcups = v["total_cells"] / (v["ms"] * 1e-3) / 1e9
        clocks = v["clocks"]
        # roofline of the dominant kernels (score pass of the first windows): 33 fp32 ops per cell over their
        # device time, against the issue ceiling (one lane-op per lane and clock at the sampled SM clock)
        mhz = clocks.get("sm_mhz") or 1965.0
        nominal = 148 * 128 * mhz * 1e6 / 1e12
        achieved = OPS_PER_CELL * kern["grid_cells"] / (kern["score_ms"] * 1e-3) / 1e12
        traffic_bytes = measured_traffic() if (world == 1 and args.profiles == 20000 and args.reads_per_step == 96
                                               and args.read_len == 2000) else None
        mix = peaks.get("fadd_vimnmx3_2to1")
        e2e_cabi = {"value": c["total_cells"] / c["wall_s"] / 1e9, "unit": "GCUPS",
                    "h2d_bytes_per_step": int(c["h2d_bytes"] / args.steps), "d2h_bytes_per_step": int(c["d2h_bytes"] / args.steps),
                    "reads_per_s": args.steps * R / c["wall_s"],
                    "api": "dcpgpu_reads_set + score_grid + scores_fetch + hits_fetch + trace_pairs + match_build (libdcpgpu.so), "
                           "rank 0's bytes, wall clock, max over ranks"}
        if plugin and "value" in plugin:
            e2e = dict(plugin)
            if plugin["profiles"] == len(sizes):
                e2e["same_cells_as_value"] = bool(plugin["cells"] == v["total_cells"])
                e2e["same_lrt_windows_as_value"] = bool(plugin["lrt_windows"] == int(v["total_hits"]))
        else:
            e2e = dict(e2e_cabi)
            e2e["note"] = "dcp_scan_run leg " + (plugin or {}).get("unavailable", "skipped (--no-plugin)") + "; this is the C ABI leg"
        line = {
            "metric": "GCUPS (nt x node DP cells/s), Pfam-scale scan", "value": gcups, "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v["ms"] / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, sizes, [int(x) for x in cuts]),
            "reads_per_s": args.steps * R / (v["ms"] * 1e-3),
            "clocks": clocks,
            "e2e": e2e,
            "e2e_cabi": e2e_cabi,
            "gpu_launches": int(v["launches"]),
            "roofline": {"bound": "fp32-alu (non-tensor add/min issue; not hbm, not tensor)", "achieved": achieved,
                         "peak": nominal, "unit": "TFLOP/s", "frac": achieved / nominal, "traffic": traffic_bytes,
                         "peak_source": f"issue ceiling: 148 SMs x 128 lanes x {mhz:.0f} MHz (SM clock sampled under load), one "
                                        "fp32 add/min per lane and clock; MEASURED_PEAKS.json has no ALU entry",
                         "peak_measured_mix": mix, "frac_of_measured_mix": (achieved / mix) if mix else None,
                         "peak_measured_mix_source": "dcpgpu_alu_peak mode 6 on this GPU: FADD + three-input integer min 2:1 "
                                                     "(the row's own mix; a min3 counts as two of the 33 ops)",
                         "alu_probe": peaks,
                         "kernel": "score pass of the first windows = score_row_kernel<Q,SEG,MODE,false,STAGE> (row_kernel.cuh, "
                                   "profile-stationary CTAs, short-code rows + null/background table staged in shared memory by "
                                   "TMA): whole profiles of <= 256 nodes (SEG 32/16/8/4), 256-node segments + tail of larger ones "
                                   "(speculative B, exact redo by score_reg_kernel<Q,W>), generic_kernel<false> for the rest; rank 0",
                         "ops_per_cell": OPS_PER_CELL, "kernel_gcups": kern["grid_cells"] / (kern["score_ms"] * 1e-3) / 1e9,
                         "kernel_ms_per_step": kern["score_ms"] / min(args.steps, 5),
                         "kernel_timing": "CUDA events around the score pass of each context run alone (phase 1b), the same "
                                          "batches as the timed steps",
                         "hbm_gbs_measured": _measured_peaks().get("hbm_gbs")},
            "hits_per_step": v["hits"] / args.steps, "path_steps_per_step": v["steps"] / args.steps,
            "later_windows_per_step": v["later_windows"] / args.steps, "multi_window_profiles_rank0": multi_window,
            "db_build_s": t_build, "db_bytes_rank0": db_bytes, "score_ms_max_rank": score_ms_max / args.steps,
        }
        if strong:
            line["strong"] = strong
        if small:
            line["configs"] = small
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_record(args, sizes, pool, small)
            except Exception as e:  # the checker is optional for the measured arm
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
