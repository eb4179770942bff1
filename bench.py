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

Prints ONE JSON line (rank 0).  `value` = GCUPS with the reads already resident in HBM;
`e2e` = the same through the public API with HOST buffers (pack + H2D + D2H inside the timed
region); `roofline` = the score kernel against the measured FP32 non-tensor issue rate;
`cpu_baseline` / `--impl reference` = the reference's own viterbi.c/trellis.c (oracle/_ref,
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
        with open(os.path.join(ROOT, "profiles", "r1_traffic_v16.json")) as f:
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

def balanced_sample(sizes, cores, per_thread, seed):
    """K-stratified sample of per_thread x cores profiles, ordered so that the reference's
    contiguous count partitions (scan.c:188-208, partition_size.c:13-16: cores partitions of
    per_thread consecutive profiles) carry near-equal sum K -- what partitioning 20,000 profiles
    by count gives the real scan (1,250 profiles per thread average out), and what a sample of
    one or two profiles per thread does not."""
    nprof = per_thread * cores
    order = np.argsort(sizes, kind="stable")
    pick = order[np.linspace(0, len(order) - 1, nprof).round().astype(int)][::-1]  # largest first
    parts = [[] for _ in range(cores)]
    for r in range(per_thread):  # snake deal: round r hands one profile to every partition
        row = pick[r * cores:(r + 1) * cores]
        for j, p in enumerate(row if r % 2 == 0 else row[::-1]):
            parts[j].append(int(p))
    rng = np.random.default_rng(seed)
    for part in parts:
        rng.shuffle(part)  # mixed sizes inside a partition, like a real database
    return np.asarray([p for part in parts for p in part], dtype=np.int64)


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
    pick = balanced_sample(sizes, cores, per_thread, args.seed)
    profs = []
    for p in pick:
        ids, bmk = profile_nodes(args.seed, int(p), sizes[p], pool)
        tr, em = pool.trans[ids], pool.emission[ids]
        pr = Profile("s%d" % p, 1, "", int(sizes[p]), pool.null_emission, pool.bg_emission,
                     np.concatenate([tr, tr[-1:]]), np.concatenate([em, em[-1:]]), bmk)
        profs.append(ref.profile(pr.costs()))
    probe_reads = make_reads(args.seed, 0, 2, args.read_len, sizes, pool)
    ref.scan(profs, probe_reads, True, False, cores)  # first touch of the tables
    t = ref.scan(profs, probe_reads, True, False, cores)
    rate = t["cells"] / max(t["seconds"], 1e-9)
    per_read = t["cells"] / 2
    nreads = int(max(2, min(4000, round(target_seconds * rate / per_read))))
    reads = make_reads(args.seed, 0, nreads, args.read_len, sizes, pool)
    r = ref.scan(profs, reads, True, False, cores)
    ts = r["thread_seconds"]
    gcups = r["cells"] / r["seconds"] / 1e9
    sample = (f"{len(pick)} K-stratified profiles = {per_thread} per thread x {cores} threads, sum-K-balanced count "
              f"partitions (sum K = {int(sizes[pick].sum())}) x {nreads} reads of {args.read_len} nt, every window.c "
              f"window, xtrans + null + alt Viterbi per window, trellis+unzip for lrt>=0 ({r['hits']} hits), "
              f"{r['cells']:.3e} cells in {r['seconds']:.2f} s; per-thread busy {ts.min():.2f}..{ts.max():.2f} s, "
              f"parallel efficiency {r['parallel_efficiency']:.3f}, {gcups / cores:.4f} GCUPS per core; "
              f"{os.path.basename(ref.path)}")
    return {"gcups": gcups, "reads_per_s": None, "cores": cores, "sample": sample,
            "seconds": r["seconds"], "cells": r["cells"], "profs": profs, "ref": ref, "nreads": nreads,
            "sumK_sample": int(sizes[pick].sum()), "parallel_efficiency": r["parallel_efficiency"],
            "reads": reads}


def cpu_baseline_record(args, sizes, pool):
    """cpu_baseline of the measured arm: the widest SIMD build this host runs, plus the reference
    Makefile's default level (-mavx2, c-core/Makefile:10-12) on a third of the sample."""
    c = cpu_scan_sample(args, sizes, pool, args.cpu_seconds)
    rec = {"value": c["gcups"], "unit": "GCUPS", "cores": c["cores"], "kind": "reference", "sample": c["sample"],
           "parallel_efficiency": c["parallel_efficiency"]}
    wide = os.path.basename(c["ref"].path)
    c.clear()
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
        "parallelism": f"profile-sharded x{args.gpus}, no collective",
        "l2_policy": "inputs larger than L2 (profile tables >> 126 MB); no flush",
        "shard": shard,
    }


# ---- B200 leg -------------------------------------------------------------------------------

def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from deciphon_b200.device import PAIR_DTYPE, Device

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (deciphon_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line unless the caller asks for more
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pool = synth.NodePool()
    sizes = synth.core_sizes(np.random.default_rng(args.seed), args.profiles)
    cuts = shard_bounds(sizes, world)
    p0, p1 = cuts[rank], cuts[rank + 1]

    dev = Device(local_rank)
    stream = torch.cuda.current_stream()
    dev.set_stream(stream.cuda_stream)
    t_build = time.time()
    first = dev.pool_add(pool.emission, pool.trans)
    for p in range(p0, p1):
        ids, bmk = profile_nodes(args.seed, p, sizes[p], pool)
        dev.profile_add(int(sizes[p]), bmk, pool.null_emission, pool.bg_emission, ids + first)
    dev.sync()
    t_build = time.time() - t_build
    nprof = p1 - p0
    R, L = args.reads_per_step * world, args.read_len  # the batch grows with the GPUs: per-rank work fixed
    nsteps_total = args.warmup + args.steps
    reads = make_reads(args.seed, 0, nsteps_total * R, L, sizes, pool)
    win = np.minimum(np.minimum(sizes[p0:p1] * 50, 100000), L).astype(np.int32)  # first window per profile

    def hit_pairs(seq0):
        idx = dev.hits_fetch()
        pr = np.zeros(len(idx), dtype=PAIR_DTYPE)
        pr["profile"] = idx // R
        pr["seq"] = seq0 + idx % R
        pr["start"] = 0
        pr["len"] = win[idx // R]
        return pr

    def step_resident(i, stats):
        seq0 = i * R
        dev.score_grid(0, nprof, seq0, seq0 + R, True, False)
        pr = hit_pairs(seq0)
        stats["score_ms"] += dev.last_kernel_ms()
        stats["cells"] += dev.last_cells()
        if len(pr):
            _, off, _ids, _sz = dev.trace_pairs_flat(pr, True, False)
            stats["steps"] += int(off[-1])
        stats["hits"] += len(pr)

    # ---- phase 1: reads resident in HBM ("value") ----
    dev.set_reads(reads)
    stats = {"score_ms": 0.0, "cells": 0.0, "hits": 0, "steps": 0}
    for i in range(args.warmup):
        step_resident(i, stats)
    stats = {"score_ms": 0.0, "cells": 0.0, "hits": 0, "steps": 0}
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = dev.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.warmup, nsteps_total):
        step_resident(i, stats)
    e1.record(stream)
    barrier()
    elapsed_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.summary()
    launches = dev.launch_count() - launches0
    total_cells = sum_over_ranks(stats["cells"])
    score_ms_max = max_over_ranks(stats["score_ms"])

    # ---- phase 2: end to end through the public API with host buffers ("e2e") ----
    pinned = []
    for i in range(nsteps_total):
        sym = torch.from_numpy(np.concatenate(reads[i * R:(i + 1) * R])).pin_memory()
        pinned.append(sym)
    offs = np.arange(R + 1, dtype=np.int64) * L
    h2d = d2h = 0

    def step_e2e(i):
        nonlocal h2d, d2h
        dev.set_reads_packed(pinned[i].numpy(), offs)
        dev.score_grid(0, nprof, 0, R, True, False)
        nul, alt = dev.scores_fetch(nprof * R)
        pr = hit_pairs(0)
        nst = 0
        if len(pr):
            _, off, _ids, _sz = dev.trace_pairs_flat(pr, True, False)
            nst = int(off[-1])
        h2d = (R * L + 3) // 4 + R * 12 + len(pr) * 16
        d2h = nprof * R * 8 + len(pr) * 8 + nst * 3
        return float(nul[0]) + float(alt[-1])

    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, nsteps_total):
        step_e2e(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)

    if rank == 0:
        gcups = total_cells / (elapsed_ms * 1e-3) / 1e9
        e2e_gcups = total_cells / e2e_s / 1e9
        reads_per_s = args.steps * R / (elapsed_ms * 1e-3)
        # roofline of the dominant kernel (score pass): algorithmic fp32 ops / its measured device time
        try:
            peak = dev.alu_peak(0)
            peak_src = "measured here: FADD+FMNMX 1:1 issue-rate microbenchmark (dcpgpu_alu_peak mode 0)"
        except Exception:
            peak = 148 * 128 * 1.965e9 / 1e12
            peak_src = "fallback: nominal 148 SM x 128 lanes x 1.965 GHz"
        achieved = OPS_PER_CELL * stats["cells"] / (stats["score_ms"] * 1e-3) / 1e12
        # DRAM bytes of one step's score pass, ncu capture of this workload (default sizes, one GPU)
        traffic_bytes = measured_traffic() if (world == 1 and args.profiles == 20000 and args.reads_per_step == 96
                                               and args.read_len == 2000) else None
        line = {
            "metric": "GCUPS (nt x node DP cells/s), Pfam-scale scan", "value": gcups, "unit": "GCUPS",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, sizes, [int(c) for c in cuts]),
            "reads_per_s": reads_per_s,
            "clocks": clocks,
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "reads_per_s": args.steps * R / e2e_s},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32-alu (non-tensor add/min issue; not hbm, not tensor)", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic_bytes,
                         "kernel": "score pass = score_sub_kernel<Q,G> (K<=128) + score_reg_kernel<Q,1> (K<=256) + score_lstrip/subtail_kernel over 256-node segments (K<=2048, exact redo of failed speculation) + generic_kernel<false> (rest), rank 0",
                         "ops_per_cell": OPS_PER_CELL, "kernel_gcups": stats["cells"] / (stats["score_ms"] * 1e-3) / 1e9,
                         "kernel_ms_per_step": stats["score_ms"] / args.steps, "peak_source": peak_src,
                         "hbm_gbs_measured": _measured_peaks().get("hbm_gbs")},
            "hits_per_step": stats["hits"] / args.steps, "path_steps_per_step": stats["steps"] / args.steps,
            "db_build_s": t_build, "db_bytes_rank0": dev.profile_bytes, "score_ms_max_rank": score_ms_max / args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline_record(args, sizes, pool)
            except Exception as e:  # the checker is optional for the measured arm
                line["cpu_baseline"] = {"value": None, "unit": "GCUPS", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"unavailable: {e}"}
        print(json.dumps(line), flush=True)
    dev.close()
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
