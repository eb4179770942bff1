"""Window waves over the C ABI: the per-(sequence, profile) window loop of c-core/thread.c:49-86
and window.c:13-37, run as batched GPU passes from Python (the C++ twin lives in
host/deciphon_b200.cpp:run_shard and is what dcp_scan_run uses).

Wave 0 is the first window of every pair (``Device.score_grid``); pairs whose sequence is longer
than that window keep producing windows, each a function of the previous one and of the position
of the last hit found in it (``window_set_last_hit_position``, thread.c:162).
"""
from __future__ import annotations

import numpy as np

from .device import PAIR_DTYPE, Device

MAX_WINDOW = 100000  # window.c:29


def window_next(start, stop, last_hit, seq_len, K):
    """window.c:13-37 on arrays: returns (alive, new_start, new_stop)."""
    alive = stop != seq_len
    start_miss = np.maximum(start + 1, start + last_hit + 1)
    start_miss = np.maximum(start_miss, stop + 1 - 4 * K)
    new_stop = np.minimum(start_miss + np.minimum(50 * K, MAX_WINDOW), seq_len)
    return alive, start_miss, new_stop


def first_windows(Ks: np.ndarray, seq_lens: np.ndarray):
    """[P, S] length of wave 0's windows."""
    return np.minimum(np.minimum(Ks * 50, MAX_WINDOW)[:, None], seq_lens[None, :])


def trace_hits(dev: Device, pairs: np.ndarray, multi_hits: bool, hmmer3_compat: bool, budget: float = 16e9,
               Ks: np.ndarray | None = None):
    """Trace pass + on-device hit extents of `pairs` in rounds bounded by the value-dump bytes.
    Returns (hit flag, hit_start, hit_stop, number of path steps)."""
    n = len(pairs)
    hit = np.zeros(n, dtype=np.int32)
    hs = np.zeros(n, dtype=np.int32)
    he = np.zeros(n, dtype=np.int32)
    steps = 0
    if n == 0:
        return hit, hs, he, steps
    cost = (pairs["len"].astype(np.float64) + 1) * (12.0 * (Ks[pairs["profile"]] if Ks is not None else 2048) + 32.0)
    i0 = 0
    while i0 < n:
        i1 = i0 + 1
        acc = cost[i0]
        while i1 < n and acc + cost[i1] <= budget:
            acc += cost[i1]
            i1 += 1
        _, off, _ids, _sz = dev.trace_pairs_flat(pairs[i0:i1], multi_hits, hmmer3_compat)
        steps += int(off[-1])
        hit[i0:i1], hs[i0:i1], he[i0:i1] = dev.match_build()
        i0 = i1
    return hit, hs, he, steps


def later_waves(dev: Device, Ks: np.ndarray, seq0: int, seq_lens: np.ndarray, hit_pairs0: np.ndarray,
                hit0: np.ndarray, hstop0: np.ndarray, multi_hits: bool = True, hmmer3_compat: bool = False):
    """Waves 1.. of the pairs whose sequence outlasts its first window.

    Ks: core sizes of the resident profiles [P]; sequences seq0 .. seq0+S of the resident reads
    with lengths seq_lens [S]; hit_pairs0 / hit0 / hstop0: wave 0's traced pairs with their hit
    flags and window-relative hit stops.  Returns a dict of counters (pairs, hits, path steps,
    waves); the DP cells are accumulated by the device (``Device.counters``)."""
    Ks = np.asarray(Ks, dtype=np.int64)
    seq_lens = np.asarray(seq_lens, dtype=np.int64)
    w0 = first_windows(Ks, seq_lens)
    pi, si = np.nonzero(w0 < seq_lens[None, :])
    out = {"pairs": 0, "hits": 0, "steps": 0, "waves": 0}
    if len(pi) == 0:
        return out
    start = np.zeros(len(pi), dtype=np.int64)
    stop = w0[pi, si].astype(np.int64)
    last_hit = np.full(len(pi), -1, dtype=np.int64)
    if len(hit_pairs0):  # thread.c:162 for wave 0's hits
        key = pi * len(seq_lens) + si
        hk = hit_pairs0["profile"].astype(np.int64) * len(seq_lens) + (hit_pairs0["seq"].astype(np.int64) - seq0)
        sel = np.nonzero(hit0 != 0)[0]
        pos = np.searchsorted(key, hk[sel])
        ok = (pos < len(key)) & (key[np.minimum(pos, len(key) - 1)] == hk[sel])
        last_hit[pos[ok]] = hstop0[sel[ok]] - 1
    while len(pi):
        alive, nstart, nstop = window_next(start, stop, last_hit, seq_lens[si], Ks[pi])
        pi, si, start, stop, last_hit = pi[alive], si[alive], nstart[alive], nstop[alive], last_hit[alive]
        if len(pi) == 0:
            break
        pairs = np.zeros(len(pi), dtype=PAIR_DTYPE)
        pairs["profile"], pairs["seq"], pairs["start"], pairs["len"] = pi, seq0 + si, start, stop - start
        nul, alt = dev.score_pairs(pairs, multi_hits, hmmer3_compat)
        with np.errstate(invalid="ignore", over="ignore"):
            lrt = np.float32(-2) * ((-nul) - (-alt))  # lrt.h:6-9
        gate = np.isfinite(lrt) & (lrt >= 0)  # thread.c:121
        idx = np.nonzero(gate)[0]
        hit, _hs, he, steps = trace_hits(dev, pairs[idx], multi_hits, hmmer3_compat, Ks=Ks)
        got = idx[hit != 0]
        last_hit[got] = he[hit != 0] - 1
        out["pairs"] += len(pairs)
        out["hits"] += len(idx)
        out["steps"] += steps
        out["waves"] += 1
    return out
