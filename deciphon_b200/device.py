"""Host-side handle of one B200: resident profiles + packed reads + score/trace passes.

Thin object layer over the C ABI (include/dcpgpu.h).  It plays the role of the reference's
per-thread ``work``/``viterbi`` objects (c-core/work.c:24-51, thread.c:98-128), but for a
whole batch of (window, profile) pairs at once.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import HMMER3_COMPAT, KEEP_TRELLIS, MULTI_HITS, DcpGpuError, lib

PAIR_DTYPE = np.dtype([("profile", "<i4"), ("seq", "<i4"), ("start", "<i4"), ("len", "<i4")])


def flags_of(multi_hits: bool, hmmer3_compat: bool, keep_trellis: bool = False) -> int:
    return ((MULTI_HITS if multi_hits else 0) | (HMMER3_COMPAT if hmmer3_compat else 0)
            | (KEEP_TRELLIS if keep_trellis else 0))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Device:
    def __init__(self, index: int = 0):
        h = C.c_void_p()
        rc = lib.dcpgpu_open(C.byref(h), index)
        if rc:
            raise DcpGpuError(rc, lib.dcpgpu_strerror(rc).decode())
        self._h = h
        self.index = index

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc:
            detail = lib.dcpgpu_last_error(self._h).decode()
            raise DcpGpuError(rc, f"{lib.dcpgpu_strerror(rc).decode()} ({detail})")

    def close(self):
        if getattr(self, "_h", None):
            lib.dcpgpu_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream: int | None):
        self._check(lib.dcpgpu_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        self._check(lib.dcpgpu_sync(self._h))

    @property
    def sm_count(self) -> int:
        return int(lib.dcpgpu_device_info(self._h, 0))

    @property
    def profile_bytes(self) -> int:
        return int(lib.dcpgpu_device_info(self._h, 3))

    def free_bytes(self) -> int:
        return int(lib.dcpgpu_device_info(self._h, 2))

    # -- profiles ------------------------------------------------------------------------
    def pool_add(self, emission: np.ndarray, trans: np.ndarray) -> int:
        """Upload nodes in .dcp (log-prob) form; returns the id of the first one."""
        emission = np.ascontiguousarray(emission, dtype=np.float32)
        trans = np.ascontiguousarray(trans, dtype=np.float32)
        n = emission.shape[0]
        assert emission.shape == (n, 1364) and trans.shape == (n, 7)
        first = C.c_int64()
        self._check(lib.dcpgpu_pool_add(self._h, n, _ptr(emission), _ptr(trans), C.byref(first)))
        return first.value

    def pool_release(self):
        self._check(lib.dcpgpu_pool_release(self._h))

    def profile_add(self, K: int, BMk, null_emission, bg_emission, node_ids=None, first_node_id: int = 0) -> int:
        BMk = np.ascontiguousarray(BMk, dtype=np.float32)
        nul = np.ascontiguousarray(null_emission, dtype=np.float32)
        bg = np.ascontiguousarray(bg_emission, dtype=np.float32)
        ids = None if node_ids is None else np.ascontiguousarray(node_ids, dtype=np.int64)
        assert BMk.size == K and nul.size == 1364 and bg.size == 1364
        assert ids is None or ids.size == K
        idx = C.c_int32()
        self._check(lib.dcpgpu_profile_add(self._h, K, _ptr(ids), first_node_id, _ptr(BMk), _ptr(nul), _ptr(bg),
                                           C.byref(idx)))
        return idx.value

    def add_profile(self, prof) -> int:
        """Make a ``dcp_file.Profile`` resident (the GPU analogue of work_setup, work.c:24-46)."""
        K = prof.core_size
        first = self.pool_add(prof.emission[:K], prof.trans[:K])
        idx = self.profile_add(K, prof.BMk, prof.null_emission, prof.bg_emission, None, first)
        self.pool_release()
        return idx

    @property
    def num_profiles(self) -> int:
        return lib.dcpgpu_profile_count(self._h)

    def core_size(self, profile: int) -> int:
        return lib.dcpgpu_profile_core_size(self._h, profile)

    # -- reads ---------------------------------------------------------------------------
    def set_reads(self, reads):
        """reads: list of uint8 arrays of symbols 0..3 (A,C,G,T/U)."""
        off = np.zeros(len(reads) + 1, dtype=np.int64)
        if len(reads):
            off[1:] = np.cumsum([len(r) for r in reads])
            sym = np.ascontiguousarray(np.concatenate(reads), dtype=np.uint8)
        else:
            sym = np.zeros(1, dtype=np.uint8)
        self._check(lib.dcpgpu_reads_set(self._h, len(reads), _ptr(sym), _ptr(off)))

    def set_reads_packed(self, symbols: np.ndarray, offsets: np.ndarray):
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(lib.dcpgpu_reads_set(self._h, len(offsets) - 1, _ptr(symbols), _ptr(offsets)))

    # -- score pass ----------------------------------------------------------------------
    def score_pairs(self, pairs: np.ndarray, multi_hits=True, hmmer3_compat=False):
        """pairs: structured array of PAIR_DTYPE (or int32[n,4]).  Returns (null_cost, alt_cost)."""
        pairs = np.ascontiguousarray(pairs)
        if pairs.dtype != PAIR_DTYPE:
            pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 4).view(PAIR_DTYPE).reshape(-1)
        n = pairs.shape[0]
        nul = np.empty(n, dtype=np.float32)
        alt = np.empty(n, dtype=np.float32)
        self._check(lib.dcpgpu_score_pairs(self._h, n, _ptr(pairs), flags_of(multi_hits, hmmer3_compat),
                                           _ptr(nul), _ptr(alt)))
        return nul, alt

    def score_grid(self, prof0: int, prof1: int, seq0: int, seq1: int, multi_hits=True, hmmer3_compat=False):
        """First window of every seq in [seq0,seq1) x every profile in [prof0,prof1).  Asynchronous on
        the context's stream unless profiles of more than 256 nodes are present."""
        self._check(lib.dcpgpu_score_grid(self._h, prof0, prof1, seq0, seq1, flags_of(multi_hits, hmmer3_compat)))

    def scores_fetch(self, n: int):
        nul = np.empty(n, dtype=np.float32)
        alt = np.empty(n, dtype=np.float32)
        self._check(lib.dcpgpu_scores_fetch(self._h, n, _ptr(nul), _ptr(alt)))
        return nul, alt

    def hits_fetch(self) -> np.ndarray:
        n = C.c_int64()
        self._check(lib.dcpgpu_hits_fetch(self._h, 0, None, C.byref(n)))
        idx = np.empty(n.value, dtype=np.int64)
        if n.value:
            self._check(lib.dcpgpu_hits_fetch(self._h, n.value, _ptr(idx), C.byref(n)))
        return idx

    def scores_gather(self, index: np.ndarray):
        """(null_cost, alt_cost) of the given pairs of the last score pass (e.g. its hit list)."""
        index = np.ascontiguousarray(index, dtype=np.int64)
        nul = np.empty(len(index), dtype=np.float32)
        alt = np.empty(len(index), dtype=np.float32)
        self._check(lib.dcpgpu_scores_gather(self._h, len(index), _ptr(index), _ptr(nul), _ptr(alt)))
        return nul, alt

    def last_cells(self) -> float:
        return float(lib.dcpgpu_last_cells(self._h))

    def last_kernel_ms(self) -> float:
        return float(lib.dcpgpu_last_kernel_ms(self._h))

    def last_redo(self) -> int:
        return int(lib.dcpgpu_last_redo(self._h))

    def launch_count(self) -> int:
        return int(lib.dcpgpu_launch_count(self._h))

    def alu_peak(self, mode: int = 0) -> float:
        """Measured FP32 non-tensor issue rate, tera lane-ops/s (see include/dcpgpu.h)."""
        v = C.c_double()
        self._check(lib.dcpgpu_alu_peak(self._h, mode, C.byref(v)))
        return v.value

    def frame_tables(self, nuclt_lprobs, codon_marg_lprobs, epsilon: float) -> np.ndarray:
        """Emission log-prob tables [n][1364] of n frame states (press path, include/dcpgpu.h)."""
        nu = np.ascontiguousarray(nuclt_lprobs, dtype=np.float32).reshape(-1, 4)
        cm = np.ascontiguousarray(codon_marg_lprobs, dtype=np.float32).reshape(-1, 125)
        if len(nu) != len(cm):
            raise ValueError("one base distribution and one codon marginal table per state")
        out = np.empty((len(nu), 1364), dtype=np.float32)
        self._check(lib.dcpgpu_frame_tables(self._h, len(nu), _ptr(nu), _ptr(cm), float(epsilon), _ptr(out)))
        return out

    # -- trace pass ----------------------------------------------------------------------
    def trace_pairs_flat(self, pairs: np.ndarray, multi_hits=True, hmmer3_compat=False, keep_trellis=False):
        """Returns (alt_cost[n], offsets[n+1], state_ids uint16[total], seqsizes uint8[total])."""
        pairs = np.ascontiguousarray(pairs)
        if pairs.dtype != PAIR_DTYPE:
            pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 4).view(PAIR_DTYPE).reshape(-1)
        n = pairs.shape[0]
        alt = np.empty(n, dtype=np.float32)
        nsteps = np.zeros(n, dtype=np.int32)
        self._check(lib.dcpgpu_trace_pairs(self._h, n, _ptr(pairs),
                                           flags_of(multi_hits, hmmer3_compat, keep_trellis),
                                           _ptr(alt), _ptr(nsteps)))
        self._last_traced = n
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum(nsteps)
        ids = np.zeros(max(int(off[-1]), 1), dtype=np.uint16)
        sz = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        self._check(lib.dcpgpu_trace_fetch(self._h, _ptr(off), _ptr(ids), _ptr(sz)))
        return alt, off, ids, sz

    def trace_pairs(self, pairs: np.ndarray, multi_hits=True, hmmer3_compat=False, keep_trellis=False):
        """Returns (alt_cost[n], paths) with paths[i] = (state_ids uint16[], seqsizes uint8[]).
        keep_trellis=True also materialises the whole trellis (see trace_trellis)."""
        pairs = np.ascontiguousarray(pairs)
        if pairs.dtype != PAIR_DTYPE:
            pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 4).view(PAIR_DTYPE).reshape(-1)
        n = pairs.shape[0]
        alt = np.empty(n, dtype=np.float32)
        nsteps = np.zeros(n, dtype=np.int32)
        self._check(lib.dcpgpu_trace_pairs(self._h, n, _ptr(pairs),
                                           flags_of(multi_hits, hmmer3_compat, keep_trellis),
                                           _ptr(alt), _ptr(nsteps)))
        self._last_traced = n
        off = np.zeros(n + 1, dtype=np.int64)
        off[1:] = np.cumsum(nsteps)
        ids = np.zeros(max(int(off[-1]), 1), dtype=np.uint16)
        sz = np.zeros(max(int(off[-1]), 1), dtype=np.uint8)
        self._check(lib.dcpgpu_trace_fetch(self._h, _ptr(off), _ptr(ids), _ptr(sz)))
        paths = [(ids[off[i]:off[i + 1]].copy(), sz[off[i]:off[i + 1]].copy()) for i in range(n)]
        return alt, paths

    # -- post-processing of the traced paths on the device -----------------------------------
    def set_decoder(self, profile: int, node_dists, null_dist, bg_dist, gencode64: str):
        """Decode tables of a resident profile: node_dists f32[K,129], null/bg f32[129] (4 base
        log-probs + 125 codon marginals) and the 64 amino letters of its genetic code (TCAG order)."""
        nd = np.ascontiguousarray(node_dists, dtype=np.float32)
        nu = np.ascontiguousarray(null_dist, dtype=np.float32)
        bg = np.ascontiguousarray(bg_dist, dtype=np.float32)
        assert nd.shape[1] == 129 and nu.size == 129 and bg.size == 129 and len(gencode64) == 64
        self._check(lib.dcpgpu_profile_set_decoder(self._h, profile, _ptr(nd), _ptr(nu), _ptr(bg), gencode64.encode()))

    def match_build(self, epsilon: float = 0.01, is_rna: bool = False, want_text: bool = False):
        """Hit flag and window-relative extent [start, stop) of every pair of the preceding trace
        pass (thread.c:130-166); with want_text also the bytes of each row's match column."""
        n = self._last_traced
        hit = np.zeros(n, dtype=np.int32)
        hs = np.zeros(n, dtype=np.int32)
        he = np.zeros(n, dtype=np.int32)
        off = np.zeros(n + 1, dtype=np.int64)
        self._check(lib.dcpgpu_match_build(self._h, float(epsilon), int(is_rna), _ptr(hit), _ptr(hs), _ptr(he),
                                           _ptr(off) if want_text else None))
        if not want_text:
            return hit, hs, he
        text = np.zeros(int(off[-1]) + 1, dtype=np.uint8)
        self._check(lib.dcpgpu_match_fetch(self._h, _ptr(text)))
        raw = text.tobytes()
        return hit, hs, he, [raw[off[i]:off[i + 1]].decode() for i in range(n)]

    def counters(self) -> dict:
        """Cumulative H2D / D2H bytes, kernel launches and DP cells of this context."""
        names = ("h2d_bytes", "d2h_bytes", "launches", "cells")
        return {k: float(lib.dcpgpu_counter(self._h, i)) for i, k in enumerate(names)}

    def trace_trellis(self, i: int, length: int, K: int):
        xn = np.zeros(length + 1, dtype=np.uint32)
        nd = np.zeros((length + 1) * K, dtype=np.uint16)
        self._check(lib.dcpgpu_trace_trellis(self._h, i, _ptr(xn), _ptr(nd)))
        return xn, nd


def xtrans(window_len: int, multi_hits=True, hmmer3_compat=False) -> np.ndarray:
    """The 13 special-transition costs the device uses for a window (host computation)."""
    out = np.zeros(13, dtype=np.float32)
    rc = lib.dcpgpu_xtrans(window_len, flags_of(multi_hits, hmmer3_compat), _ptr(out))
    if rc:
        raise DcpGpuError(rc, lib.dcpgpu_strerror(rc).decode())
    return out
