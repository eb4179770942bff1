"""ctypes front-end for the oracle -- TEST INFRASTRUCTURE ONLY.

``Oracle``  wraps oracle/libdcporacle.so (our scalar restatement, dcp_oracle.c).
``Reference`` wraps oracle/_ref/libdcpref_*.so (the reference's own viterbi.c /
trellis.c compiled where they lie, see oracle/Makefile and ref_driver.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
X_NAMES = ("RR", "SN", "NN", "SB", "NB", "EB", "JB", "EJ", "JJ", "EC", "CC", "ET", "CT")

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(dtype=np.uint16, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def encode(seq: str) -> np.ndarray:
    """ACGT(U) -> 0..3 (alphabet order of the golden header)."""
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate("ACGT"):
        lut[ord(ch)] = i
    lut[ord("U")] = 3
    a = lut[np.frombuffer(seq.upper().encode(), dtype=np.uint8)]
    if (a == 255).any():
        raise ValueError("sequence holds symbols outside ACGTU; disambiguate first")
    return np.ascontiguousarray(a)


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "libdcporacle.so")
        if not os.path.exists(path):
            build()
        L = self.lib = C.CDLL(path)
        L.orc_xtrans.argtypes = [C.c_int, C.c_int, C.c_int, f32p]
        L.orc_code.argtypes = [u8p, C.c_int, C.c_int]
        L.orc_core_costs.argtypes = [C.c_int, f32p, f32p, f32p]
        L.orc_null.argtypes = [f32p, f32p, u8p, C.c_int]
        L.orc_null.restype = C.c_float
        L.orc_alt.argtypes = [C.c_int, f32p, f32p, f32p, f32p, f32p, u8p, C.c_int]
        L.orc_alt.restype = C.c_float
        L.orc_trace.argtypes = [C.c_int, f32p, f32p, f32p, f32p, f32p, u8p, C.c_int, u32p, u16p]
        L.orc_trace.restype = C.c_float
        L.orc_unzip.argtypes = [C.c_int, C.c_int, u32p, u16p, u16p, u8p, C.c_int]
        L.orc_lrt.argtypes = [C.c_float, C.c_float]
        L.orc_lrt.restype = C.c_float
        L.orc_hit_extent.argtypes = [C.c_int, u16p, u8p] + [C.POINTER(C.c_int)] * 4
        L.orc_window_next.argtypes = [C.POINTER(C.c_int * 4), C.c_int, C.c_int]
        L.orc_state_name.argtypes = [C.c_int, C.c_char_p]

    def xtrans(self, window_len: int, multi_hits: bool, hmmer3_compat: bool) -> np.ndarray:
        out = np.empty(13, dtype=np.float32)
        self.lib.orc_xtrans(window_len, int(multi_hits), int(hmmer3_compat), out)
        return out

    def core_costs(self, K, BMk, trans) -> np.ndarray:
        out = np.empty((8, K), dtype=np.float32)
        self.lib.orc_core_costs(K, np.ascontiguousarray(BMk, np.float32),
                                np.ascontiguousarray(trans, np.float32).reshape(-1), out.reshape(-1))
        return out

    def null(self, nul, xt, x) -> np.float32:
        return np.float32(self.lib.orc_null(nul, xt, x, len(x)))

    def alt(self, costs, xt, x) -> np.float32:
        nul, bg, em, core = costs
        return np.float32(self.lib.orc_alt(em.shape[0], nul, bg, em.reshape(-1), core.reshape(-1), xt, x, len(x)))

    def trace(self, costs, xt, x):
        nul, bg, em, core = costs
        K, L = em.shape[0], len(x)
        xn = np.zeros(L + 1, dtype=np.uint32)
        nd = np.zeros((L + 1) * K, dtype=np.uint16)
        alt = self.lib.orc_trace(K, nul, bg, em.reshape(-1), core.reshape(-1), xt, x, L, xn, nd)
        return np.float32(alt), xn, nd

    def unzip(self, K, L, xnodes, nodes):
        cap = L + 2 * K + 64
        ids = np.zeros(cap, dtype=np.uint16)
        sz = np.zeros(cap, dtype=np.uint8)
        n = self.lib.orc_unzip(K, L, xnodes, nodes, ids, sz, cap)
        if n < 0:
            raise RuntimeError(f"orc_unzip failed ({n})")
        return ids[:n].copy(), sz[:n].copy()

    def path(self, costs, xt, x):
        _, xn, nd = self.trace(costs, xt, x)
        return self.unzip(costs[2].shape[0], len(x), xn, nd)

    def lrt(self, null_cost, alt_cost) -> np.float32:
        return np.float32(self.lib.orc_lrt(float(null_cost), float(alt_cost)))

    def hit_extent(self, ids, sizes):
        v = [C.c_int() for _ in range(4)]
        ok = self.lib.orc_hit_extent(len(ids), np.ascontiguousarray(ids), np.ascontiguousarray(sizes),
                                     *[C.byref(a) for a in v])
        return (v[0].value, v[1].value, v[2].value, v[3].value) if ok else None

    def windows(self, seq_len: int, core_size: int, last_hit_pos=None):
        """Generator over (idx, start, stop); send() the window-relative last hit position."""
        st = (C.c_int * 4)(-1, 0, -1, -1)
        while self.lib.orc_window_next(C.byref(st), seq_len, core_size):
            lhp = yield (st[2], st[0], st[1])
            if lhp is not None:
                st[3] = lhp

    def state_name(self, sid: int) -> str:
        buf = C.create_string_buffer(16)
        if self.lib.orc_state_name(int(sid), buf):
            raise ValueError("invalid state id")
        return buf.value.decode()


def ref_lib_path() -> str | None:
    """Pick the widest reference build this CPU can run (AVX-512 else AVX2)."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    want = os.environ.get("DCP_REF_SIMD", "")
    names = []
    if want != "avx2" and " avx512f" in flags:
        names.append("libdcpref_avx512.so")
    names.append("libdcpref_avx2.so")
    if want == "avx2":
        names = ["libdcpref_avx2.so"]
    for n in names:
        p = os.path.join(HERE, "_ref", n)
        if os.path.exists(p):
            return p
    return None


def ref_xtrans(window_lens, multi_hits: bool, hmmer3_compat: bool) -> np.ndarray:
    """The 13 special-transition costs the REFERENCE loads for windows of the given lengths:
    its own xtrans.c (oracle/_ref/libdcpref_xtrans.so), called like thread.c:112 with
    seq_size = max(L / 3, 1).  Returns float32 [n][13] in viterbi.h:4-19 order."""
    path = os.path.join(HERE, "_ref", "libdcpref_xtrans.so")
    if not os.path.exists(path) and os.path.isdir("/root/reference/c-core"):
        build()
    lib = C.CDLL(path)
    lens = np.atleast_1d(np.asarray(window_lens, dtype=np.int64))
    ss = np.ascontiguousarray(np.maximum(lens // 3, 1).astype(np.int32))
    out = np.empty((len(ss), 13), dtype=np.float32)
    lib.ref_xtrans_many.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.ref_xtrans_many(len(ss), ss.ctypes.data, int(multi_hits), int(hmmer3_compat), out.ctypes.data)
    return out


class Reference:
    """The reference's own viterbi.c/trellis.c/xtrans.c (oracle/_ref)."""

    def __init__(self, path: str | None = None):
        path = path or ref_lib_path()
        if path is None:
            if os.path.isdir("/root/reference/c-core"):
                build()
                path = ref_lib_path()
        if path is None:
            raise FileNotFoundError("oracle/_ref is not built and /root/reference is absent")
        self.path = path
        L = self.lib = C.CDLL(path)
        L.ref_profile_new.argtypes = [C.c_int, f32p, f32p, f32p, f32p]
        L.ref_profile_new.restype = C.c_void_p
        L.ref_profile_del.argtypes = [C.c_void_p]
        L.ref_set_xtrans.argtypes = [C.c_void_p, f32p]
        L.ref_null.argtypes = [C.c_void_p, u8p, C.c_int]
        L.ref_null.restype = C.c_float
        L.ref_cost.argtypes = [C.c_void_p, u8p, C.c_int]
        L.ref_cost.restype = C.c_float
        L.ref_path.argtypes = [C.c_void_p, u8p, C.c_int, u16p, u8p, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_scan.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, u8p, i64p, C.c_int, C.c_int,
                               C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                               C.POINTER(C.c_int64), C.c_void_p]
        L.ref_scan.restype = C.c_double
        L.ref_num_lanes.restype = C.c_int
        self.lanes = L.ref_num_lanes()

    def profile(self, costs):
        nul, bg, em, core = costs
        h = self.lib.ref_profile_new(em.shape[0], nul, bg, em.reshape(-1), core.reshape(-1))
        if not h:
            raise MemoryError
        return RefProfile(self, h, em.shape[0])

    def scan(self, profiles, reads, multi_hits=True, hmmer3_compat=False, nthreads=1, want_scores=False):
        """Reference CPU scan over all (profile, read) pairs; returns dict with seconds, cells, hits."""
        n = len(profiles)
        arr = (C.c_void_p * n)(*[p.h for p in profiles])
        off = np.zeros(len(reads) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(r) for r in reads])
        x = np.ascontiguousarray(np.concatenate(reads) if len(reads) else np.zeros(0, np.uint8))
        cells = C.c_double()
        hits = C.c_int64()
        nul = alt = None
        pn = pa = None
        if want_scores:
            nul = np.zeros((n, len(reads)), dtype=np.float32)
            alt = np.zeros((n, len(reads)), dtype=np.float32)
            pn, pa = nul.ctypes.data, alt.ctypes.data
        nthreads = max(1, min(int(nthreads), n))
        tsec = np.zeros(nthreads, dtype=np.float64)
        sec = self.lib.ref_scan(n, arr, len(reads), x, off, int(multi_hits), int(hmmer3_compat),
                                nthreads, pn, pa, C.byref(cells), C.byref(hits), tsec.ctypes.data)
        return {"seconds": sec, "cells": cells.value, "hits": hits.value, "null": nul, "alt": alt,
                "thread_seconds": tsec, "threads": nthreads,
                "parallel_efficiency": float(tsec.sum() / (nthreads * max(sec, 1e-12)))}


class RefProfile:
    def __init__(self, ref: Reference, h, K):
        self.ref, self.h, self.K = ref, h, K

    def __del__(self):
        try:
            self.ref.lib.ref_profile_del(self.h)
        except Exception:
            pass

    def set_xtrans(self, xt):
        self.ref.lib.ref_set_xtrans(self.h, xt)

    def null(self, x):
        return np.float32(self.ref.lib.ref_null(self.h, x, len(x)))

    def cost(self, x):
        return np.float32(self.ref.lib.ref_cost(self.h, x, len(x)))

    def path(self, x, want_trellis=False):
        L = len(x)
        cap = L + 2 * self.K + 64
        ids = np.zeros(cap, dtype=np.uint16)
        sz = np.zeros(cap, dtype=np.uint8)
        xn = nd = None
        pxn = pnd = None
        if want_trellis:
            xn = np.zeros(L + 1, dtype=np.uint32)
            nd = np.zeros((L + 1) * self.K, dtype=np.uint16)
            pxn, pnd = xn.ctypes.data, nd.ctypes.data
        n = self.ref.lib.ref_path(self.h, x, L, ids, sz, cap, pxn, pnd)
        if n < 0:
            raise RuntimeError(f"ref_path failed ({n})")
        if want_trellis:
            return ids[:n].copy(), sz[:n].copy(), xn, nd
        return ids[:n].copy(), sz[:n].copy()
