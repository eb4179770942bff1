"""Seeded synthetic workloads for tests and bench.py (SURVEY 8(d) recipes).

Profiles are bootstrapped from a pool of REAL profile nodes (the 576 nodes of the golden
minifam database, shipped as tests/golden/minifam.npz): every synthetic node copies one real
node's 1364-entry emission table and its 7 transitions, so the value distribution of the
tables is the real one; the B->M_k entry costs follow the occupancy recipe of the reference
(c-core/model.c:284-309).  Reads are uniform ACGT, optionally with a back-translated profile
consensus embedded, then passed through an error channel (per base 1/3 substitution,
1/3 insertion, 1/3 deletion).
"""
from __future__ import annotations

import os

import numpy as np

from .dcp_file import Profile

GOLDEN_NPZ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                          "minifam.npz")


def load_golden_profiles(path: str = GOLDEN_NPZ):
    """The three golden minifam profiles as ``Profile`` objects (from the committed fixture)."""
    z = np.load(path, allow_pickle=False)
    out = []
    for i in range(int(z["num_profiles"])):
        K = int(z[f"p{i}_core_size"])
        out.append(Profile(
            accession=str(z[f"p{i}_accession"]), gencode=int(z[f"p{i}_gencode"]),
            consensus=str(z[f"p{i}_consensus"]), core_size=K,
            null_emission=z[f"p{i}_null_emission"], bg_emission=z[f"p{i}_bg_emission"],
            trans=z[f"p{i}_trans"], emission=z[f"p{i}_emission"], BMk=z[f"p{i}_BMk"],
            null_nuclt=(z[f"p{i}_null_nuclt4"], z[f"p{i}_null_nuclt125"]),
            bg_nuclt=(z[f"p{i}_bg_nuclt4"], z[f"p{i}_bg_nuclt125"]),
            node_nuclt=(z[f"p{i}_node_nuclt4"], z[f"p{i}_node_nuclt125"])))
    return out


class NodePool:
    """Real nodes to bootstrap from: emission[n,1364], trans[n,7] (log-probs)."""

    def __init__(self, profiles=None):
        profiles = profiles or load_golden_profiles()
        # interior nodes only: the last node's MD/DD are -inf by construction
        self.emission = np.ascontiguousarray(np.concatenate([p.emission[: p.core_size - 1] for p in profiles]))
        self.trans = np.ascontiguousarray(np.concatenate([p.trans[: p.core_size - 1] for p in profiles]))
        self.null_emission = profiles[0].null_emission
        self.bg_emission = profiles[0].bg_emission
        # decode tables of the same nodes (nuclt_dist of a .dcp record: 4 base + 125 codon marginal log-probs)
        self.nuclt4 = np.ascontiguousarray(np.concatenate([p.node_nuclt[0][: p.core_size - 1] for p in profiles]))
        self.nuclt125 = np.ascontiguousarray(np.concatenate([p.node_nuclt[1][: p.core_size - 1] for p in profiles]))
        self.null_nuclt = profiles[0].null_nuclt
        self.bg_nuclt = profiles[0].bg_nuclt
        # most likely codon of every node: argmax over the 64 3-mers (codes 20..83)
        self.codon = np.argmax(self.emission[:, 20:84], axis=1).astype(np.int64)

    def __len__(self):
        return self.emission.shape[0]


def occupancy_entry(trans: np.ndarray) -> np.ndarray:
    """B->M_k log-probs from match occupancy (model.c:284-309 recipe, float64 then f32)."""
    K = trans.shape[0]
    t = trans.astype(np.float64)
    locc = np.empty(K)
    locc[0] = np.logaddexp(t[0, 1], t[0, 0])
    for i in range(1, K):
        v0 = locc[i - 1] + np.logaddexp(t[i, 0], t[i, 1])
        v1 = np.log1p(-min(np.exp(locc[i - 1]), 1.0 - 1e-12)) + t[i, 5]
        locc[i] = np.logaddexp(v0, v1)
    logz = np.logaddexp.reduce(locc + np.log(K - np.arange(K)))
    return (locc - logz).astype(np.float32)


def synth_profile_nodes(rng: np.random.Generator, K: int, pool: NodePool):
    """Node ids into the pool + entry log-probs for one synthetic profile of K nodes."""
    ids = rng.integers(0, len(pool), size=K)
    return ids.astype(np.int64), occupancy_entry(pool.trans[ids])


def synth_profile(rng: np.random.Generator, K: int, pool: NodePool, name: str = "SYN") -> Profile:
    ids, bmk = synth_profile_nodes(rng, K, pool)
    em = pool.emission[ids]
    tr = pool.trans[ids]
    n4, n125 = pool.nuclt4[ids], pool.nuclt125[ids]
    return Profile(accession=name, gencode=1, consensus="x" * K, core_size=K,
                   null_emission=pool.null_emission, bg_emission=pool.bg_emission,
                   trans=np.concatenate([tr, tr[-1:]]), emission=np.concatenate([em, em[-1:]]), BMk=bmk,
                   null_nuclt=pool.null_nuclt, bg_nuclt=pool.bg_nuclt,
                   node_nuclt=(np.concatenate([n4, n4[-1:]]), np.concatenate([n125, n125[-1:]])))


def core_sizes(rng: np.random.Generator, n: int, mean: float = 200.0, sigma: float = 0.7, lo: int = 20,
               hi: int = 2000) -> np.ndarray:
    """Clipped log-normal core sizes with the requested mean (Pfam-like)."""
    mu = np.log(mean) - 0.5 * sigma * sigma
    return np.clip(np.rint(rng.lognormal(mu, sigma, size=n)), lo, hi).astype(np.int64)


def random_read(rng: np.random.Generator, L: int) -> np.ndarray:
    return rng.integers(0, 4, size=L, dtype=np.uint8)


def mutate(rng: np.random.Generator, x: np.ndarray, rate: float) -> np.ndarray:
    """Error channel: each base is hit with probability `rate`; 1/3 sub, 1/3 ins, 1/3 del."""
    n = len(x)
    if n == 0:
        return np.zeros(1, dtype=np.uint8)
    hit = rng.random(n) < rate
    kind = rng.integers(0, 3, size=n)
    rnd = rng.integers(0, 4, size=n, dtype=np.uint8)
    sub = hit & (kind == 0)
    ins = hit & (kind == 1)
    dele = hit & (kind == 2)
    base = np.where(sub, (x + 1 + rnd % 3) % 4, x).astype(np.uint8)
    counts = np.where(dele, 0, np.where(ins, 2, 1))
    out = np.repeat(base, counts)
    # the first copy of an inserted position becomes the random base
    first = np.cumsum(counts) - counts
    out[first[ins]] = rnd[ins]
    if out.size == 0:
        out = x[:1].copy()
    return np.ascontiguousarray(out, dtype=np.uint8)


def consensus_dna(pool: NodePool, node_ids: np.ndarray) -> np.ndarray:
    """Back-translation of a synthetic profile: most likely codon of each node."""
    c = pool.codon[node_ids]
    return np.stack([c // 16, (c // 4) % 4, c % 4], axis=1).reshape(-1).astype(np.uint8)


def fixed_length(rng: np.random.Generator, x: np.ndarray, L: int) -> np.ndarray:
    """Trim or pad with random bases to exactly L nucleotides."""
    if len(x) >= L:
        return np.ascontiguousarray(x[:L])
    return np.concatenate([x, random_read(rng, L - len(x))])


# ---- a synthetic database as a .dcp file (bench.py's plugin leg, tests) ----------------------

def _node_record_dtype():
    """One node of a protein record in the current encoding (dcp_file.write_dcp, protein.c:262-275):
    "nuclt_dist": [bin f32[4], bin f32[125]], "trans": bin f32[7], "emission": bin f32[1364] --
    6,037 bytes with constant framing, so a profile's node block is one gather from the pool."""
    from .dcp_file import _mp_arr, _mp_str
    h0 = _mp_str("nuclt_dist") + _mp_arr(2) + bytes([0xC4, 16])
    h1 = bytes([0xC5, 0x01, 0xF4])  # bin16, 500 bytes
    h2 = _mp_str("trans") + bytes([0xC4, 28])
    h3 = _mp_str("emission") + bytes([0xC5, 0x15, 0x50])  # bin16, 5456 bytes
    names, formats, offsets, o = [], [], [], 0
    for name, fmt, size in (("h0", f"S{len(h0)}", len(h0)), ("n4", ("<f4", 4), 16), ("h1", "S3", 3),
                            ("n125", ("<f4", 125), 500), ("h2", f"S{len(h2)}", len(h2)), ("tr", ("<f4", 7), 28),
                            ("h3", f"S{len(h3)}", len(h3)), ("em", ("<f4", 1364), 5456)):
        names.append(name)
        formats.append(fmt)
        offsets.append(o)
        o += size
    dt = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": o})
    return dt, (h0, h1, h2, h3)


def pool_node_records(pool: NodePool) -> np.ndarray:
    dt, (h0, h1, h2, h3) = _node_record_dtype()
    rec = np.zeros(len(pool), dtype=dt)
    # numpy "S" fields strip trailing NULs on read but store the bytes given: write through a byte view
    raw = rec.view(np.uint8).reshape(len(pool), dt.itemsize)
    for name, h in (("h0", h0), ("h1", h1), ("h2", h2), ("h3", h3)):
        o = dt.fields[name][1]
        raw[:, o:o + len(h)] = np.frombuffer(h, dtype=np.uint8)
    rec["n4"], rec["n125"], rec["tr"], rec["em"] = pool.nuclt4, pool.nuclt125, pool.trans, pool.emission
    return rec


def write_synth_dcp(path: str, sizes, pool: NodePool, nodes_of, epsilon: float = 0.01, first: int = 0,
                    count: int | None = None) -> dict:
    """Stream profiles first .. first+count of a synthetic database into a .dcp file, byte for byte
    what dcp_file.write_dcp writes for the same Profile objects (tests/test_host_scan.py checks).
    nodes_of(p) -> (pool node ids [K], BMk [K]) of profile p."""
    import struct

    from .dcp_file import MAGIC_NUMBER, _mp_arr, _mp_f32bin, _mp_int, _mp_map, _mp_str
    count = len(sizes) - first if count is None else count
    rec = pool_node_records(pool)
    node_bytes = rec.dtype.itemsize
    raw = rec.view(np.uint8).reshape(len(pool), node_bytes)
    nuc = lambda d: _mp_arr(2) + _mp_f32bin(d[0]) + _mp_f32bin(d[1])  # noqa: E731
    fixed = (_mp_str("null_nuclt_dist") + nuc(pool.null_nuclt) + _mp_str("null_emission") + _mp_f32bin(pool.null_emission)
             + _mp_str("bg_nuclt_dist") + nuc(pool.bg_nuclt) + _mp_str("bg_emission") + _mp_f32bin(pool.bg_emission))

    def head(p, K):
        return (_mp_map(10) + _mp_str("accession") + _mp_str("SYN%05d" % p) + _mp_str("gencode") + _mp_int(1)
                + _mp_str("consensus") + _mp_str("x" * K) + _mp_str("core_size") + _mp_int(K) + fixed
                + _mp_str("nodes") + _mp_map(3 * (K + 1)))

    def tail_len(K):
        n = 4 * K
        return len(_mp_str("BMk")) + (2 if n < 256 else 3 if n < 1 << 16 else 5) + n

    rec_sizes = [len(head(p, int(sizes[p]))) + (int(sizes[p]) + 1) * node_bytes + tail_len(int(sizes[p]))
                 for p in range(first, first + count)]

    def abc(sym, typeid):
        return (_mp_map(4) + _mp_str("symbols") + _mp_str(sym) + _mp_str("idx") + bytes([0xC7, 94, 0]) + bytes([0x7F] * 94)
                + _mp_str("any_symbol_id") + _mp_int(55) + _mp_str("typeid") + _mp_int(typeid))

    hdr = (_mp_map(8) + _mp_str("magic_number") + _mp_int(MAGIC_NUMBER) + _mp_str("version") + _mp_int(1)
           + _mp_str("entry_dist") + _mp_int(2) + _mp_str("epsilon") + bytes([0xCA]) + struct.pack(">f", epsilon)
           + _mp_str("abc") + abc("ACGT", 4) + _mp_str("amino") + abc("ACDEFGHIKLMNPQRSTVWY", 2)
           + _mp_str("has_ga") + bytes([0xC3])
           + _mp_str("protein_sizes") + _mp_arr(count) + b"".join(_mp_int(v) for v in rec_sizes))
    total = 0
    with open(path, "wb", buffering=0) as fh:
        top = _mp_map(2) + _mp_str("header") + hdr + _mp_str("proteins") + _mp_arr(count)
        fh.write(top)
        total += len(top)
        for i, p in enumerate(range(first, first + count)):
            K = int(sizes[p])
            ids, bmk = nodes_of(p)
            h, t = head(p, K), _mp_str("BMk") + _mp_f32bin(bmk)
            assert len(h) + (K + 1) * node_bytes + len(t) == rec_sizes[i]
            fh.write(h)
            fh.write(np.take(raw, np.append(ids, ids[-1]), axis=0).data)
            fh.write(t)
            total += rec_sizes[i]
    return {"bytes": total, "profiles": count, "nodes": int(np.sum(sizes[first:first + count]))}
