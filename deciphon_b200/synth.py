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
    return Profile(accession=name, gencode=1, consensus="x" * K, core_size=K,
                   null_emission=pool.null_emission, bg_emission=pool.bg_emission,
                   trans=np.concatenate([tr, tr[-1:]]), emission=np.concatenate([em, em[-1:]]), BMk=bmk)


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
