"""Oracle of the press path (.hmm -> per-profile tables) -- TEST INFRASTRUCTURE ONLY.

A float64 numpy restatement of what c-core does between reading a HMMER3 ASCII profile and
packing a protein record, so that the product's press (deciphon_b200/host/press.cpp and the
frame-table kernel in csrc/) can be checked.  Restated from:

  hmm_reader.c:19-103   node 0 = begin transitions, then (match emissions, transitions) per node;
                        file values are -ln p, "*" = probability zero; null = Swiss-Prot 50.8
                        background (hmm_reader.c:78-103)
  model.c:62-96         match log-odds = match - null amino log-probs
  model.c:390-441       amino -> codon log-probs (split evenly over the synonymous codons of the
                        genetic code, stop codons impossible, normalised), base log-probs
                        (mean over the three codon positions), codon marginals with wildcards
  model.c:284-309       occupancy-based entry distribution B -> M_k
  protein.c:67-120      node i of the record = match state of node min(i, K-1), transitions out of it
  third-party imm_frame_state (imm_score_table_scores, protein.c:102): the 1..5-nt emission
                        log-probabilities of a frame state; NOT in the reference tree.

Pinning.  The frame-state table is third-party math, restated here as the closed form below and
pinned on the reference's own golden file: all 576 nodes of control/tests/files/minifam.dcp
(stored base log-probs + codon marginals in, stored emission[1364] out) plus the null and
background states reproduce to float32 rounding (tests/test_press_oracle.py).  The rest of this
file is pinned end to end by pressing c-core/minifam.hmm (committed copy of its three profiles'
numbers is NOT needed: the test reads the committed golden npz and a committed copy of the .hmm).

Frame state with base distribution b, codon distribution p and indel rate e (z = emitted
fragment, "_" = any base; single(x) = p(x,_,_)+p(_,x,_)+p(_,_,x), pair(x,y) = p(_,x,y)+p(x,_,y)+p(x,y,_)):
  |z|=1  e^2(1-e)^2/3 * single(z1)
  |z|=2  2e(1-e)^3/3 * pair(z1,z2) + e^3(1-e)/3 * [b(z1) single(z2) + b(z2) single(z1)]
  |z|=3  (1-e)^4 p(z1,z2,z3) + 4e^2(1-e)^2/9 * sum_i b(zi) pair(rest) + e^4/9 * sum_i b(zj) b(zk) single(zi)
  |z|=4  e(1-e)^3/2 * sum_i b(zi) p(rest) + e^3(1-e)/9 * sum_{i<j} b(zi) b(zj) pair(rest)
  |z|=5  e^2(1-e)^2/10 * sum_{i<j} b(zi) b(zj) p(rest)
"""
from __future__ import annotations

import itertools

import numpy as np

AMINO = "ACDEFGHIKLMNPQRSTVWY"  # imm_amino_iupac order = HMMER3 column order
NUCLT = "ACGT"
OFF = {1: 0, 2: 4, 3: 20, 4: 84, 5: 340}
# HMMER3 amino background, Swiss-Prot 50.8 (hmm_reader.c:78-103)
NULL_AMINO = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092,
                       0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364,
                       0.0540687, 0.0673417, 0.0114135, 0.0304133])

# NCBI translation tables in TCAG x TCAG x TCAG order (imm_gencode): id -> 64 amino letters
GENCODES = {
    1: "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    2: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG",
    3: "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    4: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    5: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG",
    6: "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    9: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
    10: "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    11: "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    12: "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    13: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG",
    14: "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
    15: "FFLLSSSSYY*QCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    16: "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    21: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG",
    22: "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    23: "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    24: "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG",
    25: "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    26: "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    27: "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    28: "FFLLSSSSYYQQCCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    29: "FFLLSSSSYYYYCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    30: "FFLLSSSSYYEECC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    31: "FFLLSSSSYYEECCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG",
    33: "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG",
}
_TCAG = {"T": 0, "C": 1, "A": 2, "G": 3}


def codon_amino(gencode: int, a: int, b: int, c: int) -> str:
    """Amino letter of codon (a, b, c) given in ACGT indices."""
    t = [_TCAG[NUCLT[i]] for i in (a, b, c)]
    return GENCODES[gencode][t[0] * 16 + t[1] * 4 + t[2]]


def logsumexp(x):
    x = np.asarray(x, dtype=np.float64)
    m = np.max(x)
    if not np.isfinite(m):
        return m
    return m + np.log(np.sum(np.exp(x - m)))


def nuclt_dist(gencode: int, amino_lprobs):
    """setup_nuclt_dist (model.c:428-441): amino log-probs (or log-odds) -> (base lprobs[4],
    codon marginals[5][5][5] with index 4 = any)."""
    amino_lprobs = np.asarray(amino_lprobs, dtype=np.float64)
    count = {aa: 0 for aa in AMINO}
    for a, b, c in itertools.product(range(4), repeat=3):
        aa = codon_amino(gencode, a, b, c)
        if aa in count:
            count[aa] += 1
    codon = np.full((4, 4, 4), -np.inf)
    for a, b, c in itertools.product(range(4), repeat=3):
        aa = codon_amino(gencode, a, b, c)
        if aa in count:  # stop codons stay impossible (model.c:401-421)
            codon[a, b, c] = amino_lprobs[AMINO.index(aa)] - np.log(count[aa])
    codon -= logsumexp(codon.reshape(-1))  # imm_codon_lprob_normalize
    base = np.full(4, -np.inf)
    for x in range(4):  # model.c:366-388: each codon position contributes lprob - log 3
        terms = []
        for a, b, c in itertools.product(range(4), repeat=3):
            for pos in (a, b, c):
                if pos == x and np.isfinite(codon[a, b, c]):
                    terms.append(codon[a, b, c] - np.log(3))
        base[x] = logsumexp(terms) if terms else -np.inf
    marg = np.full((5, 5, 5), -np.inf)
    P = np.exp(codon)
    for a, b, c in itertools.product(range(5), repeat=3):
        sl = tuple(slice(None) if i == 4 else i for i in (a, b, c))
        s = P[sl].sum()
        marg[a, b, c] = np.log(s) if s > 0 else -np.inf
    return base, marg


def frame_table(eps: float, base_lprobs, marg_lprobs) -> np.ndarray:
    """Emission log-probabilities of a frame state for every 1..5-mer: float64 [1364] in the
    scan's code order (off[len] + big-endian base-4 index)."""
    e = float(eps)
    b = np.exp(np.asarray(base_lprobs, dtype=np.float64))
    P = np.exp(np.asarray(marg_lprobs, dtype=np.float64).reshape(5, 5, 5))
    A = 4

    def single(x):
        return P[x, A, A] + P[A, x, A] + P[A, A, x]

    def pair(x, y):
        return P[A, x, y] + P[x, A, y] + P[x, y, A]

    out = np.zeros(1364)
    for n in range(1, 6):
        for z in itertools.product(range(4), repeat=n):
            if n == 1:
                v = e * e * (1 - e) ** 2 / 3 * single(z[0])
            elif n == 2:
                v = 2 * e * (1 - e) ** 3 / 3 * pair(z[0], z[1]) + e ** 3 * (1 - e) / 3 * (
                    b[z[0]] * single(z[1]) + b[z[1]] * single(z[0]))
            elif n == 3:
                v = (1 - e) ** 4 * P[z[0], z[1], z[2]]
                v += 4 * e * e * (1 - e) ** 2 / 9 * (b[z[0]] * pair(z[1], z[2]) + b[z[1]] * pair(z[0], z[2]) + b[z[2]] * pair(z[0], z[1]))
                v += e ** 4 / 9 * (b[z[1]] * b[z[2]] * single(z[0]) + b[z[0]] * b[z[2]] * single(z[1]) + b[z[0]] * b[z[1]] * single(z[2]))
            elif n == 4:
                v = 0.0
                for i in range(4):
                    r = [z[k] for k in range(4) if k != i]
                    v += e * (1 - e) ** 3 / 2 * b[z[i]] * P[r[0], r[1], r[2]]
                for i, j in itertools.combinations(range(4), 2):
                    r = [z[k] for k in range(4) if k not in (i, j)]
                    v += e ** 3 * (1 - e) / 9 * b[z[i]] * b[z[j]] * pair(r[0], r[1])
            else:
                v = 0.0
                for i, j in itertools.combinations(range(5), 2):
                    r = [z[k] for k in range(5) if k not in (i, j)]
                    v += e * e * (1 - e) ** 2 / 10 * b[z[i]] * b[z[j]] * P[r[0], r[1], r[2]]
            code = 0
            for x in z:
                code = code * 4 + x
            out[OFF[n] + code] = v
    with np.errstate(divide="ignore"):
        return np.log(out)


def frame_joint(eps: float, base_lprobs, marg_lprobs, codon, z) -> float:
    """log p(codon, fragment z) of the frame state: the table's formula with every marginal
    p(pattern) replaced by p(codon) * [codon matches pattern] (third-party imm_frame_cond_lprob,
    called through imm_frame_cond_decode at decoder.c:38-58).  Summed over the 64 codons it gives
    frame_table -- the property tests/test_press_oracle.py checks."""
    e = float(eps)
    b = np.exp(np.asarray(base_lprobs, dtype=np.float64))
    marg = np.asarray(marg_lprobs, dtype=np.float64).reshape(5, 5, 5)
    pc = np.exp(marg[codon[0], codon[1], codon[2]])
    A = 4

    def M(a, bb, c):
        ok = all(p == A or p == q for p, q in zip((a, bb, c), codon))
        return pc if ok else 0.0

    def single(x):
        return M(x, A, A) + M(A, x, A) + M(A, A, x)

    def pair(x, y):
        return M(A, x, y) + M(x, A, y) + M(x, y, A)

    n = len(z)
    if n == 1:
        v = e * e * (1 - e) ** 2 / 3 * single(z[0])
    elif n == 2:
        v = 2 * e * (1 - e) ** 3 / 3 * pair(z[0], z[1]) + e ** 3 * (1 - e) / 3 * (b[z[0]] * single(z[1]) + b[z[1]] * single(z[0]))
    elif n == 3:
        v = (1 - e) ** 4 * M(z[0], z[1], z[2])
        v += 4 * e * e * (1 - e) ** 2 / 9 * (b[z[0]] * pair(z[1], z[2]) + b[z[1]] * pair(z[0], z[2]) + b[z[2]] * pair(z[0], z[1]))
        v += e ** 4 / 9 * (b[z[1]] * b[z[2]] * single(z[0]) + b[z[0]] * b[z[2]] * single(z[1]) + b[z[0]] * b[z[1]] * single(z[2]))
    elif n == 4:
        v = 0.0
        for i in range(4):
            r = [z[k] for k in range(4) if k != i]
            v += e * (1 - e) ** 3 / 2 * b[z[i]] * M(r[0], r[1], r[2])
        for i, j in itertools.combinations(range(4), 2):
            r = [z[k] for k in range(4) if k not in (i, j)]
            v += e ** 3 * (1 - e) / 9 * b[z[i]] * b[z[j]] * pair(r[0], r[1])
    elif n == 5:
        v = 0.0
        for i, j in itertools.combinations(range(5), 2):
            r = [z[k] for k in range(5) if k not in (i, j)]
            v += e * e * (1 - e) ** 2 / 10 * b[z[i]] * b[z[j]] * M(r[0], r[1], r[2])
    else:
        raise ValueError("fragment length must be 1..5")
    return float(np.log(v)) if v > 0 else -np.inf


def frame_decode(eps: float, base_lprobs, marg_lprobs, z):
    """Most likely codon of a fragment: argmax over the 64 codons in ACGT-major order, the first
    maximum wins (imm_frame_cond_decode; the iteration order is read from its callers' results on
    the golden rows only, where no tie occurs).  Returns ((a, b, c), log-prob)."""
    best, arg = -np.inf, None
    for codon in itertools.product(range(4), repeat=3):
        v = frame_joint(eps, base_lprobs, marg_lprobs, codon, z)
        if arg is None or v > best:
            best, arg = v, codon
    return arg, best


# ---- HMMER3 ASCII reader (what hmmer_reader hands to hmm_reader.c) ----------------------------

def _num(tok: str) -> float:
    return -np.inf if tok == "*" else -float(tok)


def read_hmm(path: str):
    """Yields dict(acc, name, leng, match[K][20] lprobs, trans[K+1][7] lprobs (MM,MI,MD,IM,II,DM,DD;
    row 0 = begin node), consensus, has_ga) per profile."""
    with open(path) as fh:
        lines = fh.read().split("\n")
    i = 0
    while i < len(lines):
        if not lines[i].startswith("HMMER3/"):
            i += 1
            continue
        meta = {"acc": "", "name": "", "ga": ""}
        while not lines[i].startswith("HMM "):
            parts = lines[i].split(None, 1)
            if parts and parts[0] in ("NAME", "ACC", "LENG", "GA"):
                meta[parts[0].lower()] = parts[1].strip() if len(parts) > 1 else ""
            i += 1
        i += 2  # "HMM ..." and the transition header line
        if lines[i].split()[0] == "COMPO":
            i += 1
        i += 1  # node 0 insert emissions
        trans = [[_num(t) for t in lines[i].split()[:7]]]
        i += 1
        match, cons = [], []
        while lines[i].strip() != "//":
            toks = lines[i].split()
            match.append([_num(t) for t in toks[1:21]])
            cons.append(toks[22] if len(toks) > 22 else "-")
            trans.append([_num(t) for t in lines[i + 2].split()[:7]])
            i += 3
        i += 1
        yield {"acc": meta["acc"], "name": meta["name"], "leng": int(meta.get("leng", len(match))),
               "match": np.asarray(match, dtype=np.float64), "trans": np.asarray(trans, dtype=np.float64),
               "consensus": "".join(cons), "has_ga": meta["ga"] != ""}


def occupancy(trans) -> np.ndarray:
    """calculate_occupancy (model.c:284-309): log occupancy of every match state, normalised so
    that sum_k occ_k * (K - k) = 1.  `trans` row i = transitions out of node i (row 0 = begin)."""
    K = len(trans) - 1
    MM, MI, DM = trans[:, 0], trans[:, 1], trans[:, 5]
    locc = np.zeros(K)
    locc[0] = np.logaddexp(MI[0], MM[0])
    for i in range(1, K):
        v0 = locc[i - 1] + np.logaddexp(MM[i], MI[i])
        v1 = np.log1p(-np.exp(locc[i - 1])) + DM[i]
        locc[i] = np.logaddexp(v0, v1)
    logZ = logsumexp([locc[i] + np.log(K - i) for i in range(K)])
    return locc - logZ


def press_profile(h, gencode: int = 1, eps: float = 0.01, entry_dist: int = 2):
    """One protein record (protein_absorb, protein.c:67-120) from a read_hmm() profile."""
    K = len(h["match"])
    null_lp = np.log(NULL_AMINO.astype(np.float32).astype(np.float64))  # logf of float literals
    null_nd = nuclt_dist(gencode, null_lp)
    bg_nd = nuclt_dist(gencode, np.zeros(20))
    nodes_nd = [nuclt_dist(gencode, h["match"][k] - null_lp) for k in range(K)]
    nodes_nd.append(nodes_nd[-1])  # record node K repeats node K-1 (protein.c:99)
    emission = np.stack([frame_table(eps, *nd) for nd in nodes_nd])
    trans = np.concatenate([h["trans"][1:], h["trans"][K:]])  # record node i holds alt.trans[min(i+1, K)]
    if entry_dist == 2:
        bmk = occupancy(h["trans"])
    else:
        bmk = np.full(K, np.log(2.0 / (K * (K + 1))) * K)
    return {"accession": h["acc"], "consensus": h["consensus"], "core_size": K, "gencode": gencode,
            "null_nuclt": null_nd, "null_emission": frame_table(eps, *null_nd),
            "bg_nuclt": bg_nd, "bg_emission": frame_table(eps, *bg_nd),
            "node_nuclt": (np.stack([n[0] for n in nodes_nd]), np.stack([n[1].reshape(-1) for n in nodes_nd])),
            "trans": trans, "emission": emission, "BMk": bmk, "has_ga": h["has_ga"]}
