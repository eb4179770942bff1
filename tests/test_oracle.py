"""CPU tests: the oracle (oracle/dcp_oracle.c) against the reference's golden artefacts,
the committed reference-generated vectors and, when built, the reference's own code."""
import os

import numpy as np
import pytest

from oracle.oracle import encode

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ST_B, ST_E = (3 << 14) | 5, (3 << 14) | 6


def _rows():
    lines = open(os.path.join(GOLDEN, "snap_products.tsv")).read().splitlines()
    hdr = lines[0].split("\t")
    return [dict(zip(hdr, l.split("\t"))) for l in lines[1:]]


def test_golden_snap_lrt_and_paths(oracle, golden_profiles, golden_reads):
    """LRT 291.6 / 349.3 / 360.4 and the full match strings of the reference's snap.dcs."""
    rows = _rows()
    assert [r["lrt"] for r in rows] == ["291.6", "349.3", "360.4"]
    for row, prof, read in zip(rows, golden_profiles, golden_reads["consensus_fna"]):
        assert row["profile"] == prof.accession
        x = encode(read["data"])
        costs = prof.costs()
        xt = oracle.xtrans(len(x), True, False)
        nul, alt = oracle.null(costs[0], xt, x), oracle.alt(costs, xt, x)
        assert "%.1f" % oracle.lrt(nul, alt) == row["lrt"]
        ids, sz = oracle.path(costs, xt, x)
        hs, he, b, e = oracle.hit_extent(ids, sz)
        assert (hs, he) == (int(row["hit_start"]), int(row["hit_stop"]))
        assert (0, len(x)) == (int(row["window_start"]), int(row["window_stop"]))
        # state names and fragments of every step from the first B through the last E
        want = [(m.split(",")[0], m.split(",")[1]) for m in row["match"].split(";")]
        got, pos = [], hs
        for i in range(b, e):
            got.append((read["data"][pos:pos + int(sz[i])], oracle.state_name(ids[i])))
            pos += int(sz[i])
        assert got == want


def test_ref_vectors_scores_bitexact(oracle, golden_profiles, ref_vectors):
    """Oracle == reference viterbi_null/viterbi_cost, bit for bit, 600 pairs x flags."""
    off, sym = ref_vectors["offsets"], ref_vectors["symbols"]
    costs = [p.costs() for p in golden_profiles]
    for f in range(4):
        mh, h3 = bool(f & 1), bool(f & 2)
        for ri in range(len(off) - 1):
            x = np.ascontiguousarray(sym[off[ri]:off[ri + 1]])
            xt = oracle.xtrans(len(x), mh, h3)
            for pi in range(3):
                assert oracle.null(costs[pi][0], xt, x).tobytes() == ref_vectors["null_cost"][f, pi, ri].tobytes()
                assert oracle.alt(costs[pi], xt, x).tobytes() == ref_vectors["alt_cost"][f, pi, ri].tobytes()


def test_ref_vectors_paths(oracle, golden_profiles, ref_vectors):
    """Oracle trace + unzip == reference viterbi_path + trellis_unzip on every hit."""
    off, sym = ref_vectors["offsets"], ref_vectors["symbols"]
    costs = [p.costs() for p in golden_profiles]
    at = 0
    for f, pi, ri, n in ref_vectors["path_key"]:
        x = np.ascontiguousarray(sym[off[ri]:off[ri + 1]])
        xt = oracle.xtrans(len(x), bool(f & 1), bool(f & 2))
        ids, sz = oracle.path(costs[pi], xt, x)
        assert np.array_equal(ids, ref_vectors["path_ids"][at:at + n])
        assert np.array_equal(sz, ref_vectors["path_sizes"][at:at + n])
        assert int(sz.sum()) == len(x)
        at += n
    assert at == len(ref_vectors["path_ids"])


def test_core_costs_transform(oracle, golden_profiles):
    """orc_core_costs (protein.c:353-383 restated in C) == the numpy transform."""
    for p in golden_profiles:
        a = oracle.core_costs(p.core_size, p.BMk, p.trans)
        b = p.costs()[3]
        assert a.tobytes() == b.tobytes()
        assert np.isinf(a[1:, 0][[0, 2, 3, 5, 6]]).all()  # MM,MD,IM,DM,DD into node 0
        assert np.isinf(a[[2, 5], -1]).all()              # MI,II of node K-1


def test_emission_mass_kat(golden_profiles):
    """Known answer for the tables themselves: per-length probability mass
    {e^2(1-e)^2, 2e(1-e)^3, rest, 2e(1-e)^3, e^2(1-e)^2} with epsilon = 0.01."""
    e = 0.01
    want = [e * e * (1 - e) ** 2, 2 * e * (1 - e) ** 3, None, 2 * e * (1 - e) ** 3, e * e * (1 - e) ** 2]
    bounds = [0, 4, 20, 84, 340, 1364]
    p = golden_profiles[0]
    for table in (p.null_emission, p.bg_emission, p.emission[0], p.emission[p.core_size // 2]):
        mass = [np.exp(table[bounds[i]:bounds[i + 1]].astype(np.float64)).sum() for i in range(5)]
        for m, w in zip(mass, want):
            if w is not None:
                assert abs(m - w) / w < 1e-3
        assert abs(sum(mass) - 1) < 1e-3


def test_windows(oracle):
    """window.c:13-37: first window, overlap 4K, last-hit skip, stickiness."""
    assert list(oracle.windows(519, 173)) == [(0, 0, 519)]
    g = oracle.windows(1000, 3)  # 150-nt windows, overlap <= 12 - 1
    w0 = next(g)
    assert w0 == (0, 0, 150)
    w1 = g.send(None)
    assert w1 == (1, 150 + 1 - 12, 150 + 1 - 12 + 150)
    w2 = g.send(140)  # a hit ending at window-relative 141 -> next start skips past it
    assert w2[1] == max(w1[1] + 1, w1[1] + 140 + 1, w1[2] + 1 - 12)
    ws = [w0, w1, w2] + list(g)
    assert ws[-1][2] == 1000 and all(b[1] > a[1] for a, b in zip(ws, ws[1:]))


def test_oracle_vs_reference_random(oracle, reference, golden_profiles, node_pool):
    """Oracle vs the reference's compiled code on synthetic profiles of odd sizes."""
    from deciphon_b200 import synth
    rng = np.random.default_rng(7)
    for K in (2, 3, 7, 8, 9, 16, 17, 31, 33, 100):
        prof = synth.synth_profile(rng, K, node_pool)
        costs = prof.costs()
        rp = reference.profile(costs)
        cons = synth.consensus_dna(node_pool, rng.integers(0, len(node_pool), size=K))
        for trial in range(6):
            if trial < 3:
                x = synth.mutate(rng, np.concatenate([synth.random_read(rng, 10), cons, synth.random_read(rng, 7)]), 0.1)
            else:
                x = synth.random_read(rng, int(rng.integers(1, 200)))
            f = int(rng.integers(0, 4))
            xt = oracle.xtrans(len(x), bool(f & 1), bool(f & 2))
            rp.set_xtrans(xt)
            assert oracle.null(costs[0], xt, x).tobytes() == rp.null(x).tobytes()
            alt = oracle.alt(costs, xt, x)
            assert alt.tobytes() == rp.cost(x).tobytes()
            talt, xn, nd = oracle.trace(costs, xt, x)
            assert talt.tobytes() == alt.tobytes()
            rids, rsz, rxn, rnd = rp.path(x, want_trellis=True)
            ids, sz = oracle.unzip(K, len(x), xn, nd)
            assert np.array_equal(ids, rids) and np.array_equal(sz, rsz)
