"""CPU tests of the press oracle (oracle/press.py): the frame-state emission formula, the amino ->
codon model and the HMMER3 reader against the reference's own golden database
(control/tests/files/minifam.dcp, committed as tests/golden/minifam.npz) pressed from
c-core/minifam.hmm (committed copy: tests/golden/minifam.hmm)."""
import itertools
import os

import numpy as np
import pytest

from oracle import press

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5  # natural-log units; the golden values are float32 (about 3e-6 observed)


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(GOLDEN, "minifam.npz"), allow_pickle=True)


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a))
    return float(np.max(np.abs(a[fin] - b[fin]))) <= tol


def test_frame_table_reproduces_every_golden_state(z):
    """All 579 node records + null + background of the three golden profiles: stored base
    log-probs and codon marginals in, stored emission[1364] out."""
    eps = float(z["epsilon"])
    seen = set()
    for pi in range(3):
        n4, cm, em = z[f"p{pi}_node_nuclt4"], z[f"p{pi}_node_nuclt125"], z[f"p{pi}_emission"]
        for node in range(len(n4)):
            key = n4[node].tobytes() + cm[node].tobytes()
            if key in seen:
                continue
            seen.add(key)
            assert _close(press.frame_table(eps, n4[node], cm[node]), em[node]), (pi, node)
        for nm in ("null", "bg"):
            assert _close(press.frame_table(eps, z[f"p{pi}_{nm}_nuclt4"], z[f"p{pi}_{nm}_nuclt125"]), z[f"p{pi}_{nm}_emission"])
    assert len(seen) >= 570


def test_press_of_minifam_hmm_matches_golden_dcp(z):
    """HMMER3 text -> nuclt_dist, transitions, occupancy entry distribution, consensus."""
    hs = list(press.read_hmm(os.path.join(GOLDEN, "minifam.hmm")))
    assert [h["acc"] for h in hs] == [str(z[f"p{i}_accession"]) for i in range(3)]
    null_lp = np.log(press.NULL_AMINO.astype(np.float32).astype(np.float64))
    for pi, h in enumerate(hs):
        K = len(h["match"])
        assert K == int(z[f"p{pi}_core_size"]) and h["consensus"] == str(z[f"p{pi}_consensus"]) and h["has_ga"]
        for k in (0, 1, K // 2, K - 1):
            b, m = press.nuclt_dist(1, h["match"][k] - null_lp)
            assert _close(b, z[f"p{pi}_node_nuclt4"][k]) and _close(m.reshape(-1), z[f"p{pi}_node_nuclt125"][k])
        assert _close(np.concatenate([h["trans"][1:], h["trans"][K:]]), z[f"p{pi}_trans"])
        assert _close(press.occupancy(h["trans"]), z[f"p{pi}_BMk"])
        b, m = press.nuclt_dist(1, null_lp)
        assert _close(b, z[f"p{pi}_null_nuclt4"]) and _close(m.reshape(-1), z[f"p{pi}_null_nuclt125"])
        b, m = press.nuclt_dist(1, np.zeros(20))
        assert _close(b, z[f"p{pi}_bg_nuclt4"]) and _close(m.reshape(-1), z[f"p{pi}_bg_nuclt125"])


def test_joint_marginalises_to_the_table(z):
    """sum over the 64 codons of p(codon, fragment) = p(fragment): ties the decoder's formula to the
    emission table the golden file pins -- for every fragment length, including stop codons."""
    eps = float(z["epsilon"])
    n4, cm = z["p1_node_nuclt4"][7], z["p1_node_nuclt125"][7]
    tab = press.frame_table(eps, n4, cm)
    rng = np.random.default_rng(5)
    for n in range(1, 6):
        for _ in range(6):
            frag = tuple(int(v) for v in rng.integers(0, 4, n))
            tot = sum(np.exp(press.frame_joint(eps, n4, cm, c, frag)) for c in itertools.product(range(4), repeat=3))
            code = 0
            for v in frag:
                code = code * 4 + v
            # the stored marginals are float32: "any" entries differ from the sums of their parts by ~1e-8
            assert abs(np.log(tot) - tab[press.OFF[n] + code]) < 1e-6


def test_decode_properties(z):
    eps = float(z["epsilon"])
    n4, cm = z["p0_node_nuclt4"][3], z["p0_node_nuclt125"][3]
    marg = cm.reshape(5, 5, 5)
    # an admissible, likely codon decodes to itself
    best = np.unravel_index(np.argmax(marg[:4, :4, :4]), (4, 4, 4))
    assert press.frame_decode(eps, n4, cm, tuple(int(v) for v in best))[0] == tuple(int(v) for v in best)
    # a stop codon (probability zero under table 1) never decodes to itself
    taa = (3, 0, 0)
    assert not np.isfinite(marg[taa])
    got, lp = press.frame_decode(eps, n4, cm, taa)
    assert got != taa and np.isfinite(lp) and np.isfinite(marg[got])
    # 1-, 2-, 4- and 5-nt fragments decode to a codon that contains the fragment's bases in order
    for frag in [(2,), (0, 3), (1, 1, 2, 0), (3, 2, 1, 0, 2)]:
        got, lp = press.frame_decode(eps, n4, cm, frag)
        assert np.isfinite(lp) and np.isfinite(marg[got])
        if len(frag) < 3:  # deletions only: the fragment is a subsequence of the codon
            it = iter(got)
            assert all(any(x == y for y in it) for x in frag)
        if len(frag) == 4:  # one insertion is the leading term: the codon is a subsequence of the fragment
            it = iter(frag)
            assert all(any(x == y for y in it) for x in got)


def test_gencode_tables():
    assert press.codon_amino(1, 0, 3, 2) == "M" and press.codon_amino(1, 3, 0, 0) == "*"  # ATG, TAA
    assert press.codon_amino(4, 3, 2, 0) == "W" and press.codon_amino(1, 3, 2, 0) == "*"  # TGA
    for gid, t in press.GENCODES.items():
        assert len(t) == 64 and set(t) <= set(press.AMINO + "*"), gid
