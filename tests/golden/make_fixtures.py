"""Regenerates the committed fixtures under tests/golden/ (run in the build container,
where /root/reference exists; the GPU box only ever reads the outputs).

  minifam.npz         parameters of the reference's golden control/tests/files/minifam.dcp
                      (3 profiles, 576 nodes), re-encoded as plain float32 arrays
  reads.json          the reference's test reads: control/tests/files/consensus.fna and
                      c-core/test_consensus.h
  snap_products.tsv   products.tsv of the reference's golden control/tests/files/snap.dcs
  ref_vectors.npz     outputs of the REFERENCE's own viterbi.c/trellis.c (oracle/_ref) on
                      seeded inputs: null/alt costs for every pair, decoded paths for hits

Usage:  python tests/golden/make_fixtures.py
"""
import json
import os
import re
import sys
import zipfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

from deciphon_b200 import synth  # noqa: E402
from deciphon_b200.dcp_file import read_dcp  # noqa: E402
from oracle.oracle import Oracle, Reference, encode  # noqa: E402


def write_minifam():
    db = read_dcp(f"{REF}/control/tests/files/minifam.dcp")
    d = {"num_profiles": np.int64(len(db.proteins)), "epsilon": np.float32(db.epsilon),
         "entry_dist": np.int64(db.entry_dist)}
    for i, p in enumerate(db.proteins):
        d[f"p{i}_accession"] = np.str_(p.accession)
        d[f"p{i}_consensus"] = np.str_(p.consensus)
        d[f"p{i}_gencode"] = np.int64(p.gencode)
        d[f"p{i}_core_size"] = np.int64(p.core_size)
        d[f"p{i}_null_emission"] = p.null_emission
        d[f"p{i}_bg_emission"] = p.bg_emission
        d[f"p{i}_trans"] = p.trans
        d[f"p{i}_emission"] = p.emission
        d[f"p{i}_BMk"] = p.BMk
        d[f"p{i}_null_nuclt4"], d[f"p{i}_null_nuclt125"] = p.null_nuclt
        d[f"p{i}_bg_nuclt4"], d[f"p{i}_bg_nuclt125"] = p.bg_nuclt
        d[f"p{i}_node_nuclt4"], d[f"p{i}_node_nuclt125"] = p.node_nuclt
    np.savez_compressed(os.path.join(OUT, "minifam.npz"), **d)
    return db


def write_reads():
    fna = []
    for line in open(f"{REF}/control/tests/files/consensus.fna"):
        line = line.strip()
        if line.startswith(">"):
            fna.append({"id": len(fna), "name": line[1:], "data": ""})
        elif line:
            fna[-1]["data"] += line
    src = open(f"{REF}/c-core/test_consensus.h").read()
    cons = []
    for m in re.finditer(r"\{(\d+),\s*\"([^\"]+)\",\s*((?:\"[ACGT]*\"\s*)+)\}", src):
        data = "".join(re.findall(r"\"([ACGT]*)\"", m.group(3)))
        cons.append({"id": int(m.group(1)), "name": m.group(2), "data": data})
    assert len(fna) == 3 and len(cons) == 8, (len(fna), len(cons))
    json.dump({"consensus_fna": fna, "test_consensus_h": cons}, open(os.path.join(OUT, "reads.json"), "w"), indent=1)
    return fna, cons


def write_snap():
    z = zipfile.ZipFile(f"{REF}/control/tests/files/snap.dcs")
    open(os.path.join(OUT, "snap_products.tsv"), "wb").write(z.read("snap/products.tsv"))


def write_ref_vectors(db, cons):
    """Seeded cases through the reference's compiled viterbi.c/trellis.c."""
    orc, ref = Oracle(), Reference()
    rng = np.random.default_rng(20261018)
    costs = [p.costs() for p in db.proteins]
    rprofs = [ref.profile(c) for c in costs]
    base = [encode(c["data"]) for c in cons]
    reads = list(base)                                    # 8 clean consensus reads
    for rate in (0.02, 0.10, 0.25):                       # mutated copies
        for b in base:
            reads.append(synth.mutate(rng, b, rate))
    for L in (1, 2, 3, 4, 5, 6, 7, 11, 16, 17, 31, 33, 64, 150, 400):   # short/random reads
        reads.append(synth.random_read(rng, L))
    for b in base[:3]:                                    # fragments
        a = int(rng.integers(0, len(b) // 2))
        reads.append(np.ascontiguousarray(b[a:a + int(rng.integers(30, len(b) // 2))]))
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in reads])
    out = {"symbols": np.concatenate(reads), "offsets": off}
    nul = np.zeros((4, 3, len(reads)), dtype=np.float32)
    alt = np.zeros_like(nul)
    path_ids, path_sz, path_key = [], [], []
    for f in range(4):
        mh, h3 = bool(f & 1), bool(f & 2)
        for pi, rp in enumerate(rprofs):
            for ri, x in enumerate(reads):
                rp.set_xtrans(orc.xtrans(len(x), mh, h3))
                nul[f, pi, ri] = rp.null(x)
                alt[f, pi, ri] = rp.cost(x)
                lrt = -2 * ((-nul[f, pi, ri]) - (-alt[f, pi, ri]))
                if np.isfinite(lrt) and lrt >= 0:
                    ids, sz = rp.path(x)
                    path_key.append((f, pi, ri, len(ids)))
                    path_ids.append(ids)
                    path_sz.append(sz)
    out["null_cost"] = nul
    out["alt_cost"] = alt
    out["path_key"] = np.asarray(path_key, dtype=np.int64)
    out["path_ids"] = np.concatenate(path_ids)
    out["path_sizes"] = np.concatenate(path_sz)
    out["ref_lanes"] = np.int64(ref.lanes)
    np.savez_compressed(os.path.join(OUT, "ref_vectors.npz"), **out)
    print("ref vectors:", nul.size, "pairs,", len(path_key), "hit paths,", int(out["path_ids"].size), "steps")


if __name__ == "__main__":
    db = write_minifam()
    fna, cons = write_reads()
    write_snap()
    write_ref_vectors(db, cons)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
