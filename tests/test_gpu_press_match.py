"""GPU tests of the rows SURVEY 8(f) marks "next" around the hot path: the press (.hmm -> .dcp,
frame tables on the device), the on-device post-processing of traced paths (hit extent, codons,
aminos, match column), the multi-GPU shards and the per-chunk callback / interrupt / progress of
dcp_scan_run -- all through the C ABI / the reference API."""
import os

import numpy as np
import pytest

from oracle import press as opress

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu
TOL = 2e-5


def _close(a, b, tol=TOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a))
    return float(np.max(np.abs(a[fin] - b[fin]))) <= tol


def test_frame_tables_match_golden_nodes(device, golden_profiles):
    """dcpgpu_frame_tables on all 579 golden node records (+ null, background) against the emission
    tables the reference stored for them, and against the float64 oracle."""
    for p in golden_profiles:
        n4, n125 = p.node_nuclt
        got = device.frame_tables(n4, n125, 0.01)
        assert _close(got, p.emission)
        assert _close(device.frame_tables(p.null_nuclt[0], p.null_nuclt[1], 0.01)[0], p.null_emission)
        assert _close(device.frame_tables(p.bg_nuclt[0], p.bg_nuclt[1], 0.01)[0], p.bg_emission)
    p = golden_profiles[1]
    for k in (0, 17, 240):
        assert _close(device.frame_tables(p.node_nuclt[0][k], p.node_nuclt[1][k], 0.03)[0],
                      opress.frame_table(0.03, p.node_nuclt[0][k], p.node_nuclt[1][k]), 5e-6)


def test_press_minifam_size_tables_and_rescan(tmp_path, golden_profiles, golden_reads):
    """c-core/test_press.c + python-core/tests/test_press.py: minifam.hmm pressed with gencode 1,
    epsilon 0.01 is exactly 3,609,858 bytes; every table agrees with the reference's golden
    minifam.dcp; scanning the pressed file reproduces the golden snap rows (test_scan.c)."""
    from deciphon_b200.dcp_file import read_dcp
    from deciphon_b200.scan import Batch, PressContext, Scan, Sequence
    db = tmp_path / "press.dcp"
    with PressContext(os.path.join(GOLDEN, "minifam.hmm"), 1, 0.01, str(db)) as ctx:
        assert ctx.nproteins == 3
        n = 0
        while True:
            ctx.next()
            if ctx.end():
                break
            n += 1
        assert n == 3
    assert os.path.getsize(db) == 3609858  # test_press.c:26
    assert not os.path.exists(str(db) + ".records.tmp")
    got = read_dcp(str(db))
    assert got.has_ga and abs(got.epsilon - 0.01) < 1e-9 and got.entry_dist == 2
    for a, b in zip(got.proteins, golden_profiles):
        assert (a.accession, a.consensus, a.core_size, a.gencode) == (b.accession, b.consensus, b.core_size, b.gencode)
        assert _close(a.emission, b.emission) and _close(a.trans, b.trans) and _close(a.BMk, b.BMk)
        assert _close(a.null_emission, b.null_emission) and _close(a.bg_emission, b.bg_emission)
        assert _close(a.node_nuclt[0], b.node_nuclt[0]) and _close(a.node_nuclt[1], b.node_nuclt[1])
    batch = Batch()
    for r in golden_reads["consensus_fna"]:
        batch.add(Sequence(r["id"], r["name"], r["data"]))
    with Scan(str(db), 0, 1, True, False, False) as scan:
        scan.run(str(tmp_path / "snap"), batch)
    rows = (tmp_path / "snap" / "products.tsv").read_text().splitlines()
    want = open(os.path.join(GOLDEN, "snap_products.tsv")).read().splitlines()
    key = lambda l: tuple(l.split("\t")[i] for i in (0, 7))
    gotmap = {key(l): l for l in rows[1:]}
    for l in want[1:]:
        assert gotmap[key(l)] == l


def _expected_match(prof, x, ids, sizes, ext, eps=0.01):
    """The match column of a row, restated: per step of the B..E segment
    "<fragment>,<state>,<codon>,<amino>" (match.c:66-90, product_thread.c:112-148) with the
    frame-state decoder over the state's nuclt_dist (decoder.c:38-58)."""
    from oracle.oracle import Oracle
    names = Oracle()
    hs, he, b, e = ext
    pos = hs
    out = []
    for j in range(b, e):
        sid, sz = int(ids[j]), int(sizes[j])
        frag = x[pos:pos + sz]
        name = names.state_name(sid)
        msb = sid >> 14
        mute = name in ("S", "B", "E", "T") or msb == 2
        if mute:
            out.append(",".join(["", name, "", ""]))
        else:
            if msb == 1:
                nd = prof.bg_nuclt
            elif msb == 0:
                k = (sid & 0x3FFF) - 1
                nd = (prof.node_nuclt[0][k], prof.node_nuclt[1][k])
            else:
                nd = prof.null_nuclt
            codon, lp = opress.frame_decode(eps, nd[0], nd[1], tuple(int(v) for v in frag))
            assert np.isfinite(lp)
            out.append(",".join(["".join("ACGT"[v] for v in frag), name, "".join("ACGT"[v] for v in codon),
                                 opress.codon_amino(prof.gencode, *codon)]))
        pos += sz
    return ";".join(out)


def test_match_column_with_indels_matches_restated_decoder(tmp_path, golden_profiles, oracle):
    """Reads with insertions and deletions force 1-, 2-, 4- and 5-nt steps and stop-codon triplets:
    every row of products.tsv (extent, lrt, match column byte for byte) against the oracle's path
    and the restated frame-state decoder."""
    from deciphon_b200 import synth
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    rng = np.random.default_rng(11)
    path = str(tmp_path / "mini.dcp")
    write_dcp(path, golden_profiles)
    reads, batch = [], Batch()
    for i, p in enumerate(golden_profiles):
        K = p.core_size
        cons = np.argmax(p.emission[:K, 20:84], axis=1)
        cons = np.stack([cons // 16, (cons // 4) % 4, cons % 4], axis=1).reshape(-1).astype(np.uint8)
        for err in (0.04, 0.10):
            x = np.concatenate([synth.random_read(rng, 40), synth.mutate(rng, cons, err), synth.random_read(rng, 25)])
            reads.append((i, x))
            batch.add(Sequence(len(reads), f"r{len(reads)}", "".join("ACGT"[v] for v in x)))
    with Scan(path, 0, 1, True, False, False) as scan:
        scan.run(str(tmp_path / "o"), batch)
    rows = [l.split("\t") for l in (tmp_path / "o" / "products.tsv").read_text().splitlines()[1:]]
    got = {(int(r[0]), r[7]): r for r in rows}
    sizes_seen = set()
    checked = 0
    for ri, (pi, x) in enumerate(reads):
        p = golden_profiles[pi]
        costs = p.costs()
        xt = oracle.xtrans(len(x), True, False)
        lrt = oracle.lrt(oracle.null(costs[0], xt, x), oracle.alt(costs, xt, x))
        assert np.isfinite(lrt) and lrt >= 0
        ids, sz = oracle.path(costs, xt, x)
        ext = oracle.hit_extent(ids, sz)
        assert ext
        row = got[(ri + 1, p.accession)]
        assert (int(row[5]), int(row[6]), row[9]) == (ext[0], ext[1], "%.1f" % lrt)
        assert row[11] == _expected_match(p, x, ids, sz, ext)
        sizes_seen |= {int(s) for s in sz[ext[2]:ext[3]]}
        checked += 1
    assert checked == 6 and {2, 3, 4} <= sizes_seen  # indel steps were exercised


def test_interrupt_progress_and_chunked_callbacks(tmp_path, golden_profiles, golden_reads, monkeypatch):
    """thread.c:74-79 / scan.c:218-227 at chunk granularity: the callback fires after every chunk
    of profiles, progress advances with it and an interrupt stops the run before the next chunk."""
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    path = str(tmp_path / "mini.dcp")
    write_dcp(path, golden_profiles)
    batch = Batch()
    for r in golden_reads["consensus_fna"]:
        batch.add(Sequence(r["id"], r["name"], r["data"]))
    monkeypatch.setenv("DCP_CHUNK_CELLS", "1")  # one profile per chunk
    seen = []
    with Scan(path, 0, 1, True, False, False, on_callback=lambda s: seen.append(s.progress())) as scan:
        scan.run(str(tmp_path / "full"), batch)
        assert seen == [33, 66, 100] and scan.progress() == 100
    full = (tmp_path / "full" / "products.tsv").read_text().splitlines()

    def stop(s):
        s.interrupt()

    with Scan(path, 0, 1, True, False, False, on_callback=stop) as scan:
        scan.run(str(tmp_path / "part"), batch)  # returns 0 like thread_run after an interrupt
        assert scan.progress() == 33
    part = (tmp_path / "part" / "products.tsv").read_text().splitlines()
    assert part[0] == full[0] and 1 <= len(part) < len(full) and part == full[:len(part)]
    monkeypatch.delenv("DCP_CHUNK_CELLS")
    with Scan(path, 0, 1, True, False, False) as scan:  # chunking does not change the rows
        scan.run(str(tmp_path / "one"), batch)
    assert (tmp_path / "one" / "products.tsv").read_text().splitlines() == full


def test_two_gpus_give_the_same_products(tmp_path, golden_profiles, golden_reads):
    """Profile shards over two GPUs inside dcp_scan_setup/run (scan.c:95-152,188-208): the merged
    products.tsv is byte-identical to the one-GPU file."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    path = str(tmp_path / "mini.dcp")
    write_dcp(path, golden_profiles)
    batch = Batch()
    for r in golden_reads["consensus_fna"]:
        batch.add(Sequence(r["id"], r["name"], r["data"]))
    outs = []
    for threads in (1, 2):
        with Scan(path, 0, threads, True, False, False) as scan:
            assert scan.num_gpus == threads
            scan.run(str(tmp_path / f"o{threads}"), batch)
            assert scan.progress() == 100
        outs.append((tmp_path / f"o{threads}" / "products.tsv").read_bytes())
    assert outs[0] == outs[1] and outs[0].count(b"\n") >= 4


def test_snap_archive_layout(tmp_path, golden_profiles, golden_reads):
    """NewSnapFile.make_archive: one root directory with products.tsv and hmmer/ inside the .dcs zip
    (snap/deciphon_snap/snap_file.py:18-33)."""
    import zipfile
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence, make_snap_archive
    path = str(tmp_path / "mini.dcp")
    write_dcp(path, golden_profiles)
    batch = Batch()
    for r in golden_reads["consensus_fna"][:1]:
        batch.add(Sequence(r["id"], r["name"], r["data"]))
    with Scan(path, 0, 1, True, False, False) as scan:
        scan.run(str(tmp_path / "snap"), batch)
    dcs = make_snap_archive(str(tmp_path / "snap"), str(tmp_path / "snap.dcs"))
    names = zipfile.ZipFile(dcs).namelist()
    assert "snap/products.tsv" in names and "snap/hmmer/" in names and not (tmp_path / "snap").exists()


def test_window_waves_over_the_c_abi_match_dcp_scan_run(tmp_path, node_pool):
    """deciphon_b200.waves (bench.py's measured leg: window.c waves driven over the C ABI) and
    dcp_scan_run (the C++ host loop) score exactly the same windows: same DP cells, same number of
    windows through the lrt gate -- every later window's start depends on the hits decoded before it."""
    from deciphon_b200 import synth, waves
    from deciphon_b200.device import PAIR_DTYPE, Device
    from deciphon_b200.scan import Batch, Scan, Sequence
    rng = np.random.default_rng(5)
    sizes = np.asarray([20, 24, 31, 39, 40, 64, 300], dtype=np.int64)
    nodes = [synth.synth_profile_nodes(np.random.default_rng([5, 1, p]), int(sizes[p]), node_pool) for p in range(len(sizes))]
    reads = []
    for i in range(12):
        x = synth.random_read(rng, 2000)
        if i % 3 == 0:  # embed the consensus of a small profile a few times so that later windows hold hits
            q = int(rng.integers(0, 4))
            cons = synth.consensus_dna(node_pool, nodes[q][0])
            parts = []
            for _ in range(6):
                parts += [synth.random_read(rng, int(rng.integers(100, 300))), cons]
            x = synth.fixed_length(rng, np.concatenate(parts), 2000)
        reads.append(synth.mutate(rng, x, 0.05)[:1900 + 10 * i])
    db = str(tmp_path / "w.dcp")
    synth.write_synth_dcp(db, sizes, node_pool, lambda p: nodes[p])
    batch = Batch()
    for i, r in enumerate(reads):
        batch.add(Sequence(i, f"r{i}", "".join("ACGT"[v] for v in r)))
    with Scan(db, 0, 1, True, False, False) as scan:
        scan.run(str(tmp_path / "o"), batch)
        want = scan.counters()
    with Device(0) as dev:
        first = dev.pool_add(node_pool.emission, node_pool.trans)
        for p in range(len(sizes)):
            dev.profile_add(int(sizes[p]), nodes[p][1], node_pool.null_emission, node_pool.bg_emission, nodes[p][0] + first)
        dev.set_reads(reads)
        lens = np.asarray([len(r) for r in reads], dtype=np.int64)
        R = len(reads)
        dev.score_grid(0, len(sizes), 0, R, True, False)
        idx = dev.hits_fetch()
        pr = np.zeros(len(idx), dtype=PAIR_DTYPE)
        pr["profile"], pr["seq"] = idx // R, idx % R
        pr["len"] = waves.first_windows(sizes, lens)[idx // R, idx % R]
        hit, _hs, he, _n = waves.trace_hits(dev, pr, True, False, Ks=sizes)
        w = waves.later_waves(dev, sizes, 0, lens, pr, hit, he, True, False)
        got = dev.counters()
    assert w["pairs"] > 0 and w["hits"] > 0  # later windows exist and some hold hits
    assert got["cells"] == want["cells"]
    assert len(idx) + w["hits"] == want["lrt_windows"] and len(sizes) * R + w["pairs"] == want["windows"]


def test_batched_amino_fasta_for_the_hmmer_stage(tmp_path, golden_profiles, golden_reads, monkeypatch):
    """DCP_WRITE_AMINOS=1: hmmer/aminos.fa holds, row by row, the amino-acid sequence the reference
    would send to the HMMER daemon for that hit (thread.c:168-190: the amino letters of the non-mute
    steps of the match) -- checked against the golden rows' own match columns."""
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    path = str(tmp_path / "mini.dcp")
    write_dcp(path, golden_profiles)
    batch = Batch()
    for r in golden_reads["consensus_fna"]:
        batch.add(Sequence(r["id"], r["name"], r["data"]))
    monkeypatch.setenv("DCP_WRITE_AMINOS", "1")
    with Scan(path, 0, 1, True, False, False) as scan:
        scan.run(str(tmp_path / "snap"), batch)
    rows = [l.split("\t") for l in (tmp_path / "snap" / "products.tsv").read_text().splitlines()[1:]]
    fa = (tmp_path / "snap" / "hmmer" / "aminos.fa").read_text().splitlines()
    assert len(fa) == 2 * len(rows) and len(rows) >= 3
    golden = {(l.split("\t")[0], l.split("\t")[7]): l.split("\t")[11]
              for l in open(os.path.join(GOLDEN, "snap_products.tsv")).read().splitlines()[1:]}
    seen = 0
    for row, head, seq in zip(rows, fa[0::2], fa[1::2]):
        assert head.startswith(f">{row[0]}/{row[1]}/{row[7]} ") and f"hit=[{row[5]},{row[6]})" in head
        want = "".join(step.split(",")[3] for step in row[11].split(";") if step.split(",")[3])
        assert seq == want and len(seq) > 100
        if (row[0], row[7]) in golden:
            assert seq == "".join(st.split(",")[3] for st in golden[(row[0], row[7])].split(";") if st.split(",")[3])
            seen += 1
    assert seen == 3
