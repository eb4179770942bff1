"""The reference-facing host library (libdeciphon_b200.so, include/deciphon_b200.h): same API
as c-core/deciphon.h.  CPU tests cover symbols, .dcp parsing (both float encodings), batch
clean-up and error behaviour; the GPU test reproduces the reference's golden snap.dcs rows."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DCP = "/root/reference/control/tests/files/minifam.dcp"


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:dcpb200|dcp)_[a-z0-9_]+)\s*\(", src)))


def test_host_library_exports_reference_api():
    from deciphon_b200 import scan
    names = _declared("deciphon_b200.h")
    # every function of c-core/deciphon.h:9-32 is there
    for n in ("dcp_scan_new dcp_scan_del dcp_scan_setup dcp_scan_run dcp_scan_interrupt dcp_scan_progress "
              "dcp_press_new dcp_press_setup dcp_press_open dcp_press_nproteins dcp_press_next dcp_press_end "
              "dcp_press_close dcp_press_del dcp_batch_new dcp_batch_del dcp_batch_add dcp_batch_reset "
              "dcp_error_string").split():
        assert n in names
    raw = ctypes.CDLL(scan.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), n
    assert sorted(n for n, _, _ in scan.SYMBOLS) == names


def test_error_strings_and_press_errors(tmp_path):
    import torch
    from deciphon_b200 import scan
    assert scan.lib.dcp_error_string(20) == b"out of memory"
    assert b"no CUDA device" in scan.lib.dcp_error_string(81)
    assert scan.lib.dcp_error_string(999).startswith(b"unknown error #999")
    p = scan.lib.dcp_press_new()
    assert scan.lib.dcp_press_open(p, b"x.hmm", b"x.dcp") == 49  # DCP_ESETGENCODE: setup comes first
    assert scan.lib.dcp_press_setup(p, 7, 0.01) == 50            # DCP_EGENCODEID (press.c:55-56): no NCBI table 7
    assert scan.lib.dcp_press_setup(p, 1, 0.01) == 0
    assert scan.lib.dcp_press_open(p, str(tmp_path / "missing.hmm").encode(), str(tmp_path / "x.dcp").encode()) == 22
    if not torch.cuda.is_available():  # no CPU fallback: the frame tables are computed on the GPU
        hmm = os.path.join(GOLDEN, "minifam.hmm")
        assert scan.lib.dcp_press_open(p, hmm.encode(), str(tmp_path / "x.dcp").encode()) == 81
        assert not (tmp_path / "x.dcp.records.tmp").exists()
    scan.lib.dcp_press_del(p)


def test_batch_cleanup_errors():
    from deciphon_b200.scan import Batch, DeciphonError, Sequence
    b = Batch()
    b.add(Sequence(1, "ok", "acgtnnryACGT"))
    with pytest.raises(DeciphonError) as e:
        b.add(Sequence(2, "bad", "ACGTU"))  # both T and U: DCP_ENUCLTSEQTU (disambiguate.c:54)
    assert e.value.errno == 74
    b.reset()


def test_dcp_roundtrip_both_encodings(tmp_path, golden_profiles):
    """write_dcp -> python reader and -> the C++ reader of the host library, bin and ext."""
    from deciphon_b200 import scan
    from deciphon_b200.dcp_file import read_dcp, write_dcp
    for enc in ("bin", "ext"):
        path = str(tmp_path / f"mini_{enc}.dcp")
        write_dcp(path, golden_profiles, encoding=enc)
        db = read_dcp(path)
        assert [p.accession for p in db.proteins] == [p.accession for p in golden_profiles]
        for a, b in zip(db.proteins, golden_profiles):
            assert a.emission.tobytes() == b.emission.tobytes() and a.trans.tobytes() == b.trans.tobytes()
            assert a.BMk.tobytes() == b.BMk.tobytes() and a.null_emission.tobytes() == b.null_emission.tobytes()
        n, total, eps = scan.db_info(path)
        assert (n, total) == (3, 173 + 241 + 162) and abs(eps - 0.01) < 1e-9


@pytest.mark.skipif(not os.path.exists(REF_DCP), reason="reference tree absent (GPU box)")
def test_cpp_reader_parses_reference_golden_file(golden_profiles):
    from deciphon_b200 import scan
    from deciphon_b200.dcp_file import read_dcp
    assert scan.db_info(REF_DCP)[:2] == (3, 576)
    db = read_dcp(REF_DCP)  # and the committed fixture is exactly that file's content
    for a, b in zip(db.proteins, golden_profiles):
        assert a.emission.tobytes() == b.emission.tobytes() and a.BMk.tobytes() == b.BMk.tobytes()


def test_setup_errors_without_gpu(tmp_path, golden_profiles):
    from deciphon_b200.scan import DeciphonError, Scan
    with pytest.raises(DeciphonError) as e:
        Scan(str(tmp_path / "missing.dcp"), 0, 1, True, False, False)
    assert e.value.errno == 21  # DCP_EOPENDB
    bad = tmp_path / "bad.dcp"
    bad.write_bytes(b"\x82\xa6header\x88garbage")
    with pytest.raises(DeciphonError) as e:
        Scan(str(bad), 0, 1, True, False, False)
    assert e.value.errno == 69  # DCP_ENOTDBFILE
    with pytest.raises(DeciphonError) as e:
        Scan(str(bad), 51371, 1, True, False, False)
    assert e.value.errno == 51  # DCP_EH3CDIAL: no HMMER client in this build
    with pytest.raises(DeciphonError) as e:
        Scan(str(bad), 0, 129, True, False, False)
    assert e.value.errno == 42  # DCP_EMANYTHREADS
    import torch
    if not torch.cuda.is_available():
        from deciphon_b200.dcp_file import write_dcp
        path = str(tmp_path / "mini.dcp")
        write_dcp(path, golden_profiles[:1])
        with pytest.raises(DeciphonError) as e:
            Scan(path, 0, 1, True, False, False)
        assert e.value.errno == 81  # DCP_EGPUNODEVICE: no CPU fallback


@pytest.mark.gpu
def test_scan_reproduces_golden_snap(tmp_path, golden_profiles, golden_reads):
    """Press-less version of python-core/tests/test_scan.py and c-core/test_scan.c: scan the
    consensus reads against minifam and compare products.tsv with the reference's golden
    control/tests/files/snap.dcs row by row (every column; evalue is 0 in both)."""
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    path = str(tmp_path / "minifam.dcp")
    write_dcp(path, golden_profiles)
    batch = Batch()
    for r in golden_reads["consensus_fna"]:
        batch.add(Sequence(r["id"], r["name"], r["data"]))
    out = tmp_path / "snap"
    with Scan(path, 0, 2, True, False, False) as scan:
        scan.run(str(out), batch)
        assert scan.progress() == 100
        got = (out / "products.tsv").read_text().splitlines()
        want = open(os.path.join(GOLDEN, "snap_products.tsv")).read().splitlines()
        assert got[0] == want[0]
        # the golden file holds the three true hits; a scan also reports any other lrt >= 0 window
        # (the reference would pass those to HMMER).  Compare the rows the golden file has.
        key = lambda l: tuple(l.split("\t")[i] for i in (0, 7))
        gotmap = {key(l): l for l in got[1:]}
        for l in want[1:]:
            assert gotmap[key(l)] == l
        # the scan object is reusable (test_scan.c:57-69)
        scan.run(str(tmp_path / "snap2"), batch)
        assert (tmp_path / "snap2" / "products.tsv").read_text().splitlines() == got


@pytest.mark.gpu
@pytest.mark.parametrize("K,copies", [(20, 8), (3, 12)])  # K = 3: the shape of c-core/massive.hmm (config 2)
def test_scan_windows_long_read(tmp_path, node_pool, oracle, K, copies):
    """A read longer than min(50K, 100000) is cut into overlapping window.c windows whose starts
    depend on the previous window's hit; rows must match an oracle-driven scan of the same read."""
    from deciphon_b200 import synth
    from deciphon_b200.dcp_file import write_dcp
    from deciphon_b200.scan import Batch, Scan, Sequence
    rng = np.random.default_rng(77)
    prof = synth.synth_profile(rng, K, node_pool, name="SYN%d" % K)
    cons = np.argmax(prof.emission[:K, 20:84], axis=1)
    cons = np.stack([cons // 16, (cons // 4) % 4, cons % 4], axis=1).reshape(-1).astype(np.uint8)
    parts = []
    for i in range(copies):
        parts += [synth.random_read(rng, int(rng.integers(150, 500))), synth.mutate(rng, cons, 0.03)]
    x = np.concatenate(parts + [synth.random_read(rng, 333)])
    data = "".join("ACGT"[v] for v in x)
    path = str(tmp_path / "syn.dcp")
    write_dcp(path, [prof])
    batch = Batch()
    batch.add(Sequence(7, "long", data))
    with Scan(path, 0, 1, True, False, False) as scan:
        scan.run(str(tmp_path / "o"), batch)
    got = [l.split("\t") for l in (tmp_path / "o" / "products.tsv").read_text().splitlines()[1:]]
    # oracle-driven reference loop (thread.c:49-208 minus HMMER)
    costs = prof.costs()
    want = []
    g = oracle.windows(len(x), K)
    try:
        w = next(g)
        while True:
            idx, a, b = w
            win = np.ascontiguousarray(x[a:b])
            xt = oracle.xtrans(len(win), True, False)
            lrt = oracle.lrt(oracle.null(costs[0], xt, win), oracle.alt(costs, xt, win))
            last = None
            if np.isfinite(lrt) and lrt >= 0:
                ids, sz = oracle.path(costs, xt, win)
                ext = oracle.hit_extent(ids, sz)
                if ext:
                    want.append((idx, a, b, ext[0], ext[1], "%.1f" % lrt, [oracle.state_name(s) for s in ids[ext[2]:ext[3]]]))
                    last = ext[1] - 1
            w = g.send(last)
    except StopIteration:
        pass
    assert (len(want) >= 3 or K < 10) and len(got) == len(want)
    for row, (idx, a, b, hs, he, lrt, names) in zip(got, want):
        assert (int(row[1]), int(row[2]), int(row[3]), int(row[5]), int(row[6]), row[9]) == (idx, a, b, hs, he, lrt)
        assert [m.split(",")[1] for m in row[11].split(";")] == names


@pytest.mark.gpu
def test_config5_long_read_full_traceback(device, oracle, node_pool):
    """BASELINE.json config 5: one 24,000-nt read (random + embedded genes + 10 % errors) against
    profiles of 50..2000 nodes, window.c windows driven by the decoded hits, full trellis traceback.
    Every window's null/alt cost is bit-exact with the oracle and every hit path identical."""
    from deciphon_b200 import synth
    from deciphon_b200.device import PAIR_DTYPE
    rng = np.random.default_rng(2024)
    sizes = [50, 200, 500, 1000, 2000]
    profs = [synth.synth_profile(rng, K, node_pool, name=f"C5_{K}") for K in sizes]
    parts = []
    for p in profs:
        cons = np.argmax(p.emission[:p.core_size, 20:84], axis=1)
        cons = np.stack([cons // 16, (cons // 4) % 4, cons % 4], axis=1).reshape(-1).astype(np.uint8)
        parts += [synth.random_read(rng, 600), cons[: min(len(cons), 3000)]]
    x = synth.mutate(rng, np.concatenate(parts + [synth.random_read(rng, 25000)]), 0.10)[:24000]
    assert len(x) == 24000
    base = device.num_profiles
    for p in profs:
        device.add_profile(p)
    device.set_reads([x])
    nwin = nhit = 0
    for pi, p in enumerate(profs):
        costs = p.costs()
        g = oracle.windows(len(x), p.core_size)
        last = None
        try:
            w = next(g)
            while True:
                idx, a, b = w
                win = np.ascontiguousarray(x[a:b])
                pair = np.asarray([(base + pi, 0, a, b - a)], dtype=np.int32).view(PAIR_DTYPE).reshape(-1)
                nul, alt = device.score_pairs(pair, True, False)
                xt = oracle.xtrans(len(win), True, False)
                assert nul[0].tobytes() == oracle.null(costs[0], xt, win).tobytes(), (p.core_size, idx)
                assert alt[0].tobytes() == oracle.alt(costs, xt, win).tobytes(), (p.core_size, idx)
                nwin += 1
                last = None
                lrt = oracle.lrt(nul[0], alt[0])
                if np.isfinite(lrt) and lrt >= 0:
                    talt, paths = device.trace_pairs(pair, True, False)
                    ids, sz = oracle.path(costs, xt, win)
                    assert np.array_equal(paths[0][0], ids) and np.array_equal(paths[0][1], sz), (p.core_size, idx)
                    ext = oracle.hit_extent(ids, sz)
                    if ext:
                        last = ext[1] - 1
                        nhit += 1
                w = g.send(last)
        except StopIteration:
            pass
    assert nwin >= 5 + 9 and nhit >= 4  # K = 50 alone yields ~10 windows of 2,500 nt


@pytest.mark.gpu
def test_config2_and_config5_literal_against_the_compiled_reference(reference, node_pool):
    """BASELINE.json configs 2 and 5 as bench.py times them (K = 3 x 1,000 reads of 1 kb in 150-nt
    windows; one 24-kb read x profiles of 50..2000 nodes with full traceback) through dcp_scan_run,
    against the reference's own scan loop (oracle/_ref): the same DP cells -- every window start
    depends on the hits decoded before it -- and the same number of windows through the lrt gate."""
    import argparse
    import bench
    args = argparse.Namespace(seed=20261018, tmp="")
    gpu = bench.small_config_legs(args, node_pool)
    cpu = bench.cpu_small_configs(args, node_pool, gpu)
    for name in ("config2", "config5"):
        assert gpu[name]["windows"] > 0 and cpu[name]["parity_with_gpu_leg"] is True, (name, gpu[name], cpu[name])
    assert gpu["config2"]["windows"] >= 7000 and gpu["config5"]["lrt_windows"] >= 5 and gpu["config5"]["rows"] >= 5
