"""CPU tests of the host-side pieces bench.py's legs stand on: the streaming synthetic .dcp writer
and the vectorised window.c rule of deciphon_b200/waves.py (against the oracle's window iterator)."""
import numpy as np


def _import_waves():
    # waves imports the ctypes binding of libdcpgpu.so (loadable without a GPU; no compute call here)
    from deciphon_b200 import waves
    return waves


def test_streaming_writer_equals_write_dcp(tmp_path, node_pool):
    from deciphon_b200 import synth
    from deciphon_b200.dcp_file import read_dcp, write_dcp
    sizes = np.asarray([20, 33, 257, 64, 1], dtype=np.int64)
    nodes = [synth.synth_profile_nodes(np.random.default_rng([7, 1, p]), int(sizes[p]), node_pool) for p in range(len(sizes))]
    info = synth.write_synth_dcp(str(tmp_path / "a.dcp"), sizes, node_pool, lambda p: nodes[p])
    profs = [synth.synth_profile(np.random.default_rng([7, 1, p]), int(sizes[p]), node_pool, "SYN%05d" % p)
             for p in range(len(sizes))]
    write_dcp(str(tmp_path / "b.dcp"), profs)
    a, b = (tmp_path / "a.dcp").read_bytes(), (tmp_path / "b.dcp").read_bytes()
    assert a == b and info["bytes"] == len(a) and info["nodes"] == int(sizes.sum())
    db = read_dcp(str(tmp_path / "a.dcp"))
    assert [p.core_size for p in db.proteins] == sizes.tolist()
    assert np.array_equal(db.proteins[2].emission[:257], node_pool.emission[nodes[2][0]])
    # a sub-range of a larger database
    synth.write_synth_dcp(str(tmp_path / "c.dcp"), sizes, node_pool, lambda p: nodes[p], first=1, count=2)
    assert [p.accession for p in read_dcp(str(tmp_path / "c.dcp")).proteins] == ["SYN00001", "SYN00002"]


def test_window_next_matches_the_oracle(oracle):
    waves = _import_waves()
    rng = np.random.default_rng(3)
    for _ in range(200):
        K = int(rng.integers(1, 60))
        n = int(rng.integers(1, 6000))
        g = oracle.windows(n, K)
        start, stop, last = np.asarray([0]), np.asarray([min(50 * K, 100000, n)]), np.asarray([-1])
        w = next(g)
        assert (w[1], w[2]) == (0, int(stop[0]))
        while True:
            lhp = None
            if rng.random() < 0.3:  # a hit somewhere in the window
                lhp = int(rng.integers(0, stop[0] - start[0]))
                last = np.asarray([lhp])
            alive, ns, ne = waves.window_next(start, stop, last, np.asarray([n]), np.asarray([K]))
            try:
                w = g.send(lhp)
            except StopIteration:
                assert not alive[0]
                break
            assert alive[0] and (w[1], w[2]) == (int(ns[0]), int(ne[0]))
            start, stop = ns, ne
