"""python-core binds the reference through a cffi API-mode extension, ``deciphon_core._cffi``
(python-core/build_ext.py:88-112), and its classes only ever touch ``ffi`` and ``lib`` from it
(python-core/deciphon_core/scan.py:5, batch.py:3, press.py:3, error.py:1).  These tests drive
libdeciphon_b200.so through the same extension, built by deciphon_b200/compat/build_cffi.py, the
way python-core's Scan / Batch / DeciphonError do."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def cffi_mod():
    from deciphon_b200.compat import build_cffi
    site = os.path.join(ROOT, "deciphon_b200", "compat", "_site")
    pkg = os.path.join(site, "deciphon_core")
    built = os.path.isdir(pkg) and any(f.startswith("_cffi.") and f.endswith(".so") for f in os.listdir(pkg))
    if not built:  # normally __graft_entry__.build() has done it
        build_cffi.build(site)
    sys.path.insert(0, site)
    try:
        import deciphon_core._cffi as m
    finally:
        sys.path.remove(site)
    return m


def test_cffi_module_exports_the_interface(cffi_mod):
    ffi, lib = cffi_mod.ffi, cffi_mod.lib
    for name in ("dcp_press_new dcp_press_setup dcp_press_open dcp_press_nproteins dcp_press_next dcp_press_end "
                 "dcp_press_close dcp_press_del dcp_scan_new dcp_scan_del dcp_scan_setup dcp_scan_run dcp_scan_interrupt "
                 "dcp_scan_progress dcp_batch_new dcp_batch_del dcp_batch_add dcp_batch_reset dcp_error_string callback").split():
        assert hasattr(lib, name), name
    assert ffi.string(lib.dcp_error_string(21)).decode() == "could not open database file"
    # error path without a GPU: the file is looked at before any device (scan.c:102-108 order)
    scan = lib.dcp_scan_new()
    assert scan != ffi.NULL

    @ffi.def_extern()
    def callback(userdata):
        pass

    rc = lib.dcp_scan_setup(scan, b"/nonexistent/x.dcp", 0, 1, True, False, False, lib.callback, ffi.NULL)
    assert rc == 21  # DCP_EOPENDB
    lib.dcp_scan_del(scan)
    b = lib.dcp_batch_new()
    assert lib.dcp_batch_add(b, 1, b"s", b"ACGTNNRY") == 0 and lib.dcp_batch_add(b, 2, b"t", b"ACGTU") == 74  # T and U mixed
    lib.dcp_batch_del(b)


@pytest.mark.gpu
def test_python_core_style_scan_through_cffi(cffi_mod, tmp_path, golden_profiles, golden_reads):
    """python-core/tests/test_scan.py's flow on the golden database: handle + extern "Python" callback
    (scan.py:12-20), Batch.add, Scan.run, progress -- rows identical to the reference's snap."""
    from deciphon_b200.dcp_file import write_dcp
    ffi, lib = cffi_mod.ffi, cffi_mod.lib
    db = str(tmp_path / "minifam.dcp")
    write_dcp(db, golden_profiles)
    calls = []

    class Owner:
        pass

    owner = Owner()
    handle = ffi.new_handle(owner)

    @ffi.def_extern()
    def callback(userdata):
        assert ffi.from_handle(userdata) is owner
        calls.append(1)

    scan = lib.dcp_scan_new()
    assert lib.dcp_scan_setup(scan, db.encode(), 0, 1, True, False, False, lib.callback, handle) == 0
    batch = lib.dcp_batch_new()
    for r in golden_reads["consensus_fna"]:
        assert lib.dcp_batch_add(batch, r["id"], r["name"].encode(), r["data"].encode()) == 0
    assert lib.dcp_scan_run(scan, batch, str(tmp_path / "snap").encode()) == 0
    assert lib.dcp_scan_progress(scan) == 100 and calls
    lib.dcp_batch_del(batch)
    lib.dcp_scan_del(scan)
    rows = (tmp_path / "snap" / "products.tsv").read_text().splitlines()
    want = open(os.path.join(GOLDEN, "snap_products.tsv")).read().splitlines()
    key = lambda l: tuple(l.split("\t")[i] for i in (0, 7))  # noqa: E731
    gotmap = {key(l): l for l in rows[1:]}
    assert rows[0] == want[0]
    for l in want[1:]:
        assert gotmap[key(l)] == l
