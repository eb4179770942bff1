"""Mirror of python-core's Scan / Batch / Sequence / PressContext (python-core/deciphon_core/
scan.py:23-83, batch.py:8-33, sequence.py, press.py:9-56) over libdeciphon_b200.so -- same
constructor arguments, same methods, same error type behaviour (a DeciphonError carrying the C
error code) -- plus the snap archive step of deciphon_schema.NewSnapFile.make_archive
(schema/deciphon_schema/__init__.py:221-226): products.tsv + hmmer/ zipped into a .dcs file."""
from __future__ import annotations

import ctypes as C
import os
import shutil
from dataclasses import dataclass

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdeciphon_b200.so")

_CALLBACK = C.CFUNCTYPE(None, C.c_void_p)

# every symbol include/deciphon_b200.h declares
SYMBOLS = [
    ("dcp_scan_new", C.c_void_p, []),
    ("dcp_scan_del", None, [C.c_void_p]),
    ("dcp_scan_setup", C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_bool, C.c_bool, C.c_bool, _CALLBACK,
                                 C.c_void_p]),
    ("dcp_scan_run", C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p]),
    ("dcp_scan_interrupt", None, [C.c_void_p]),
    ("dcp_scan_progress", C.c_int, [C.c_void_p]),
    ("dcp_press_new", C.c_void_p, []),
    ("dcp_press_setup", C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    ("dcp_press_open", C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    ("dcp_press_nproteins", C.c_long, [C.c_void_p]),
    ("dcp_press_next", C.c_int, [C.c_void_p]),
    ("dcp_press_end", C.c_bool, [C.c_void_p]),
    ("dcp_press_close", C.c_int, [C.c_void_p]),
    ("dcp_press_del", None, [C.c_void_p]),
    ("dcp_batch_new", C.c_void_p, []),
    ("dcp_batch_del", None, [C.c_void_p]),
    ("dcp_batch_add", C.c_int, [C.c_void_p, C.c_long, C.c_char_p, C.c_char_p]),
    ("dcp_batch_reset", None, [C.c_void_p]),
    ("dcp_error_string", C.c_char_p, [C.c_int]),
    ("dcpb200_db_info", C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_float)]),
    ("dcpb200_scan_num_gpus", C.c_int, [C.c_void_p]),
    ("dcpb200_scan_num_shards", C.c_int, [C.c_void_p]),
    ("dcpb200_scan_counter", C.c_double, [C.c_void_p, C.c_int]),
]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C deciphon_b200/host`")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class DeciphonError(RuntimeError):
    """python-core/deciphon_core/error.py:6-9: message via dcp_error_string."""

    def __init__(self, errno: int):
        self.errno = errno
        super().__init__(lib.dcp_error_string(errno).decode())


@dataclass
class Sequence:
    id: int
    name: str
    data: str


class Batch:
    def __init__(self):
        self._cbatch = lib.dcp_batch_new()
        if not self._cbatch:
            raise MemoryError()

    def add(self, sequence: Sequence):
        if rc := lib.dcp_batch_add(self._cbatch, sequence.id, sequence.name.encode(), sequence.data.encode()):
            raise DeciphonError(rc)

    def reset(self):
        lib.dcp_batch_reset(self._cbatch)

    @property
    def cdata(self):
        return self._cbatch

    def __del__(self):
        if getattr(self, "_cbatch", None):
            lib.dcp_batch_del(self._cbatch)
            self._cbatch = None


class Scan:
    def __init__(self, dbfile, port: int, num_threads: int, multi_hits: bool, hmmer3_compat: bool, cache: bool,
                 on_callback=None):
        """on_callback(scan): optional hook run inside the C callback (python-core's is a no-op that
        only lets Python signal handlers run, scan.py:13-21); tests use it to interrupt a run."""
        self._cscan = lib.dcp_scan_new()
        if not self._cscan:
            raise MemoryError()
        self.interrupted = False
        self.callbacks = 0

        def _cb(_userdata):
            self.callbacks += 1
            if on_callback is not None:
                on_callback(self)

        self._cb = _CALLBACK(_cb)  # keep alive
        path = str(getattr(dbfile, "path", dbfile)).encode()
        if rc := lib.dcp_scan_setup(self._cscan, path, port, num_threads, multi_hits, hmmer3_compat, cache, self._cb,
                                    None):
            lib.dcp_scan_del(self._cscan)
            self._cscan = None
            raise DeciphonError(rc)

    def run(self, snap, batch: Batch):
        """snap: a directory path or an object with .basedir (deciphon_schema.NewSnapFile)."""
        self.interrupted = False
        basedir = str(getattr(snap, "basedir", snap)).encode()
        if rc := lib.dcp_scan_run(self._cscan, batch.cdata, basedir):
            raise DeciphonError(rc)

    def interrupt(self):
        self.interrupted = True
        lib.dcp_scan_interrupt(self._cscan)

    def progress(self) -> int:
        return lib.dcp_scan_progress(self._cscan)

    @property
    def num_gpus(self) -> int:
        """GPUs this scan runs on."""
        return lib.dcpb200_scan_num_gpus(self._cscan)

    @property
    def num_shards(self) -> int:
        """Profile shards (one host thread each; DCP_SHARDS_PER_GPU per GPU for large databases)."""
        return lib.dcpb200_scan_num_shards(self._cscan)

    def counters(self) -> dict:
        """Cumulative H2D / D2H bytes, kernel launches and DP cells of this scan's GPU contexts, windows
        scored and windows that passed the lrt gate."""
        names = ("h2d_bytes", "d2h_bytes", "launches", "cells", "windows", "lrt_windows", "speculative_windows")
        return {n: lib.dcpb200_scan_counter(self._cscan, i) for i, n in enumerate(names)}

    def free(self):
        if getattr(self, "_cscan", None):
            lib.dcp_scan_del(self._cscan)
            self._cscan = None

    def __del__(self):
        self.free()

    def __enter__(self):
        return self

    def __exit__(self, *_):
        self.free()


class PressContext:
    """python-core/deciphon_core/press.py:9-56: hmm -> .dcp, one profile per next()."""

    def __init__(self, hmm, gencode: int, epsilon: float = 0.01, dbpath=None):
        self._cpress = lib.dcp_press_new()
        if not self._cpress:
            raise MemoryError()
        self._hmm = str(getattr(hmm, "path", hmm))
        if dbpath is None:
            dbpath = getattr(getattr(hmm, "dbpath", None), "path", None) or os.path.splitext(self._hmm)[0] + ".dcp"
        self._db = str(dbpath)
        if rc := lib.dcp_press_setup(self._cpress, int(gencode), float(epsilon)):
            raise DeciphonError(rc)

    def open(self):
        if rc := lib.dcp_press_open(self._cpress, self._hmm.encode(), self._db.encode()):
            raise DeciphonError(rc)

    def close(self):
        if rc := lib.dcp_press_close(self._cpress):
            raise DeciphonError(rc)

    def end(self) -> bool:
        return bool(lib.dcp_press_end(self._cpress))

    def next(self):
        if rc := lib.dcp_press_next(self._cpress):
            raise DeciphonError(rc)

    def __enter__(self):
        self.open()
        return self

    def __exit__(self, *_):
        self.close()

    @property
    def nproteins(self) -> int:
        return lib.dcp_press_nproteins(self._cpress)

    def __del__(self):
        if getattr(self, "_cpress", None):
            lib.dcp_press_del(self._cpress)
            self._cpress = None


def press(hmm: str, dbpath: str, gencode: int = 1, epsilon: float = 0.01) -> int:
    """Press every profile of a HMMER3 file (what `deciphon press` drives, cli press loop)."""
    with PressContext(hmm, gencode, epsilon, dbpath) as ctx:
        n = 0
        while True:
            ctx.next()
            if ctx.end():
                break
            n += 1
    return n


def make_snap_archive(basedir: str, dcs_path: str) -> str:
    """NewSnapFile.make_archive: zip `<basedir>` (products.tsv + hmmer/) into `<name>.dcs`, the
    layout deciphon_snap.SnapFile reads (snap/deciphon_snap/snap_file.py:18-33: one root directory
    holding products.tsv and hmmer/).  Without the HMMER stage (port <= 0) hmmer/ is empty."""
    basedir = os.path.abspath(str(basedir))
    if not str(dcs_path).endswith(".dcs"):
        raise ValueError("must end in `.dcs`")
    x = shutil.make_archive(basedir, "zip", os.path.dirname(basedir), os.path.basename(basedir))
    shutil.move(x, str(dcs_path))
    shutil.rmtree(basedir)
    return str(dcs_path)


def db_info(path: str):
    n, total, eps = C.c_int(), C.c_long(), C.c_float()
    if rc := lib.dcpb200_db_info(str(path).encode(), C.byref(n), C.byref(total), C.byref(eps)):
        raise DeciphonError(rc)
    return n.value, total.value, eps.value
