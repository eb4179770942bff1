"""ctypes binding of the C ABI declared in include/dcpgpu.h (deciphon_b200/libdcpgpu.so).

The library is built in-tree by ``deciphon_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
There is no CPU fallback: if the shared object is missing the import of this module fails,
and if no CUDA device is present ``dcpgpu_open`` returns DCPGPU_ENODEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# DCPGPU_LIB: developer switch for A/B builds of the CUDA library (scripts/quick_perf.py); the
# product always loads the in-tree deciphon_b200/libdcpgpu.so
LIB_PATH = os.environ.get("DCPGPU_LIB") or os.path.join(HERE, "libdcpgpu.so")

# every symbol include/dcpgpu.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _u32, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
SYMBOLS = [
    ("dcpgpu_open", C.c_int, [C.POINTER(_vp), C.c_int]),
    ("dcpgpu_close", None, [_vp]),
    ("dcpgpu_strerror", C.c_char_p, [C.c_int]),
    ("dcpgpu_last_error", C.c_char_p, [_vp]),
    ("dcpgpu_set_stream", C.c_int, [_vp, _vp]),
    ("dcpgpu_sync", C.c_int, [_vp]),
    ("dcpgpu_device_info", _i64, [_vp, C.c_int]),
    ("dcpgpu_device_count", _i32, []),
    ("dcpgpu_pool_add", C.c_int, [_vp, C.c_int, _vp, _vp, C.POINTER(_i64)]),
    ("dcpgpu_profile_add", C.c_int, [_vp, C.c_int, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i32)]),
    ("dcpgpu_profile_count", C.c_int, [_vp]),
    ("dcpgpu_profile_core_size", C.c_int, [_vp, _i32]),
    ("dcpgpu_pool_release", C.c_int, [_vp]),
    ("dcpgpu_reads_set", C.c_int, [_vp, _i32, _vp, _vp]),
    ("dcpgpu_reads_count", C.c_int, [_vp]),
    ("dcpgpu_score_pairs", C.c_int, [_vp, _i64, _vp, _u32, _vp, _vp]),
    ("dcpgpu_score_grid", C.c_int, [_vp, _i32, _i32, _i32, _i32, _u32]),
    ("dcpgpu_scores_fetch", C.c_int, [_vp, _i64, _vp, _vp]),
    ("dcpgpu_hits_fetch", C.c_int, [_vp, _i64, _vp, C.POINTER(_i64)]),
    ("dcpgpu_scores_gather", C.c_int, [_vp, _i64, _vp, _vp, _vp]),
    ("dcpgpu_last_cells", C.c_double, [_vp]),
    ("dcpgpu_last_kernel_ms", _f32, [_vp]),
    ("dcpgpu_last_redo", _i64, [_vp]),
    ("dcpgpu_launch_count", _i64, [_vp]),
    ("dcpgpu_counter", C.c_double, [_vp, C.c_int]),
    ("dcpgpu_alu_peak", C.c_int, [_vp, C.c_int, C.POINTER(C.c_double)]),
    ("dcpgpu_frame_tables", C.c_int, [_vp, _i32, _vp, _vp, _f32, _vp]),
    ("dcpgpu_profile_set_decoder", C.c_int, [_vp, _i32, _vp, _vp, _vp, C.c_char_p]),
    ("dcpgpu_match_build", C.c_int, [_vp, _f32, C.c_int, _vp, _vp, _vp, _vp]),
    ("dcpgpu_match_fetch", C.c_int, [_vp, _vp]),
    ("dcpgpu_trace_pairs", C.c_int, [_vp, _i64, _vp, _u32, _vp, _vp]),
    ("dcpgpu_trace_fetch", C.c_int, [_vp, _vp, _vp, _vp]),
    ("dcpgpu_trace_trellis", C.c_int, [_vp, _i64, _vp, _vp]),
    ("dcpgpu_xtrans", C.c_int, [C.c_int, _u32, _vp]),
]

MULTI_HITS = 1
HMMER3_COMPAT = 2
KEEP_TRELLIS = 4


class DcpGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"dcpgpu error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C deciphon_b200/csrc` "
            "(deciphon_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()
