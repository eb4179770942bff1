"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/dcpgpu.h declares; host-only entry points agree with the oracle; without a GPU
the product fails loudly instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:dcpgpu|dcp)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from deciphon_b200 import _lib
    names = _declared("dcpgpu.h")
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/dcpgpu.h but not exported"
    assert sorted(n for n, _, _ in _lib.SYMBOLS) == names


def test_xtrans_matches_oracle(oracle):
    from deciphon_b200 import device
    for mh in (False, True):
        for h3 in (False, True):
            for L in list(range(1, 400)) + [999, 1000, 2000, 24000, 99999, 100000]:
                assert device.xtrans(L, mh, h3).tobytes() == oracle.xtrans(L, mh, h3).tobytes()


def test_xtrans_matches_reference_xtrans_c(oracle):
    """dcpgpu_xtrans and the oracle against the REFERENCE's own xtrans.c (compiled where it lies
    into oracle/_ref/libdcpref_xtrans.so), bit for bit, every window length 1..100000 (the
    maximum window, window.c:7) x the four multi_hits x hmmer3_compat combinations."""
    from deciphon_b200 import device
    from oracle.oracle import ref_xtrans
    lens = np.arange(1, 100001)
    for mh in (False, True):
        for h3 in (False, True):
            want = ref_xtrans(lens, mh, h3)
            # costs depend on L only through max(L / 3, 1): one device/oracle call per distinct value
            first = np.unique(np.maximum(lens // 3, 1), return_index=True)[1]
            for i in first:
                L = int(lens[i])
                assert device.xtrans(L, mh, h3).tobytes() == want[i].tobytes(), (L, mh, h3)
                assert oracle.xtrans(L, mh, h3).tobytes() == want[i].tobytes(), (L, mh, h3)
            # and the expansion back to every length is the reference's
            same = np.maximum(lens // 3, 1)
            assert np.array_equal(want.view(np.uint32), want[first][np.searchsorted(same[first], same)].view(np.uint32))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from deciphon_b200.device import DcpGpuError, Device
    with pytest.raises(DcpGpuError) as e:
        Device(0)
    assert e.value.code == 1  # DCPGPU_ENODEVICE


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under deciphon_b200/ may reference it."""
    pkg = os.path.join(ROOT, "deciphon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower().replace("oracle-generated", ""), f
                assert "dcporacle" not in src and "dcpref" not in src, f


def test_layout_positions_are_a_permutation():
    """Device row layout (layout.cuh): pos() must be a bijection on [0, Kpad)."""
    def shape(K):
        w = 1
        while 32 * w * 8 < K:
            w *= 2
        return (K + 32 * w - 1) // (32 * w), w

    def pos(k, Q, VL):
        vl, q = divmod(k, Q)
        n4 = Q & ~3
        if q < n4:
            q0, w = q & ~3, 4
        elif (Q & 2) and q < n4 + 2:
            q0, w = n4, 2
        else:
            q0, w = n4 + (Q & 2), 1
        return VL * q0 + vl * w + (q - q0)

    for K in (1, 2, 3, 31, 32, 33, 100, 200, 224, 255, 256, 257, 511, 513, 1000, 2000, 5000, 16384):
        Q, W = shape(K)
        assert Q <= 8
        Kpad = 32 * W * Q
        assert Kpad >= K
        assert sorted(pos(k, Q, 32 * W) for k in range(Kpad)) == list(range(Kpad))
