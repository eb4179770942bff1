"""Reader for Deciphon's ``.dcp`` profile database (host side, pure Python + numpy).

The on-disk layout is the MessagePack-family stream written by the reference's press
path (c-core/database_writer.c:95-193, c-core/protein.c:234-281) and read back by
c-core/database_reader.c:26-80 and c-core/protein.c:283-351:

    map(2){ "header": map(8){magic_number, version, entry_dist, epsilon, abc, amino,
                             has_ga, protein_sizes},
            "proteins": array(n){ map(10){accession, gencode, consensus, core_size,
                                          null_nuclt_dist, null_emission, bg_nuclt_dist,
                                          bg_emission, nodes, BMk} } }

Two encodings of the float arrays exist in the wild and both are accepted:
the current writer (c-core/write.c:59-66) emits ``bin`` + host-endian floats, while the
golden ``control/tests/files/minifam.dcp`` carries ``ext`` type 8 big-endian floats (and
``protein_sizes`` as ``ext`` type 6 big-endian u32).

All values are natural-log probabilities; the scan path negates them into costs
(c-core/protein.c:353-394), which is what :func:`Profile.costs` does.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

MAGIC_NUMBER = 0xC6F1  # c-core/magic_number.h
NUM_CODES = 1364  # c-core/protein_node_size.h:4-9 = 4 + 16 + 64 + 256 + 1024
TRANS_NAMES = ("MM", "MI", "MD", "IM", "II", "DM", "DD")  # c-core/trans.h
CORE_NAMES = ("BM", "MM", "MI", "MD", "IM", "II", "DM", "DD")  # order used by the C ABI


class DcpFormatError(ValueError):
    pass


class _Ext:
    __slots__ = ("type", "data")

    def __init__(self, type_: int, data: bytes):
        self.type = type_
        self.data = data


class _Bin:
    __slots__ = ("data",)

    def __init__(self, data: bytes):
        self.data = data


class _Stream:
    """Minimal MessagePack tokenizer (only what lite-pack emits)."""

    def __init__(self, data: bytes):
        self.d = memoryview(data)
        self.p = 0

    def _u(self, n: int) -> int:
        v = int.from_bytes(self.d[self.p : self.p + n], "big")
        self.p += n
        return v

    def _raw(self, n: int) -> bytes:
        v = bytes(self.d[self.p : self.p + n])
        if len(v) != n:
            raise DcpFormatError("truncated stream")
        self.p += n
        return v

    def read(self):
        b = self.d[self.p]
        self.p += 1
        if b <= 0x7F:
            return b
        if 0x80 <= b <= 0x8F:
            return self._map(b & 0xF)
        if 0x90 <= b <= 0x9F:
            return [self.read() for _ in range(b & 0xF)]
        if 0xA0 <= b <= 0xBF:
            return self._raw(b & 0x1F).decode()
        if b == 0xC0:
            return None
        if b == 0xC2:
            return False
        if b == 0xC3:
            return True
        if 0xC4 <= b <= 0xC6:
            return _Bin(self._raw(self._u(1 << (b - 0xC4))))
        if 0xC7 <= b <= 0xC9:
            n = self._u(1 << (b - 0xC7))
            t = self._u(1)
            return _Ext(t, self._raw(n))
        if b == 0xCA:
            return struct.unpack(">f", self._raw(4))[0]
        if b == 0xCB:
            return struct.unpack(">d", self._raw(8))[0]
        if 0xCC <= b <= 0xCF:
            return self._u(1 << (b - 0xCC))
        if 0xD0 <= b <= 0xD3:
            n = 1 << (b - 0xD0)
            return int.from_bytes(self._raw(n), "big", signed=True)
        if 0xD4 <= b <= 0xD8:
            n = 1 << (b - 0xD4)
            t = self._u(1)
            return _Ext(t, self._raw(n))
        if 0xD9 <= b <= 0xDB:
            return self._raw(self._u(1 << (b - 0xD9))).decode()
        if b == 0xDC:
            return [self.read() for _ in range(self._u(2))]
        if b == 0xDD:
            return [self.read() for _ in range(self._u(4))]
        if b == 0xDE:
            return self._map(self._u(2))
        if b == 0xDF:
            return self._map(self._u(4))
        if b >= 0xE0:
            return b - 256
        raise DcpFormatError(f"unsupported token 0x{b:02x} at {self.p - 1}")

    def _map(self, n: int):
        # keys repeat inside "nodes", so keep an ordered list of pairs
        return [(self.read(), self.read()) for _ in range(n)]


def _f32(tok, n: int | None = None) -> np.ndarray:
    if isinstance(tok, _Ext):  # golden/older encoding: big-endian
        a = np.frombuffer(tok.data, dtype=">f4").astype("<f4")
    elif isinstance(tok, _Bin):  # current writer: raw host-endian
        a = np.frombuffer(tok.data, dtype="<f4").copy()
    elif isinstance(tok, list):
        a = np.asarray(tok, dtype="<f4")
    else:
        raise DcpFormatError("expected an f32 array")
    if n is not None and a.size != n:
        raise DcpFormatError(f"expected {n} floats, found {a.size}")
    return a


def _nuclt_dist(tok):
    """nuclt_dist = {nucleotide lprobs f32[4], codon marginal lprobs f32[125]}
    (c-core/nuclt_dist.c:13-20; the golden encoding is array(2){f32[4], f32[125]})."""
    if isinstance(tok, list) and len(tok) == 2 and not isinstance(tok[0], tuple):
        return _f32(tok[0], 4), _f32(tok[1], 125)
    if isinstance(tok, list) and tok and isinstance(tok[0], tuple):  # map form
        vals = [v for _, v in tok]
        arrs = [_f32(v) for v in vals if isinstance(v, (_Ext, _Bin, list))]
        a4 = next(a for a in arrs if a.size == 4)
        a125 = next(a for a in arrs if a.size == 125)
        return a4, a125
    raise DcpFormatError("unrecognised nuclt_dist encoding")


@dataclass
class Profile:
    """One protein profile in ``.dcp`` (log-prob) form."""

    accession: str
    gencode: int
    consensus: str
    core_size: int
    null_emission: np.ndarray  # f32[1364]
    bg_emission: np.ndarray  # f32[1364]
    trans: np.ndarray  # f32[K+1, 7]  transitions OUT of node i (MM,MI,MD,IM,II,DM,DD)
    emission: np.ndarray  # f32[K+1, 1364]  (node K duplicates node K-1, protein.c:99)
    BMk: np.ndarray  # f32[K]
    null_nuclt: tuple = field(default=None, repr=False)
    bg_nuclt: tuple = field(default=None, repr=False)
    node_nuclt: tuple = field(default=None, repr=False)  # (f32[K+1,4], f32[K+1,125])

    def costs(self):
        """Negate into the DP's cost form, the transform of c-core/protein.c:353-394.

        Returns (nul[1364], bg[1364], em[K,1364], core[8,K]) with core rows in
        ``CORE_NAMES`` order, indexed by DESTINATION node: MM,MD,IM,DM,DD of node k
        sit at k+1, MI,II at k; node 0's incoming and node K-1's MI/II are +inf."""
        K = self.core_size
        core = np.full((8, K), np.inf, dtype=np.float32)
        core[0, :] = -self.BMk
        t = self.trans
        if K > 1:
            core[1, 1:] = -t[: K - 1, 0]  # MM
            core[2, : K - 1] = -t[: K - 1, 1]  # MI
            core[3, 1:] = -t[: K - 1, 2]  # MD
            core[4, 1:] = -t[: K - 1, 3]  # IM
            core[5, : K - 1] = -t[: K - 1, 4]  # II
            core[6, 1:] = -t[: K - 1, 5]  # DM
            core[7, 1:] = -t[: K - 1, 6]  # DD
        nul = (-self.null_emission).astype(np.float32)
        bg = (-self.bg_emission).astype(np.float32)
        em = np.ascontiguousarray(-self.emission[:K]).astype(np.float32)
        return nul, bg, em, core


@dataclass
class Database:
    entry_dist: int
    epsilon: float
    abc_symbols: str
    abc_typeid: int
    amino_symbols: str
    has_ga: bool
    protein_sizes: list
    proteins: list


def _profile(tok) -> Profile:
    if not (isinstance(tok, list) and len(tok) == 10):
        raise DcpFormatError("protein record must be a map of 10")
    keys = [k for k, _ in tok]
    want = ["accession", "gencode", "consensus", "core_size", "null_nuclt_dist", "null_emission",
            "bg_nuclt_dist", "bg_emission", "nodes", "BMk"]
    if keys != want:  # expect_key order, c-core/protein.c:283-351
        raise DcpFormatError(f"unexpected protein keys {keys}")
    m = dict(tok)
    K = int(m["core_size"])
    nodes = m["nodes"]
    if len(nodes) != 3 * (K + 1):
        raise DcpFormatError("nodes map must hold 3*(core_size+1) entries")
    trans = np.empty((K + 1, 7), dtype=np.float32)
    emis = np.empty((K + 1, NUM_CODES), dtype=np.float32)
    nd4 = np.empty((K + 1, 4), dtype=np.float32)
    nd125 = np.empty((K + 1, 125), dtype=np.float32)
    for i in range(K + 1):
        (k0, v0), (k1, v1), (k2, v2) = nodes[3 * i : 3 * i + 3]
        if (k0, k1, k2) != ("nuclt_dist", "trans", "emission"):
            raise DcpFormatError("unexpected node keys")
        nd4[i], nd125[i] = _nuclt_dist(v0)
        trans[i] = _f32(v1, 7)
        emis[i] = _f32(v2, NUM_CODES)
    return Profile(
        accession=m["accession"],
        gencode=int(m["gencode"]),
        consensus=m["consensus"],
        core_size=K,
        null_emission=_f32(m["null_emission"], NUM_CODES),
        bg_emission=_f32(m["bg_emission"], NUM_CODES),
        trans=trans,
        emission=emis,
        BMk=_f32(m["BMk"], K),
        null_nuclt=_nuclt_dist(m["null_nuclt_dist"]),
        bg_nuclt=_nuclt_dist(m["bg_nuclt_dist"]),
        node_nuclt=(nd4, nd125),
    )


def read_dcp(path: str) -> Database:
    """Parse a whole ``.dcp`` file (c-core/database_reader.c:26-80 + protein.c:283-351)."""
    with open(path, "rb") as fh:
        data = fh.read()
    s = _Stream(data)
    top = s.read()
    if not (isinstance(top, list) and len(top) == 2 and top[0][0] == "header" and top[1][0] == "proteins"):
        raise DcpFormatError("not a deciphon database (DCP_ENOTDBFILE)")
    hdr = dict(top[0][1])
    if hdr.get("magic_number") != MAGIC_NUMBER:
        raise DcpFormatError("bad magic number (DCP_ENOTDBFILE)")
    if hdr.get("version") != 1:
        raise DcpFormatError("unsupported database version (DCP_EDBVERSION)")
    abc = dict(hdr["abc"])
    amino = dict(hdr["amino"])
    ps = hdr["protein_sizes"]
    if isinstance(ps, _Ext):
        sizes = np.frombuffer(ps.data, dtype=">u4").astype(np.int64).tolist()
    elif isinstance(ps, _Bin):
        sizes = np.frombuffer(ps.data, dtype="<u4").astype(np.int64).tolist()
    else:
        sizes = [int(v) for v in ps]
    prots = [_profile(t) for t in top[1][1]]
    if len(prots) != len(sizes):
        raise DcpFormatError("protein count mismatch (DCP_EINVALNUMPROTEINS)")
    return Database(
        entry_dist=int(hdr["entry_dist"]),
        epsilon=float(hdr["epsilon"]),
        abc_symbols=abc.get("symbols", "ACGT"),
        abc_typeid=int(abc.get("typeid", 4)),
        amino_symbols=amino.get("symbols", ""),
        has_ga=bool(hdr["has_ga"]),
        protein_sizes=sizes,
        proteins=prots,
    )


# ---- writer (current encoding: c-core/write.c:59-66 `bin` + host-endian floats) ---------------

def _mp_str(s: str) -> bytes:
    b = s.encode()
    if len(b) < 32:
        return bytes([0xA0 | len(b)]) + b
    if len(b) < 256:
        return bytes([0xD9, len(b)]) + b
    return bytes([0xDA]) + struct.pack(">H", len(b)) + b


def _mp_int(v: int) -> bytes:
    if 0 <= v < 128:
        return bytes([v])
    if 0 <= v < 1 << 16:
        return bytes([0xCD]) + struct.pack(">H", v)
    return bytes([0xCE]) + struct.pack(">I", v)


def _mp_map(n: int) -> bytes:
    return bytes([0x80 | n]) if n < 16 else (bytes([0xDE]) + struct.pack(">H", n) if n < 1 << 16 else bytes([0xDF]) + struct.pack(">I", n))


def _mp_arr(n: int) -> bytes:
    return bytes([0x90 | n]) if n < 16 else (bytes([0xDC]) + struct.pack(">H", n) if n < 1 << 16 else bytes([0xDD]) + struct.pack(">I", n))


def _mp_f32bin(a) -> bytes:
    raw = np.ascontiguousarray(a, dtype="<f4").tobytes()
    n = len(raw)
    head = bytes([0xC4, n]) if n < 256 else (bytes([0xC5]) + struct.pack(">H", n) if n < 1 << 16 else bytes([0xC6]) + struct.pack(">I", n))
    return head + raw


def _mp_f32ext(a) -> bytes:
    raw = np.ascontiguousarray(a, dtype=">f4").tobytes()
    n = len(raw)
    head = bytes([0xC7, n]) if n < 256 else (bytes([0xC8]) + struct.pack(">H", n) if n < 1 << 16 else bytes([0xC9]) + struct.pack(">I", n))
    return head + bytes([8]) + raw


def write_dcp(path: str, profiles, epsilon: float = 0.01, entry_dist: int = 2, encoding: str = "bin",
              symbols: str = "ACGT"):
    """Write profiles as a .dcp file in the reference's key order (database_writer.c:158-193,
    protein.c:234-281).  `encoding` selects the float-array form: "bin" (current writer) or
    "ext" (the golden file's older form).  The alphabet sub-maps are written like the golden
    file's (their current encoding is defined inside third-party imm)."""
    f32 = _mp_f32bin if encoding == "bin" else _mp_f32ext

    def nuclt(d):
        a4, a125 = d if d is not None else (np.log(np.full(4, 0.25, np.float32)), np.zeros(125, np.float32))
        return _mp_arr(2) + f32(a4) + f32(a125)

    def abc(sym, typeid):
        return (_mp_map(4) + _mp_str("symbols") + _mp_str(sym) + _mp_str("idx") + bytes([0xC7, 94, 0]) + bytes([0x7F] * 94)
                + _mp_str("any_symbol_id") + _mp_int(55) + _mp_str("typeid") + _mp_int(typeid))

    recs = []
    for p in profiles:
        K = p.core_size
        nd4, nd125 = p.node_nuclt if p.node_nuclt is not None else (None, None)
        b = _mp_map(10)
        b += _mp_str("accession") + _mp_str(p.accession)
        b += _mp_str("gencode") + _mp_int(p.gencode)
        b += _mp_str("consensus") + _mp_str(p.consensus)
        b += _mp_str("core_size") + _mp_int(K)
        b += _mp_str("null_nuclt_dist") + nuclt(p.null_nuclt)
        b += _mp_str("null_emission") + f32(p.null_emission)
        b += _mp_str("bg_nuclt_dist") + nuclt(p.bg_nuclt)
        b += _mp_str("bg_emission") + f32(p.bg_emission)
        b += _mp_str("nodes") + _mp_map(3 * (K + 1))
        for i in range(K + 1):
            b += _mp_str("nuclt_dist") + nuclt(None if nd4 is None else (nd4[i], nd125[i]))
            b += _mp_str("trans") + f32(p.trans[i])
            b += _mp_str("emission") + f32(p.emission[i])
        b += _mp_str("BMk") + f32(p.BMk)
        recs.append(b)
    hdr = _mp_map(8)
    hdr += _mp_str("magic_number") + _mp_int(MAGIC_NUMBER)
    hdr += _mp_str("version") + _mp_int(1)
    hdr += _mp_str("entry_dist") + _mp_int(entry_dist)
    hdr += _mp_str("epsilon") + bytes([0xCA]) + struct.pack(">f", epsilon)
    hdr += _mp_str("abc") + abc(symbols, 4 if symbols == "ACGT" else 5)
    hdr += _mp_str("amino") + abc("ACDEFGHIKLMNPQRSTVWY", 2)
    hdr += _mp_str("has_ga") + bytes([0xC3])
    hdr += _mp_str("protein_sizes") + _mp_arr(len(recs)) + b"".join(_mp_int(len(r)) for r in recs)
    with open(path, "wb") as fh:
        fh.write(_mp_map(2) + _mp_str("header") + hdr + _mp_str("proteins") + _mp_arr(len(recs)))
        for r in recs:
            fh.write(r)
