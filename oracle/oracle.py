"""ctypes/numpy front-end of the CPU oracle (oracle/hkcsa_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs.  Nothing under
``high-order-entropy-compressed-suffix-array_b200/`` imports this module.

Parity pin: every function here is checked against outputs of the reference
itself (executed by tests/golden/make_golden.py in the authoring container and
frozen under tests/golden/) by tests/test_oracle_golden.py.

Each wrapper names the reference function it restates (file:line relative to
the reference repository root); the arithmetic lives in the C file.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhkcsa_oracle.so")

ENG96, DNA4 = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "hkcsa_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u32p, u64p, i64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64))
        L.hko_gen_text.argtypes = [C.c_int, C.c_uint64, C.c_uint64, u8p]
        L.hko_gen_text.restype = C.c_int
        L.hko_pattern_lengths.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, u32p]
        L.hko_pattern_lengths.restype = None
        L.hko_pattern_fill.argtypes = [C.c_uint64, C.c_uint64, u8p, C.c_uint64, u8p, C.c_uint32, i64p, u8p]
        L.hko_pattern_fill.restype = None
        L.hko_sa_build.argtypes = [u8p, C.c_uint64, u32p, C.c_int]
        L.hko_sa_build.restype = C.c_int
        L.hko_bwt.argtypes = [u8p, u32p, C.c_uint64, u8p]
        L.hko_bwt.restype = None
        L.hko_count_table.argtypes = [u8p, C.c_uint64, u64p, u64p]
        L.hko_count_table.restype = None
        L.hko_occ_dense.argtypes = [u8p, C.c_uint64, C.c_uint8, u32p]
        L.hko_occ_dense.restype = None
        L.hko_fm_new.argtypes = [u8p, C.c_uint64]
        L.hko_fm_new.restype = C.c_void_p
        L.hko_fm_free.argtypes = [C.c_void_p]
        L.hko_fm_free.restype = None
        L.hko_fm_rank.argtypes = [C.c_void_p, C.c_uint8, C.c_uint64]
        L.hko_fm_rank.restype = C.c_uint64
        L.hko_find_range_batch.argtypes = [C.c_void_p, u8p, i64p, C.c_uint64, i64p, i64p, C.c_int]
        L.hko_find_range_batch.restype = None
        L.hko_wt_spine.argtypes = [u8p, C.c_uint64, u8p, u64p, u8p, C.POINTER(C.c_int)]
        L.hko_wt_spine.restype = C.c_int
        L.hko_rank_support.argtypes = [u8p, C.c_uint64, u32p]
        L.hko_rank_support.restype = None
        L.hko_select.argtypes = [u32p, C.c_uint64, C.c_uint64]
        L.hko_select.restype = C.c_uint64
        L.hko_golomb_m.argtypes = [C.c_uint64, C.c_uint64]
        L.hko_golomb_m.restype = C.c_uint32
        L.hko_golomb_encode.argtypes = [u8p, C.c_uint64, C.c_uint32, u8p]
        L.hko_golomb_encode.restype = C.c_uint64
        L.hko_symbol_positions.argtypes = [u8p, C.c_uint64, u32p, u64p]
        L.hko_symbol_positions.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def as_u8(text) -> np.ndarray:
    """str (latin-1, utils/data_loader.py:4) / bytes / array -> contiguous uint8."""
    if isinstance(text, str):
        text = text.encode("latin-1")
    if isinstance(text, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(text), dtype=np.uint8).copy()
    return np.ascontiguousarray(text, dtype=np.uint8)


# ------------------------------------------------------------------ workload
def gen_text(kind: int, seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    rc = lib().hko_gen_text(kind, seed, n, _p(out, C.c_uint8))
    if rc:
        raise ValueError("unknown text kind")
    return out


def gen_patterns(seed: int, P: int, text: np.ndarray, min_len: int = 8, max_len: int = 64):
    """(bytes uint8[sum m], offsets int64[P+1]); substrings, half of them with one substitution."""
    text = as_u8(text)
    n = len(text)
    lens = np.empty(P, dtype=np.uint32)
    lib().hko_pattern_lengths(seed, P, min_len, max_len, n, _p(lens, C.c_uint32))
    off = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    alpha = np.unique(text)
    out = np.empty(int(off[-1]), dtype=np.uint8)
    lib().hko_pattern_fill(seed, P, _p(text, C.c_uint8), n, _p(alpha, C.c_uint8), len(alpha),
                           _p(off, C.c_int64), _p(out, C.c_uint8))
    return out, off


# ------------------------------------------------------------------ a1, a3, a4, a5
def build_suffix_array(text, threads: int = 0) -> np.ndarray:
    """csa/suffix_array.py:131-134."""
    t = as_u8(text)
    sa = np.empty(len(t), dtype=np.uint32)
    if len(t):
        lib().hko_sa_build(_p(t, C.c_uint8), len(t), _p(sa, C.c_uint32), threads)
    return sa


def bwt_transform(text, sa) -> np.ndarray:
    """csa/bwt.py:3-13."""
    t = as_u8(text)
    sa = np.ascontiguousarray(sa, dtype=np.uint32)
    out = np.empty(len(t), dtype=np.uint8)
    if len(t):
        lib().hko_bwt(_p(t, C.c_uint8), _p(sa, C.c_uint32), len(t), _p(out, C.c_uint8))
    return out


def build_count(text):
    """utils/utils.py:16-24 -> (cnt[256], C[256]) over byte values."""
    t = as_u8(text)
    cnt = np.zeros(256, dtype=np.uint64)
    Ct = np.zeros(256, dtype=np.uint64)
    lib().hko_count_table(_p(t, C.c_uint8), len(t), _p(cnt, C.c_uint64), _p(Ct, C.c_uint64))
    return cnt, Ct


def count_dict(text) -> dict:
    """build_count as the reference's dict {chr: C} over present symbols."""
    cnt, Ct = build_count(text)
    return {chr(c): int(Ct[c]) for c in range(256) if cnt[c]}


def occ_dense(bwt, c: int) -> np.ndarray:
    """utils/utils.py:26-32, one symbol's column: occ[c][0..n]."""
    b = as_u8(bwt)
    out = np.empty(len(b) + 1, dtype=np.uint32)
    lib().hko_occ_dense(_p(b, C.c_uint8), len(b), c, _p(out, C.c_uint32))
    return out


class FM:
    """EnhancedFMIndex over an already-built BWT (csa/enhanced_fm_index.py:7-40)."""

    def __init__(self, bwt):
        self.bwt = as_u8(bwt)
        self._h = lib().hko_fm_new(_p(self.bwt, C.c_uint8), len(self.bwt))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hko_fm_free(self._h)
            self._h = None

    def rank(self, c: int, i: int) -> int:
        return int(lib().hko_fm_rank(self._h, c, i))

    def find_range_batch(self, pats: np.ndarray, off: np.ndarray, threads: int = 0):
        pats = np.ascontiguousarray(pats, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        P = len(off) - 1
        lo = np.empty(P, dtype=np.int64)
        hi = np.empty(P, dtype=np.int64)
        if pats.size == 0:
            pats = np.zeros(1, dtype=np.uint8)
        lib().hko_find_range_batch(self._h, _p(pats, C.c_uint8), _p(off, C.c_int64), P,
                                   _p(lo, C.c_int64), _p(hi, C.c_int64), threads)
        return lo, hi

    def find_range(self, query):
        q = as_u8(query)
        lo, hi = self.find_range_batch(q, np.array([0, len(q)], dtype=np.int64), 1)
        return int(lo[0]), int(hi[0])


# ------------------------------------------------------------------ a6, a7, a8
def wt_spine(text):
    """csa/wavelet_tree.py:72-100 -> (alphabet uint8[sigma], [bitmap uint8[...] per level])."""
    t = as_u8(text)
    n = len(t)
    bits = np.empty(max(1, 8 * n), dtype=np.uint8)
    lens = np.zeros(8, dtype=np.uint64)
    alpha = np.zeros(256, dtype=np.uint8)
    sigma = C.c_int(0)
    tt = t if n else np.zeros(1, dtype=np.uint8)
    L = lib().hko_wt_spine(_p(tt, C.c_uint8), n, _p(bits, C.c_uint8), _p(lens, C.c_uint64),
                           _p(alpha, C.c_uint8), C.byref(sigma))
    out, w = [], 0
    for l in range(L):
        out.append(bits[w:w + int(lens[l])].copy())
        w += int(lens[l])
    return alpha[:sigma.value].copy(), out


def rank_support(bits) -> np.ndarray:
    """csa/wavelet_tree.py:9-12."""
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    rs = np.empty(len(b) + 1, dtype=np.uint32)
    bb = b if len(b) else np.zeros(1, dtype=np.uint8)
    lib().hko_rank_support(_p(bb, C.c_uint8), len(b), _p(rs, C.c_uint32))
    return rs


def select(rs: np.ndarray, k: int) -> int:
    """csa/wavelet_tree.py:17-25."""
    rs = np.ascontiguousarray(rs, dtype=np.uint32)
    return int(lib().hko_select(_p(rs, C.c_uint32), len(rs) - 1, k))


def golomb_m(ones: int, total: int) -> int:
    """csa/wavelet_tree.py:33-38."""
    return int(lib().hko_golomb_m(ones, total))


def golomb_encode(bits, m: int | None = None) -> np.ndarray:
    """csa/wavelet_tree.py:40-63 with m from :28-31 unless given."""
    b = np.ascontiguousarray(bits, dtype=np.uint8)
    if m is None:
        m = golomb_m(int(b.sum()), len(b))
    bb = b if len(b) else np.zeros(1, dtype=np.uint8)
    size = int(lib().hko_golomb_encode(_p(bb, C.c_uint8), len(b), m, None))
    out = np.empty(max(1, size), dtype=np.uint8)
    lib().hko_golomb_encode(_p(bb, C.c_uint8), len(b), m, _p(out, C.c_uint8))
    return out[:size]


def symbol_positions(bwt):
    """csa/csa.py:13-19: positions of every symbol in the BWT, ascending, grouped by byte."""
    b = as_u8(bwt)
    pos = np.empty(max(1, len(b)), dtype=np.uint32)
    start = np.zeros(257, dtype=np.uint64)
    bb = b if len(b) else np.zeros(1, dtype=np.uint8)
    lib().hko_symbol_positions(_p(bb, C.c_uint8), len(b), _p(pos, C.c_uint32), _p(start, C.c_uint64))
    return pos[:len(b)], start
