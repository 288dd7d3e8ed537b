#!/usr/bin/env python
"""Freeze outputs of the REFERENCE ITSELF as golden fixtures.

Run in the authoring container only (needs /root/reference, which does not
exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Writes tests/golden/golden_ref.npz (+ golden_ref.json.gz for the scalar / dict
results).  The reference holds no asserting tests or golden files of its own
(SURVEY.md section 4), so these fixtures -- produced by importing and calling
the reference's functions unmodified -- are what pins the oracle and, through
it, the CUDA path.

Reference entry points exercised (paths relative to /root/reference):
  csa/suffix_array.py:131  build_suffix_array      csa/suffix_array.py:46 ksa
  csa/bwt.py:3             bwt_transform
  utils/utils.py:16,26     build_count, build_occ
  csa/wavelet_tree.py:5    SuccinctRankSelect      :27 GolombRiceEncoder   :65 WaveletTree
  csa/enhanced_fm_index.py:7 EnhancedFMIndex (find_range / find / rank)
  main.py:6                FMIndex (identical to csa/csa.py:6, whose module import fails)
"""
import contextlib
import gzip
import importlib.util
import io
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("HKCSA_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

with contextlib.redirect_stdout(io.StringIO()):  # import-time demo prints
    from csa.suffix_array import build_suffix_array, ksa
    from csa.bwt import bwt_transform
    from csa.wavelet_tree import WaveletTree, SuccinctRankSelect, GolombRiceEncoder
    from csa.enhanced_fm_index import EnhancedFMIndex
    from utils.utils import build_count, build_occ
    spec = importlib.util.spec_from_file_location("ref_main", os.path.join(REF, "main.py"))
    ref_main = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_main)
FMIndex = ref_main.FMIndex

from oracle import oracle as O  # only for the seeded text generators (workload, not results)


def u8(s: str) -> np.ndarray:
    return np.frombuffer(s.encode("latin-1"), dtype=np.uint8)


def cases():
    lit = ["", "a", "aaaa", "ab", "banana", "mississippi", "abracadabra", "abcd", "abcde",
           "this is an example text", "a" * 100 + "bbb" + "zz" + "y", " !#ab ba!# a",
           "a$b$a$", "mississippi$" * 1000]
    for i, s in enumerate(lit):
        yield f"lit{i:02d}", s
    rng = np.random.RandomState(1234)
    for n in (1024, 16384):
        yield f"eng96_{n}", O.gen_text(O.ENG96, 42, n).tobytes().decode("latin-1")
        yield f"dna4_{n}", O.gen_text(O.DNA4, 43, n).tobytes().decode("latin-1")
        yield f"bin_{n}", bytes(rng.choice(np.frombuffer(b"ab", dtype=np.uint8), n)).decode("latin-1")
        # sigma=256 without '$' would be 255; keep all 256 byte values: '$' inside the
        # text is legal for SA/BWT/WT (EnhancedFMIndex part is skipped for it below).
        yield f"byte_{n}", bytes(rng.randint(0, 256, n).astype(np.uint8)).decode("latin-1")
    yield "eng96_65536", O.gen_text(O.ENG96, 42, 65536).tobytes().decode("latin-1")
    yield "dna4_65536", O.gen_text(O.DNA4, 43, 65536).tobytes().decode("latin-1")


def patterns_for(text: str, rnd: random.Random):
    pats = ["", "x", "$", "\x00"]
    n = len(text)
    if n:
        pats += [text[:1], text[-1:], text[-1:] + "$", text, text + "$", text[: n // 2]]
        for _ in range(24):
            m = rnd.randint(1, min(20, n))
            s = rnd.randint(0, n - m)
            p = text[s:s + m]
            pats.append(p)
            if rnd.random() < 0.5:
                k = rnd.randrange(m)
                p2 = p[:k] + rnd.choice(sorted(set(text))) + p[k + 1:]
                pats.append(p2)
    return pats


def main():
    arrays, meta = {}, {}
    rnd = random.Random(99)
    for name, text in cases():
        n = len(text)
        info = {"n": n}
        arrays[f"{name}/text"] = u8(text)
        # ---- a1 / a2 / a3 on the bare text
        sa = build_suffix_array(text)
        arrays[f"{name}/sa"] = np.asarray(sa, dtype=np.uint32)
        arrays[f"{name}/bwt"] = u8(bwt_transform(text, sa))
        try:
            info["ksa_equal"] = (ksa(text) == sa)
        except Exception as e:  # a2: the sketch raises on most inputs
            info["ksa_equal"] = type(e).__name__
        # ---- a4 / a5
        info["count"] = {str(ord(k)): v for k, v in build_count(text).items()}
        if n <= 2048:
            occ = build_occ(text)
            for k, v in occ.items():
                arrays[f"{name}/occ/{ord(k)}"] = np.asarray(v, dtype=np.uint32)
        # ---- a7 / a6 / a8 / a9 / a10
        if n > 0:
            with contextlib.redirect_stdout(io.StringIO()):
                wt = WaveletTree(text)
            info["wt_levels"] = len(wt.tree)
            info["wt_m"] = wt.m
            info["wt_alphabet_after"] = [ord(c) for c in wt.alphabet]
            info["wt_left"] = [[ord(c) for c in lvl[1]] for lvl in wt.tree]
            info["wt_right"] = [[ord(c) for c in lvl[2]] for lvl in wt.tree]
            info["wt_next_len"] = [len(lvl[3]) for lvl in wt.tree]
            for l, rs in enumerate(wt.rank_structures):
                arrays[f"{name}/wt/{l}/bits"] = np.packbits(rs.bit_vector)
                arrays[f"{name}/wt/{l}/nbits"] = np.asarray([rs.n], dtype=np.int64)
                arrays[f"{name}/wt/{l}/golomb"] = np.packbits(np.asarray(wt.tree[l][0], dtype=np.uint8))
                arrays[f"{name}/wt/{l}/golomb_len"] = np.asarray([len(wt.tree[l][0])], dtype=np.int64)
                arrays[f"{name}/wt/{l}/next"] = u8("".join(wt.tree[l][3]))
                if rs.n <= 2048:
                    arrays[f"{name}/wt/{l}/rank_support"] = rs.rank_support.copy()
                ks = sorted({0, 1, 2, int(rs.rank_support[-1]), int(rs.rank_support[-1]) + 1,
                             *[rnd.randint(0, max(1, int(rs.rank_support[-1]))) for _ in range(8)]})
                arrays[f"{name}/wt/{l}/select_k"] = np.asarray(ks, dtype=np.int64)
                arrays[f"{name}/wt/{l}/select_v"] = np.asarray([rs.select(k) for k in ks], dtype=np.int64)
            if len(wt.tree):
                last = wt.rank_structures[-1]
                qs = [i for i in (0, 1, last.n - 1) if 0 <= i < last.n]
                info["wt_rank_quirk"] = [[i, int(wt.rank("a", i))] for i in qs]
                info["wt_select_quirk"] = [[k, int(wt.select("a", k))] for k in (0, 1, 2)]
            else:
                info["wt_rank_quirk"] = [[0, int(wt.rank("a", 0))]]
                info["wt_select_quirk"] = [[1, int(wt.select("a", 1))]]
            info["wt_decompress"] = wt.decompress(wt.compress())
        # ---- a11-a13: EnhancedFMIndex (text + '$')
        if n <= 16384 or name.endswith("65536"):
            fm = EnhancedFMIndex(text)
            arrays[f"{name}/fm/sa"] = np.asarray(fm.suffix_array, dtype=np.uint32)
            arrays[f"{name}/fm/bwt"] = u8(fm.bwt)
            info["fm_count"] = {str(ord(k)): v for k, v in fm.count.items()}
            pats = patterns_for(text, rnd)
            res = []
            for p in pats:
                l, r = fm.find_range(p)
                f = fm.find(p)
                res.append({"p": [ord(c) for c in p], "l": l, "r": r,
                            "find": f if len(f) <= 64 else None,
                            "find_sorted_sum": int(sum(f)), "find_len": len(f)})
            info["fm_queries"] = res
            rq = []
            for _ in range(16):
                c = rnd.choice(sorted(set(fm.text)))
                i = rnd.randint(0, len(fm.text))
                rq.append([ord(c), i, int(fm.rank(c, i))])
            rq.append([ord("\x01"), 3, int(fm.rank("\x01", 3))])
            info["fm_rank"] = rq
        # ---- a14: FMIndex (main.py copy)
        if n <= 16384:
            fi = FMIndex(text)
            info["fmindex_rank_keys"] = [ord(k) for k in fi.rank.keys()]
            if n <= 2048:
                for k, v in fi.rank.items():
                    arrays[f"{name}/fmindex/rank/{ord(k)}"] = np.asarray(v, dtype=np.uint32)
            info["fmindex_backward_search_len"] = len(fi.backward_search(text[:3]))
            info["fmindex_find_pattern_is_sa"] = (fi.find_pattern(text[:3]) == fi.suffix_array)
        meta[name] = info
        print(name, n, file=sys.stderr)
    # a8 Golomb m spot values (csa/wavelet_tree.py:33-38)
    enc = GolombRiceEncoder([1, 0])
    gm = []
    for ones, total in [(1, 2), (1, 4), (1, 8), (3, 24), (1, 3), (1, 5), (1, 2 ** 20), (5, 5), (0, 7),
                        (7, 100), (33, 1000), (1, 1), (2, 3), (1023, 1 << 20), (1025, 1 << 20)]:
        gm.append([ones, total, enc.compute_dynamic_m(ones, total)])
    meta["_golomb_m"] = gm
    # a6 on a literal bitmap
    rs = SuccinctRankSelect([0, 1, 1, 0, 0, 1])
    meta["_rs_literal"] = {"rank": [int(rs.rank(i)) for i in range(7)],
                           "select": [int(rs.select(k)) for k in range(6)]}
    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **arrays)
    with gzip.open(os.path.join(HERE, "golden_ref.json.gz"), "wt", compresslevel=9) as f:
        json.dump(meta, f, separators=(",", ":"))
    print("wrote", len(arrays), "arrays,", len(meta), "cases", file=sys.stderr)


if __name__ == "__main__":
    main()
