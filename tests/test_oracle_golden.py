"""Pins the CPU oracle (oracle/hkcsa_oracle.c) against outputs of the reference
itself, frozen by tests/golden/make_golden.py.  CPU only."""
import numpy as np
import pytest

from conftest import golden_case_names
from oracle import oracle as O

CASES = golden_case_names()


@pytest.mark.parametrize("name", CASES)
def test_sa_bwt_count(golden, name):
    text = golden.text(name)
    sa = O.build_suffix_array(text, threads=2)
    assert np.array_equal(sa, golden.get(f"{name}/sa"))                      # a1
    assert O.bwt_transform(text, sa).tobytes() == golden.get(f"{name}/bwt").tobytes()  # a3
    want = {chr(int(k)): v for k, v in golden.meta[name]["count"].items()}
    assert O.count_dict(text) == want                                        # a4


@pytest.mark.parametrize("name", [c for c in CASES if golden_case_names and True])
def test_occ_dense(golden, name):
    text = golden.text(name)
    keys = [k for k in golden.arr.files if k.startswith(f"{name}/occ/")]
    for k in keys:                                                           # a5 (n <= 2048 cases)
        c = int(k.rsplit("/", 1)[1])
        assert np.array_equal(O.occ_dense(text, c), golden.get(k))


@pytest.mark.parametrize("name", CASES)
def test_wavelet_spine(golden, name):
    meta = golden.meta[name]
    if meta["n"] == 0:
        return
    text = golden.text(name)
    alpha, levels = O.wt_spine(text)
    assert len(levels) == meta["wt_levels"]                                  # a7
    a = list(alpha)
    for l, bits in enumerate(levels):
        n_l = int(golden.get(f"{name}/wt/{l}/nbits")[0])
        want = np.unpackbits(golden.get(f"{name}/wt/{l}/bits"))[:n_l]
        assert np.array_equal(bits, want)
        mid = len(a) // 2
        assert [int(x) for x in a[:mid]] == meta["wt_left"][l]
        assert [int(x) for x in a[mid:]] == meta["wt_right"][l]
        a = a[:mid]
        rs = O.rank_support(bits)                                            # a6
        if golden.has(f"{name}/wt/{l}/rank_support"):
            assert np.array_equal(rs, golden.get(f"{name}/wt/{l}/rank_support"))
        for k, v in zip(golden.get(f"{name}/wt/{l}/select_k"), golden.get(f"{name}/wt/{l}/select_v")):
            assert O.select(rs, int(k)) == int(v)
        m = O.golomb_m(int(bits.sum()), len(bits))                           # a8
        if l == 0:
            assert m == meta["wt_m"]
        g_len = int(golden.get(f"{name}/wt/{l}/golomb_len")[0])
        want_g = np.unpackbits(golden.get(f"{name}/wt/{l}/golomb"))[:g_len]
        assert np.array_equal(O.golomb_encode(bits, m), want_g)
    assert [int(x) for x in a] == meta["wt_alphabet_after"]


def test_golomb_m_values(golden):
    for ones, total, m in golden.meta["_golomb_m"]:
        assert O.golomb_m(ones, total) == m


def test_rank_select_literal(golden):
    rs = O.rank_support([0, 1, 1, 0, 0, 1])
    assert [int(x) for x in rs] == golden.meta["_rs_literal"]["rank"]
    assert [O.select(rs, k) for k in range(6)] == golden.meta["_rs_literal"]["select"]


@pytest.mark.parametrize("name", [c for c in CASES])
def test_fm_index(golden, name):
    meta = golden.meta[name]
    if "fm_queries" not in meta:
        return
    text = golden.text(name) + b"$"                                          # a11
    sa = O.build_suffix_array(text, threads=2)
    assert np.array_equal(sa, golden.get(f"{name}/fm/sa"))
    bwt = O.bwt_transform(text, sa)
    assert bwt.tobytes() == golden.get(f"{name}/fm/bwt").tobytes()
    want_c = {chr(int(k)): v for k, v in meta["fm_count"].items()}
    assert O.count_dict(text) == want_c
    fm = O.FM(bwt)
    for q in meta["fm_queries"]:                                             # a12, a13
        p = bytes(q["p"])
        l, r = fm.find_range(p)
        assert (l, r) == (q["l"], q["r"]), (name, p)
        found = [] if l < 0 else [int(x) for x in sa[l:r + 1]]
        assert len(found) == q["find_len"] and sum(found) == q["find_sorted_sum"]
        if q["find"] is not None:
            assert found == q["find"]
    for c, i, v in meta["fm_rank"]:
        assert fm.rank(c, i) == v


@pytest.mark.parametrize("name", CASES)
def test_fmindex_position_lists(golden, name):
    meta = golden.meta[name]
    if "fmindex_rank_keys" not in meta or meta["n"] > 2048:
        return
    bwt = golden.get(f"{name}/bwt")
    pos, start = O.symbol_positions(bwt)                                     # a14
    for c in meta["fmindex_rank_keys"]:
        want = golden.get(f"{name}/fmindex/rank/{c}")
        assert np.array_equal(pos[int(start[c]):int(start[c + 1])], want)


def test_generators_are_deterministic():
    a = O.gen_text(O.ENG96, 42, 200_000)
    b = O.gen_text(O.ENG96, 42, 200_000)
    assert np.array_equal(a, b) and 0x24 not in a
    assert len(np.unique(a)) > 60
    d = O.gen_text(O.DNA4, 43, 100_000)
    assert set(np.unique(d).tolist()) == {65, 67, 71, 84}
    pats, off = O.gen_patterns(44, 1000, a)
    lens = np.diff(off)
    assert lens.min() >= 8 and lens.max() <= 64 and pats.size == off[-1]
