"""GPU parity at the BASELINE.json sizes (C2: 100 MB DNA, C3: 200 MB English-like).  The oracle's suffix sort
still finishes in seconds on the box's host cores at these sizes, so SA and BWT are compared bit for bit; the
query path is checked against the oracle's FM index on a pattern sample and through size-independent
properties (every located position really holds the pattern; count == number of located positions)."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(cuda):
    from hkcsa import engine
    return engine


@pytest.mark.parametrize("kind,seed,n", [(1, 43, 100_000_000), (0, 42, 200_000_000)], ids=["C2_dna_100MB", "C3_eng_200MB"])
def test_build_and_query_at_baseline_size(E, kind, seed, n):
    import torch
    text = E.gen_text(kind, seed, n)
    text = torch.cat([text, torch.tensor([0x24], dtype=torch.uint8, device=text.device)])
    h_text = text.cpu().numpy()
    assert np.array_equal(h_text[:n], O.gen_text(kind, seed, n))             # generators agree at full size
    idx = E.DeviceIndex(text, sa_sample_rate=32)
    # ---- SA / BWT bit-exact against the oracle (parallel suffix sort on the host cores)
    want_sa = O.build_suffix_array(h_text)
    got_sa = idx.sa.cpu().numpy().astype(np.uint32)
    assert np.array_equal(got_sa, want_sa)
    want_bwt = O.bwt_transform(h_text, want_sa)
    assert np.array_equal(idx.bwt.cpu().numpy(), want_bwt)
    del got_sa
    # ---- C[] and occ at random positions
    cnt, Ct = O.build_count(h_text)
    assert idx.wt.count_table() == {chr(c): int(Ct[c]) for c in range(256) if cnt[c]}
    fm = O.FM(want_bwt)
    rng = np.random.RandomState(7)
    pos = rng.randint(0, n + 2, 2000).astype(np.int64)
    sym = rng.choice(np.flatnonzero(cnt), 2000).astype(np.uint8)
    got = idx.wt.rank(sym, pos).cpu().numpy()
    assert got.tolist() == [fm.rank(int(c), int(i)) for c, i in zip(sym, pos)]
    ap = rng.randint(0, n + 1, 5000).astype(np.int64)
    assert np.array_equal(idx.wt.access(ap).cpu().numpy(), want_bwt[ap])
    # ---- every level of the reference's wavelet tree (the left spine, csa/wavelet_tree.py:72-100) and its
    #      rank_support (:9-12) bit for bit at full size; select (:17-25) at 2 000 ranks per level
    _, spine = O.wt_spine(want_bwt)
    assert len(spine) >= 2 and len(spine[0]) == n + 1
    for level, bits in enumerate(spine):
        n_l = len(bits)
        assert n_l <= idx.wt.level_len(level)
        got_bits = idx.wt.bv_bits(level, 0, n_l).cpu().numpy()
        assert np.array_equal(got_bits, bits), f"level {level} bits"
        del got_bits
        rs = O.rank_support(bits)
        got_rs = idx.wt.bv_rank_range(level, 0, n_l + 1).cpu().numpy().astype(np.uint32)
        assert np.array_equal(got_rs, rs), f"level {level} rank_support"
        del got_rs
        ones = int(rs[-1])
        ks = np.unique(np.concatenate([[1, min(2, max(ones, 1)), max(ones, 1)], rng.randint(1, max(ones, 1) + 1, 2000)])).astype(np.int64)
        ks = ks[ks <= max(ones, 0)] if ones else ks[:0]
        if len(ks):
            got_sel = idx.wt.bv_select(level, ks).cpu().numpy()
            want_sel = np.searchsorted(rs, ks, side="left")          # smallest p with rank(p) >= k
            # a level holds more nodes than the reference's left-most one: select answers inside the spine prefix
            assert np.array_equal(got_sel, want_sel), f"level {level} select"
            assert want_sel.tolist()[:3] == [O.select(rs, int(k)) for k in ks[:3]]
        del rs
    # ---- count against the oracle on 20 k patterns; locate round trip
    pats, off = O.gen_patterns(44, 20_000, h_text[:n])
    d_p, d_o = torch.from_numpy(pats).cuda(), torch.from_numpy(off).cuda()
    lo, hi = idx.count_batch(d_p, d_o)
    w_lo, w_hi = fm.find_range_batch(pats, off)
    assert np.array_equal(lo.cpu().numpy(), w_lo) and np.array_equal(hi.cpu().numpy(), w_hi)
    # ---- the sampled Occ table (both layouts) gives the same ranges and the same LF walks at full size
    o_wt, p_wt = idx.locate_batch(d_p[: int(off[2000])], d_o[:2001], use_samples=True)
    for layout in (1, 0):
        idx.build_occ_table(5, layout=layout)
        lo2, hi2 = idx.count_batch(d_p, d_o, use_occ_table=True)
        assert torch.equal(lo, lo2) and torch.equal(hi, hi2)
        o_oc, p_oc = idx.locate_batch(d_p[: int(off[2000])], d_o[:2001], use_samples=True)
        assert torch.equal(o_wt, o_oc) and torch.equal(p_wt, p_oc)
        idx._occ = None
    o1, p1 = idx.locate_batch(d_p, d_o, use_samples=False)
    o2, p2 = idx.locate_batch(d_p, d_o, use_samples=True)
    assert torch.equal(o1, o2) and torch.equal(p1, p2)                       # LF walk == full SA
    o, p = o1.cpu().numpy(), p1.cpu().numpy()
    cntp = np.where(w_lo >= 0, w_hi - w_lo + 1, 0)
    assert np.array_equal(np.diff(o), cntp)
    for k in rng.choice(20_000, 300, replace=False):                         # located positions hold the pattern
        m = off[k + 1] - off[k]
        for q in p[o[k]:o[k + 1]][:5]:
            assert h_text[q:q + m].tobytes() == pats[off[k]:off[k + 1]].tobytes()
