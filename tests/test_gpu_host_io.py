"""Host text in / index out: hkcsa_h2d_staged / hkcsa_d2h_staged (pinned ring filled by several host threads) and the
str / pattern-list paths of the drop-in API that ride on them (EnhancedFMIndex takes a Python str,
csa/enhanced_fm_index.py:8-9; find_range takes str patterns, :21-32).  Byte-exact round trips; the drop-in results are
compared with the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
CHUNK = 2 << 20


@pytest.fixture(scope="module", autouse=True)
def _need_cuda(cuda):
    return cuda


@pytest.mark.parametrize("threads", [1, 3, 16, 0])
@pytest.mark.parametrize("n", [1, 17, CHUNK - 1, CHUNK, CHUNK + 1, 9 * CHUNK + 12345, 37_000_003])
def test_staged_round_trip(n, threads):
    import torch
    from hkcsa import _lib, engine as E
    L = _lib.load()
    src = np.random.RandomState(n % 9973).randint(0, 256, size=n, dtype=np.uint8)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.hkcsa_h2d_staged(d.data_ptr(), src.ctypes.data, n, threads, st))
    want = src.copy()
    src[:] = 0                       # the source may be released as soon as the call returns
    assert np.array_equal(d.cpu().numpy(), want)
    back = np.empty(n, dtype=np.uint8)
    _lib.check(L.hkcsa_d2h_staged(back.ctypes.data, d.data_ptr(), n, threads, st))
    assert np.array_equal(back, want)   # d2h returns with the data in place
    # back to back on the same ring: the second copy must not overtake the first one's DMAs
    a = np.full(n, 7, dtype=np.uint8)
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    _lib.check(L.hkcsa_h2d_staged(d.data_ptr(), a.ctypes.data, n, threads, st))
    _lib.check(L.hkcsa_h2d_staged(d2.data_ptr(), want.ctypes.data, n, threads, st))
    assert bool((d == 7).all()) and np.array_equal(d2.cpu().numpy(), want)


def test_staged_copy_argument_checks():
    from hkcsa import _lib
    L = _lib.load()
    assert L.hkcsa_h2d_staged(0, 0, 0, 0, 0) == _lib.OK            # nothing to do
    assert L.hkcsa_h2d_staged(0, 0, 16, 0, 0) == _lib.EINVAL
    assert L.hkcsa_d2h_staged(0, 0, 16, 0, 0) == _lib.EINVAL


def test_to_device_u8_sources_agree():
    """ASCII str (staged from the str's own buffer), latin-1 str (encoded), bytes, numpy: same device bytes; the tail
    lands behind them."""
    from hkcsa import engine as E
    rng = np.random.RandomState(3)
    ascii_b = rng.randint(32, 127, size=5 * CHUNK + 77, dtype=np.uint8).tobytes()
    latin_b = rng.randint(1, 256, size=300_001, dtype=np.uint8).tobytes()
    for raw in (ascii_b, latin_b, b"a", b""):
        text = raw.decode("latin-1")
        for tail in (b"", b"$"):
            want = np.frombuffer(raw + tail, dtype=np.uint8)
            for src in (text, raw, np.frombuffer(raw, dtype=np.uint8), bytearray(raw)):
                got = E.to_device_u8(src, tail=tail).cpu().numpy()
                assert np.array_equal(got, want)
    with pytest.raises(ValueError):
        E.to_device_u8("snow ☃")


def test_symbol_map_ascii_text_is_not_encoded():
    from hkcsa import engine as E
    text = "GATTACA" * 1000
    m = E.SymbolMap(text, extra="$")
    assert m.identity and m.host_bytes(text) is text and m.encode("$") == b"$"
    lat = "caf\xe9" * 10
    m2 = E.SymbolMap(lat, extra="$")
    assert m2.identity and m2.host_bytes(lat) == lat.encode("latin-1")


def test_str_index_and_pattern_list_fast_path_match_oracle():
    """EnhancedFMIndex(str) on a multi-chunk ASCII text + find_range_batch on a list of str patterns (packed by one
    join) against the oracle's ranges; a mixed / non-latin-1 list takes the per-pattern path and agrees."""
    from csa.enhanced_fm_index import EnhancedFMIndex
    rng = np.random.RandomState(11)
    raw = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=2 * CHUNK + 4321).tobytes()
    text = raw.decode("latin-1")
    fm = EnhancedFMIndex(text)
    n = len(raw) + 1
    sa = O.build_suffix_array(raw + b"$")
    ofm = O.FM(O.bwt_transform(raw + b"$", sa))
    starts = rng.randint(0, len(raw) - 40, size=3000)
    lens = rng.randint(1, 33, size=3000)
    pats = [text[a:a + b] for a, b in zip(starts.tolist(), lens.tolist())]
    pats[5] = "ACGTN"                # miss
    pats[6] = ""                     # empty pattern: the whole range (csa/enhanced_fm_index.py:22-23)
    lo, hi = fm.find_range_batch(pats)
    enc = [p.encode("latin-1") for p in pats]
    off = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    olo, ohi = ofm.find_range_batch(np.frombuffer(b"".join(enc), dtype=np.uint8), off)
    assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
    mixed = list(pats[:50])
    mixed[3] = pats[3].encode("latin-1")      # bytes among the str patterns
    mixed[4] = "AC☃"                     # beyond latin-1: a miss like an unseen symbol (:27-28)
    lo2, hi2 = fm.find_range_batch(mixed)
    assert (lo2[4], hi2[4]) == (-1, -1)
    keep = [k for k in range(50) if k != 4]
    assert np.array_equal(lo2[keep], olo[keep]) and np.array_equal(hi2[keep], ohi[keep])
    assert fm.text == text + "$" and len(fm.text) == n


def test_to_host_matches_torch_copy():
    import torch
    from hkcsa import engine as E
    g = torch.Generator(device="cuda").manual_seed(2)
    for dtype, n in ((torch.int32, 5_000_003), (torch.int64, 3_000_001), (torch.uint8, 40_000_007), (torch.int32, 100)):
        t = torch.randint(0, 100, (n,), device="cuda", generator=g).to(dtype)
        a = E.to_host(t)
        assert a.dtype == t.cpu().numpy().dtype and np.array_equal(a, t.cpu().numpy())
    t2 = torch.arange(6_000_000, device="cuda", dtype=torch.int32).view(2000, 3000).t()   # non-contiguous
    assert np.array_equal(E.to_host(t2), t2.cpu().numpy())


def test_save_load_round_trip_through_staged_copies(tmp_path):
    """save() ships the blobs through hkcsa_d2h_staged (> 8 MB each): the loaded index answers like the built one."""
    import torch
    from hkcsa import engine as E
    text = E.gen_text(E.ENG96, 9, 30_000_000)
    text = torch.cat([text, torch.tensor([36], dtype=torch.uint8, device=text.device)])
    idx = E.DeviceIndex(text, sa_sample_rate=32)
    path = str(tmp_path / "idx.npz")
    idx.save(path)
    back = E.DeviceIndex.load(path)
    pat, off = E.gen_patterns(5, 20_000, text[:-1], torch.unique(text[:-1]))
    lo, hi = idx.count_batch(pat, off)
    lo2, hi2 = back.count_batch(pat, off)
    assert torch.equal(lo, lo2) and torch.equal(hi, hi2)
    assert torch.equal(idx.wt.blob, back.wt.blob) and torch.equal(idx.ssa.blob, back.ssa.blob)
