"""world_size-2 gloo tests (CPU) of the multi-GPU query path's host logic: pattern sharding by
total symbols, all-gather of (lo, hi), and index replication.  The per-rank search is stubbed by
the CPU oracle here (test infrastructure); on the GPU box it is libhkcsa's count kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def test_shard_bounds_balance_by_symbols():
    from hkcsa.dist import shard_bounds
    rng = np.random.RandomState(1)
    lens = rng.randint(8, 65, 10_000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    for world in (1, 2, 3, 8):
        b = shard_bounds(off, world)
        assert b[0][0] == 0 and b[-1][1] == 10_000
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        work = [off[e] - off[s] for s, e in b]
        assert max(work) - min(work) <= 2 * 64
    assert shard_bounds(np.array([0], dtype=np.int64), 4) == [(0, 0)] * 4
    assert shard_bounds(np.array([0, 0, 0, 0], dtype=np.int64), 2) == [(0, 1), (1, 3)]   # empty patterns
    skew = np.array([0, 1000, 1001, 1002, 1003], dtype=np.int64)
    b = shard_bounds(skew, 2)
    assert b[0][1] >= 1 and b[-1][1] == 4
    # offsets given as a tensor are searched with torch (on the GPU box: on the device): same cuts
    for arr in (off, np.array([0], dtype=np.int64), np.array([0, 0, 0, 0], dtype=np.int64), skew):
        for world in (1, 2, 3, 8):
            assert shard_bounds(torch.from_numpy(arr), world) == shard_bounds(arr, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, text, pats, off, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hkcsa.dist import sharded_count
        t = text + b"$"
        fm = O.FM(O.bwt_transform(t, O.build_suffix_array(t, threads=1)))

        def count_fn(p, o):     # stand-in for DeviceIndex.count_batch on this rank
            lo, hi = fm.find_range_batch(p.numpy(), o.numpy(), threads=1)
            return torch.from_numpy(lo), torch.from_numpy(hi)

        lo, hi = sharded_count(count_fn, torch.from_numpy(pats), torch.from_numpy(off))
        w_lo, w_hi = fm.find_range_batch(pats, off, threads=1)
        ok = np.array_equal(lo.numpy(), w_lo) and np.array_equal(hi.numpy(), w_hi)
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_count_gloo(world):
    text = O.gen_text(O.ENG96, 42, 20_000).tobytes()
    pats, off = O.gen_patterns(44, 501, np.frombuffer(text, dtype=np.uint8))
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), text, pats, off, out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}


def test_balanced_bucket_ranges():
    from hkcsa.dist_sa import balanced_bucket_ranges
    rng = np.random.RandomState(3)
    h = rng.randint(0, 1000, 65536)
    for parts in (1, 2, 5, 8):
        r = balanced_bucket_ranges(h, parts)
        assert r[0][0] == 0 and r[-1][1] == 65536 and all(r[i][1] == r[i + 1][0] for i in range(parts - 1))
        loads = [int(h[a:b].sum()) for a, b in r]
        assert sum(loads) == int(h.sum()) and max(loads) - min(loads) <= 2 * int(h.max())
    assert balanced_bucket_ranges(np.zeros(16, dtype=np.int64), 4)[-1][1] == 16
    one = np.zeros(16, dtype=np.int64); one[7] = 100                       # one giant bucket cannot be split
    r = balanced_bucket_ranges(one, 4)
    assert sum(int(one[a:b].sum()) for a, b in r) == 100


def test_exchange_layout_regions_tile_every_destination():
    """Host logic of the distributed build's bucket exchange: the (source, destination) count matrix derived from the
    all-gathered bucket histograms gives every source a private region in every destination's receive array."""
    from hkcsa.dist_sa import exchange_layout
    rng = np.random.RandomState(3)
    for world in (1, 2, 3, 8):
        hist_all = rng.randint(0, 50, size=(world, 4096)).astype(np.int64)
        hist_all[:, 100] += 5000                       # one heavy bucket
        cuts, cnt, counts, slice_off = exchange_layout(hist_all, world)
        assert cuts[0] == 0 and cuts[-1] == 4096 and np.all(np.diff(cuts) >= 0)
        assert cnt.shape == (world, world) and cnt.sum() == hist_all.sum()
        assert np.array_equal(counts, cnt.sum(0)) and slice_off[-1] == hist_all.sum()
        for d in range(world):
            base = [int(cnt[:s_, d].sum()) for s_ in range(world)]
            ends = [base[s_] + int(cnt[s_, d]) for s_ in range(world)]
            assert base[0] == 0 and ends[-1] == counts[d] and all(ends[i] == base[i + 1] for i in range(world - 1))
        # balance: no destination exceeds the ideal share by more than the heaviest bucket
        assert counts.max() <= hist_all.sum() / world + hist_all.sum(0).max()


def _program_worker(rank, world, port, out):
    """The collective steps of the distributed build's rank program (hkcsa.dist_sa._torch_run) over gloo: a toy
    program yields the same requests the real one does -- sizes, text blocks (ragged and equal), histograms, barrier."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hkcsa import dist_sa
        full = O.gen_text(O.ENG96, 3, 10_000 + world)
        cuts_ragged = [0] + [1000 * (r + 1) + r for r in range(world - 1)] + [len(full)]
        cuts_equal = [len(full) // world * r for r in range(world)] + [len(full) // world * world]

        def program():
            got = {}
            for name, cuts in (("ragged", cuts_ragged), ("equal", cuts_equal)):
                block = torch.from_numpy(full[cuts[rank]:cuts[rank + 1]].copy())
                sizes = (yield ("gather", np.array([block.numel()], dtype=np.int64)))[:, 0]
                text, text_start, text_ready = yield ("text_async", block, sizes)
                text_start()
                text_ready()
                got[name] = (sizes.tolist(), text.numpy().copy())
            hist = np.bincount(full[cuts_ragged[rank]:cuts_ragged[rank + 1]], minlength=256).astype(np.int64)
            got["hist"] = (yield ("gather", torch.from_numpy(hist))).sum(0)
            yield ("barrier",)
            return got

        got = dist_sa._torch_run(program(), None, torch.device("cpu"))
        ok = got["ragged"][0] == list(np.diff(cuts_ragged)) and np.array_equal(got["ragged"][1], full)
        ok = ok and np.array_equal(got["equal"][1], full[: cuts_equal[-1]])
        ok = ok and np.array_equal(got["hist"], np.bincount(full, minlength=256))
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rank_program_collectives_gloo(world):
    out = mp.get_context("spawn").Manager().dict()
    mp.spawn(_program_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {r: 1 for r in range(world)}
