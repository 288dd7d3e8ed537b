"""Distributed suffix-array build for texts beyond one GPU's working set (SURVEY.md section 8e,
BASELINE config 5).  One process per GPU.

    1. every rank holds a contiguous block of the text; the blocks are all-gathered over NCCL so that
       every GPU has the whole text (8 GB fit a 180 GB B200 many times over);
    2. byte histogram and the 65536-bucket histogram of the top 16 bits of the round-0 keys are computed
       on each rank's own block and all-reduced; every rank derives the same balanced bucket ranges;
    3. rank r sorts the suffixes whose key falls into range r (libhkcsa: select + pack, onesweep radix
       sort, extension rounds that read the replicated text) -- no exchange during the sort;
    4. the slices, concatenated in rank order, are the suffix array; the BWT slice is a local gather.

Suffix ids are 32-bit (n <= 2^32 - 2); a slice holds at most 2^30 - 2 suffixes.  The collectives are the
text all-gather and two small all-reduces; an all-to-all is not needed because the text is replicated.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import SaStats, check
from .engine import _empty, _ptr, _scratch, _stream, byte_hist


def balanced_bucket_ranges(bucket_hist: np.ndarray, parts: int) -> list[tuple[int, int]]:
    """Cut the bucket axis into `parts` contiguous ranges with near-equal suffix counts (host logic)."""
    h = np.asarray(bucket_hist, dtype=np.int64)
    nb = len(h)
    cum = np.concatenate([[0], np.cumsum(h)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, parts):
        target = (total * r) // parts
        c = int(np.searchsorted(cum, target, side="left"))
        cuts.append(min(nb, max(c, cuts[-1])))
    cuts.append(nb)
    return [(cuts[r], cuts[r + 1]) for r in range(parts)]


def key_bucket_hist(text: torch.Tensor, begin: int, end: int, byte_hist_np: np.ndarray) -> torch.Tensor:
    """Histogram (int64[65536], device) of the key buckets of suffixes [begin, end) of `text`."""
    hist = torch.zeros(_lib.DIST_BUCKETS, dtype=torch.int64, device=text.device)
    bh = np.ascontiguousarray(byte_hist_np, dtype=np.uint64)
    check(_lib.load().hkcsa_sa_key_hist(_ptr(text), text.numel(), begin, end, bh.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        _ptr(hist), _stream()))
    return hist


def build_slice(text: torch.Tensor, byte_hist_np: np.ndarray, bucket_lo: int, bucket_hi: int, capacity: int,
                stats: SaStats | None = None) -> torch.Tensor:
    """Sorted suffix ids (uint32 bit patterns in an int32 tensor) of the suffixes in the bucket range."""
    L = _lib.load()
    out = _empty(max(capacity, 1), torch.int32, text.device)
    nbytes = L.hkcsa_sa_subset_scratch_bytes(capacity)
    scratch = _scratch(nbytes, text.device)
    count = C.c_uint64(0)
    bh = np.ascontiguousarray(byte_hist_np, dtype=np.uint64)
    st = stats if stats is not None else SaStats()
    check(L.hkcsa_sa_build_subset(_ptr(text), text.numel(), bh.ctypes.data_as(C.POINTER(C.c_uint64)), bucket_lo,
                                  bucket_hi, _ptr(out), capacity, C.byref(count), _ptr(scratch), nbytes, _stream(),
                                  C.byref(st)))
    return out[: count.value]


def bwt_slice(text: torch.Tensor, sa_slice: torch.Tensor) -> torch.Tensor:
    out = _empty(sa_slice.numel(), torch.uint8, text.device)
    check(_lib.load().hkcsa_bwt_slice(_ptr(text), text.numel(), _ptr(sa_slice), sa_slice.numel(), _ptr(out), _stream()))
    return out


@dataclass
class SuffixArraySlice:
    rank: int
    world: int
    n: int                      # length of the whole text
    offset: int                 # global SA position of this slice's first entry
    sa: torch.Tensor            # uint32 suffix ids (int32 storage), sorted
    bwt: torch.Tensor           # uint8, same length
    text: torch.Tensor          # the replicated text
    bucket_range: tuple
    stats: SaStats

    def sa_int64(self) -> torch.Tensor:
        return self.sa.to(torch.int64) & 0xFFFFFFFF


def distributed_suffix_array(local_block: torch.Tensor, group=None) -> SuffixArraySlice:
    """local_block: this rank's contiguous part of the text (uint8, on this rank's GPU), blocks in rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_block.device
    # 1. replicate the text
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[rank] = local_block.numel()
    dist.all_reduce(sizes, group=group)
    sizes = sizes.cpu().tolist()
    n = int(sum(sizes))
    if n > _lib.DIST_MAX_N:
        raise _lib.HkcsaError(_lib.ERANGE, f"text of {n} symbols exceeds the 32-bit suffix-id limit")
    text = torch.empty(n, dtype=torch.uint8, device=dev)
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    text[starts[rank]:starts[rank + 1]] = local_block
    for r in range(world):                                   # all-gather of unequal blocks
        if sizes[r]:
            dist.broadcast(text[starts[r]:starts[r + 1]], src=dist.get_global_rank(group, r) if group else r, group=group)
    # 2. global byte histogram and key-bucket histogram
    bh = torch.from_numpy(byte_hist(local_block).astype(np.int64)).to(dev)
    dist.all_reduce(bh, group=group)
    bh_np = bh.cpu().numpy().astype(np.uint64)
    kh = key_bucket_hist(text, int(starts[rank]), int(starts[rank + 1]), bh_np)
    dist.all_reduce(kh, group=group)
    kh_np = kh.cpu().numpy()
    ranges = balanced_bucket_ranges(kh_np, world)
    counts = [int(kh_np[a:b].sum()) for a, b in ranges]
    # 3. this rank's slice
    lo, hi = ranges[rank]
    st = SaStats()
    sa = build_slice(text, bh_np, lo, hi, counts[rank], st)
    assert sa.numel() == counts[rank]
    # 4. BWT of the slice
    return SuffixArraySlice(rank, world, n, int(sum(counts[:rank])), sa, bwt_slice(text, sa), text, (lo, hi), st)
