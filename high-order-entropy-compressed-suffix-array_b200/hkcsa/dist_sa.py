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
                stats: SaStats | None = None, wide: bool | None = None) -> torch.Tensor:
    """Sorted suffix ids of the suffixes in the bucket range: uint32 bit patterns in an int32 tensor, or -- for
    texts beyond 2^32-2 symbols, or when `wide` is forced -- an int64 tensor."""
    L = _lib.load()
    if wide is None:
        wide = text.numel() > _lib.DIST_MAX_N
    if wide:
        out = _empty(max(capacity, 1), torch.int64, text.device)
        nbytes = L.hkcsa_sa_subset64_scratch_bytes(capacity)
        scratch = _scratch(nbytes, text.device)
        count = C.c_uint64(0)
        bh = np.ascontiguousarray(byte_hist_np, dtype=np.uint64)
        st = stats if stats is not None else SaStats()
        check(L.hkcsa_sa_build_subset64(_ptr(text), text.numel(), bh.ctypes.data_as(C.POINTER(C.c_uint64)), bucket_lo,
                                        bucket_hi, _ptr(out), capacity, C.byref(count), _ptr(scratch), nbytes,
                                        _stream(), C.byref(st)))
        return out[: count.value]
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
    fn = _lib.load().hkcsa_bwt_slice64 if sa_slice.dtype == torch.int64 else _lib.load().hkcsa_bwt_slice
    check(fn(_ptr(text), text.numel(), _ptr(sa_slice), sa_slice.numel(), _ptr(out), _stream()))
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
        return self.sa if self.sa.dtype == torch.int64 else self.sa.to(torch.int64) & 0xFFFFFFFF


def distributed_suffix_array(local_block: torch.Tensor, group=None, wide: bool | None = None) -> SuffixArraySlice:
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
    if n > (1 << 40):
        raise _lib.HkcsaError(_lib.ERANGE, f"text of {n} symbols exceeds 2^40")
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
    sa = build_slice(text, bh_np, lo, hi, counts[rank], st, wide=wide)
    assert sa.numel() == counts[rank]
    # 4. BWT of the slice
    return SuffixArraySlice(rank, world, n, int(sum(counts[:rank])), sa, bwt_slice(text, sa), text, (lo, hi), st)


# ------------------------------------------------------------------ queries over a sliced index
class MultiSliceIndex:
    """FM index over a BWT that exists as slices (one per rank of the distributed build).  Every slice has its
    own wavelet tree and sampled SA; occ(c, i) = occurrences in earlier slices + a rank walk inside the slice
    holding row i (csrc/multi_slice.cu).  All slices are resident on this GPU (they are all-gathered after the
    build), so count / locate run locally and the pattern batch can be sharded exactly as for one index."""

    def __init__(self, n: int, slices, sa_sample_rate: int = 0):
        """slices: list of dicts with 'bwt' (uint8 tensor) and, for locate, 'sa' (uint32-as-int32 tensor), in
        rank order.  Builds the per-slice structures with libhkcsa K3 / the sampled-SA kernels."""
        from .engine import DeviceWaveletTree
        self.n = n
        self.device = slices[0]["bwt"].device
        self.wts, self.ssas = [], []
        starts = [0]
        for sl in slices:
            bw = sl["bwt"]
            self.wts.append(DeviceWaveletTree(bw))
            starts.append(starts[-1] + bw.numel())
            if sa_sample_rate > 0:
                self.ssas.append(_sampled_sa_slice(sl["sa"], sa_sample_rate))
        if starts[-1] != n:
            raise ValueError("slice lengths do not add up to n")
        self.starts = starts
        self.rate = sa_sample_rate
        self._assemble()

    @classmethod
    def from_parts(cls, n, starts, wt_parts, ssa_parts, rate):
        """wt_parts: [(WtPlan, blob tensor)], ssa_parts: [(SsaPlan, blob tensor)] or [] -- replicas."""
        from .engine import DeviceWaveletTree, SampledSA
        self = cls.__new__(cls)
        self.n, self.starts, self.rate = n, list(starts), rate
        self.device = wt_parts[0][1].device
        self.wts = []
        for plan, blob in wt_parts:
            wt = DeviceWaveletTree.__new__(DeviceWaveletTree)
            wt.device, wt.n, wt.hist, wt.plan, wt.blob = blob.device, int(plan.n), None, plan, blob
            self.wts.append(wt)
        self.ssas = [SampledSA(p, b) for p, b in ssa_parts]
        self._assemble()
        return self

    def _assemble(self):
        L = _lib.load()
        S = len(self.wts)
        if not 1 <= S <= _lib.MAX_SLICES:
            raise ValueError(f"1..{_lib.MAX_SLICES} slices")
        self.desc = torch.zeros(L.hkcsa_multi_desc_bytes(), dtype=torch.uint8, device=self.device)
        blobs = (C.c_void_p * S)(*[w.blob.data_ptr() for w in self.wts])
        plans = (C.POINTER(_lib.WtPlan) * S)(*[C.pointer(w.plan) for w in self.wts])
        starts = (C.c_uint64 * (S + 1))(*self.starts)
        if self.ssas:
            sblobs = (C.c_void_p * S)(*[s.blob.data_ptr() for s in self.ssas])
            splans = (C.POINTER(_lib.SsaPlan) * S)(*[C.pointer(s.plan) for s in self.ssas])
        else:
            sblobs, splans = None, None
        check(L.hkcsa_multi_desc_build(S, blobs, plans, starts, sblobs, splans, _ptr(self.desc), _stream()))

    def count_batch(self, pat: torch.Tensor, off: torch.Tensor):
        P = off.numel() - 1
        lo = _empty(P, torch.int64, self.device)
        hi = _empty(P, torch.int64, self.device)
        check(_lib.load().hkcsa_multi_count_batch(_ptr(self.desc), _ptr(pat), _ptr(off), P, _ptr(lo), _ptr(hi), _stream()))
        return lo, hi

    def locate_batch(self, pat: torch.Tensor, off: torch.Tensor):
        """CSR (offsets int64[P+1], positions int64 in SA order) through LF walks across slices."""
        if not self.ssas:
            raise ValueError("index was built without a sampled suffix array")
        L = _lib.load()
        lo, hi = self.count_batch(pat, off)
        P = lo.numel()
        cnt = torch.where(lo >= 0, hi - lo + 1, torch.zeros_like(lo))
        out_off = torch.zeros(P + 1, dtype=torch.int64, device=self.device)
        out_off[1:] = torch.cumsum(cnt, 0)
        total = int(out_off[-1].item()) if P else 0
        rows = _empty(total, torch.int64, self.device)
        check(L.hkcsa_expand_ranges64(_ptr(lo), _ptr(hi), _ptr(out_off), P, _ptr(rows), _stream()))
        out = _empty(total, torch.int64, self.device)
        check(L.hkcsa_multi_locate_rows(_ptr(self.desc), _ptr(rows), total, _ptr(out), _stream()))
        return out_off, out


def _sampled_sa_slice(sa_slice: torch.Tensor, rate: int):
    """Marks + samples for the rows of one slice (suffix ids are global, uint32 bit patterns)."""
    from .engine import SampledSA
    L = _lib.load()
    m = sa_slice.numel()
    wide = sa_slice.dtype == torch.int64
    ids = sa_slice if wide else (sa_slice.to(torch.int64) & 0xFFFFFFFF)
    n_marks = int((ids % rate == 0).sum().item())            # plumbing: sizes the sample array
    plan = _lib.SsaPlan()
    check(L.hkcsa_ssa_plan_make_slice(m, rate, n_marks, C.byref(plan)))
    blob = torch.zeros(int(plan.blob_bytes), dtype=torch.uint8, device=sa_slice.device)
    scratch = _scratch(plan.scratch_bytes, sa_slice.device)
    fn = L.hkcsa_ssa_build64 if wide else L.hkcsa_ssa_build
    check(fn(_ptr(sa_slice), C.byref(plan), _ptr(blob), _ptr(scratch), int(plan.scratch_bytes), _stream()))
    return SampledSA(plan, blob)


def replicate_sliced_index(sl: SuffixArraySlice, sa_sample_rate: int = 32, group=None) -> MultiSliceIndex:
    """After distributed_suffix_array: every rank builds the wavelet tree and sampled SA of ITS slice, the blobs are
    exchanged (one broadcast per slice: sizes differ), and every rank assembles the same MultiSliceIndex."""
    import torch.distributed as dist
    from .engine import DeviceWaveletTree
    world, rank, dev = sl.world, sl.rank, sl.sa.device
    wt = DeviceWaveletTree(sl.bwt)
    ssa = _sampled_sa_slice(sl.sa, sa_sample_rate) if sa_sample_rate > 0 else None
    meta = torch.zeros(world, 3, dtype=torch.int64, device=dev)
    meta[rank, 0] = sl.sa.numel()
    meta[rank, 1] = wt.blob.numel()
    meta[rank, 2] = ssa.blob.numel() if ssa is not None else 0
    dist.all_reduce(meta, group=group)
    meta = meta.cpu().tolist()
    starts = [0]
    for r in range(world):
        starts.append(starts[-1] + meta[r][0])

    def bcast_bytes(buf_or_none, nbytes, src):
        t = buf_or_none if rank == src else torch.empty(nbytes, dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=dist.get_global_rank(group, src) if group else src, group=group)
        return t

    def struct_bytes(obj):
        return torch.frombuffer(bytearray(bytes(obj)), dtype=torch.uint8).to(dev)

    wt_parts, ssa_parts = [], []
    for r in range(world):
        pbytes = bcast_bytes(struct_bytes(wt.plan) if rank == r else None, C.sizeof(_lib.WtPlan), r)
        plan = _lib.WtPlan.from_buffer_copy(pbytes.cpu().numpy().tobytes())
        blob = bcast_bytes(wt.blob if rank == r else None, meta[r][1], r)
        wt_parts.append((plan, blob))
        if sa_sample_rate > 0:
            sbytes = bcast_bytes(struct_bytes(ssa.plan) if rank == r else None, C.sizeof(_lib.SsaPlan), r)
            splan = _lib.SsaPlan.from_buffer_copy(sbytes.cpu().numpy().tobytes())
            sblob = bcast_bytes(ssa.blob if rank == r else None, meta[r][2], r)
            ssa_parts.append((splan, sblob))
    return MultiSliceIndex.from_parts(sl.n, starts, wt_parts, ssa_parts, sa_sample_rate)
