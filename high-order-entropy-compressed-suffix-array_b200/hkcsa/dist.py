"""Multi-GPU query path (BASELINE config 4): one process per GPU,
the index replicated (built once and broadcast over NCCL/NVLink, or rebuilt per rank), the
pattern batch split into contiguous CSR slices balanced by total symbols, results gathered.
There is no data-path collective inside the search itself -- patterns are independent.

torch.distributed is the plumbing (NCCL on the GPU box, gloo in the CPU tests); the search on
each rank is libhkcsa's count kernel.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(offsets, world: int) -> list[tuple[int, int]]:
    """Split P patterns (CSR offsets int64[P+1]) into `world` contiguous slices with near-equal
    total length (the work of a backward search is one step per symbol).  A torch tensor (any device) is
    searched where it lives -- only the `world` cut points come back to the host."""
    if isinstance(offsets, torch.Tensor):
        P = offsets.numel() - 1
        if P <= 0:
            return [(0, max(P, 0))] * world
        first, last = int(offsets[0].item()), int(offsets[-1].item())
        total = last - first
        if total == 0:
            cuts = [min(P, (P * r) // world) for r in range(world)] + [P]
        else:
            targets = torch.tensor([first + (total * r) // world for r in range(1, world)], dtype=offsets.dtype,
                                   device=offsets.device)
            mid = torch.searchsorted(offsets, targets, right=False).tolist() if world > 1 else []
            cuts = [0] + [int(x) for x in mid] + [P]
        for i in range(1, len(cuts)):
            cuts[i] = max(cuts[i], cuts[i - 1])
        return [(cuts[r], cuts[r + 1]) for r in range(world)]
    offsets = np.asarray(offsets, dtype=np.int64)
    P = len(offsets) - 1
    total = int(offsets[-1] - offsets[0]) if P > 0 else 0
    cuts = [0]
    for r in range(1, world):
        if total == 0:
            cuts.append(min(P, (P * r) // world))
        else:
            target = offsets[0] + (total * r) // world
            cuts.append(int(np.searchsorted(offsets, target, side="left")))
    cuts.append(P)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def local_slice(pat: torch.Tensor, off: torch.Tensor, begin: int, end: int):
    """The CSR slice [begin, end) re-based to start at 0."""
    o = off[begin:end + 1]
    base = int(o[0].item()) if o.numel() else 0
    stop = int(o[-1].item()) if o.numel() else 0
    return pat[base:stop], (o - base).contiguous()


def sharded_count(count_fn, pat: torch.Tensor, off: torch.Tensor, group=None):
    """Every rank holds the full batch; rank r searches slice r with `count_fn(pat, off) ->
    (lo, hi)` and the (lo, hi) pairs are all-gathered so every rank ends with all P answers."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    bounds = shard_bounds(off, world)
    b, e = bounds[rank]
    lp, lo_ = local_slice(pat, off, b, e)
    lo, hi = count_fn(lp, lo_)
    return gather_ranges(lo, hi, bounds, group=group)


def gather_ranges(lo: torch.Tensor, hi: torch.Tensor, bounds, group=None):
    """All-gather of the per-rank (lo, hi) slices into the full batch order: one collective on a padded
    (2, width) block per rank, then `world` slice copies."""
    world = dist.get_world_size(group)
    P = bounds[-1][1] if bounds else 0
    width = max(x[1] - x[0] for x in bounds) if bounds else 0
    k = lo.numel()
    send = torch.empty((2, width), dtype=torch.int64, device=lo.device)
    send[0, :k] = lo
    send[1, :k] = hi
    recv = torch.empty((world, 2, width), dtype=torch.int64, device=lo.device)
    if hasattr(dist, "all_gather_into_tensor") and lo.is_cuda:
        dist.all_gather_into_tensor(recv, send, group=group)
    else:       # gloo (CPU tests)
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        recv = torch.stack(parts)
    out_lo = torch.empty(P, dtype=torch.int64, device=lo.device)
    out_hi = torch.empty(P, dtype=torch.int64, device=lo.device)
    for r, (rb, re_) in enumerate(bounds):
        out_lo[rb:re_] = recv[r, 0, : re_ - rb]
        out_hi[rb:re_] = recv[r, 1, : re_ - rb]
    return out_lo, out_hi


def broadcast_index(index, src: int = 0, group=None, device=None, with_bwt: bool = False):
    """Replicate a built DeviceIndex from `src` to every rank: the plan structs travel as bytes,
    the wavelet-tree blob (and sampled-SA blob) as one tensor each -- < 0.5 GB for a 200 MB text.
    with_bwt: also replicate the BWT (n bytes) so every rank can build its own sampled Occ table
    (DeviceIndex.build_occ_table: 1-2 ms locally, against broadcasting a table of 14 bytes per symbol)."""
    from . import engine
    from ._lib import SsaPlan, WtPlan

    rank = dist.get_rank(group)
    have = index is not None and rank == src
    meta = torch.zeros(4, dtype=torch.int64, device=device)
    if have:
        meta[0] = index.n
        meta[1] = index.wt.blob.numel()
        meta[2] = index.ssa.blob.numel() if index.ssa is not None else 0
        meta[3] = index.ssa.plan.rate if index.ssa is not None else 0
    dist.broadcast(meta, src, group=group)
    n, wt_bytes, ssa_bytes, rate = (int(x) for x in meta.tolist())

    def bcast_struct(obj, cls):
        buf = torch.zeros(C.sizeof(cls), dtype=torch.uint8, device=device)
        if have:
            buf.copy_(torch.frombuffer(bytearray(bytes(obj)), dtype=torch.uint8))
        dist.broadcast(buf, src, group=group)
        out = cls()
        C.memmove(C.byref(out), bytes(buf.cpu().numpy().tobytes()), C.sizeof(cls))
        return out

    plan = bcast_struct(index.wt.plan if have else None, WtPlan)
    blob = index.wt.blob if have else torch.empty(wt_bytes, dtype=torch.uint8, device=device)
    dist.broadcast(blob, src, group=group)
    ssa = None
    if ssa_bytes:
        splan = bcast_struct(index.ssa.plan if have else None, SsaPlan)
        sblob = index.ssa.blob if have else torch.empty(ssa_bytes, dtype=torch.uint8, device=device)
        dist.broadcast(sblob, src, group=group)
        ssa = engine.SampledSA(splan, sblob)
    bwt = None
    if with_bwt:
        bwt = index.bwt if have else torch.empty(n, dtype=torch.uint8, device=device)
        dist.broadcast(bwt, src, group=group)
    if have:
        return index
    replica = engine.DeviceIndex.from_parts(n, plan, blob, ssa)
    replica.bwt = bwt
    return replica


class PeerRanges:
    """The (lo, hi) answer arrays of a GLOBAL pattern batch, allocated in torch symmetric memory (peer-mapped over
    NVLink): every rank's count kernel writes its slice into the arrays of all ranks, so the all-gather of the
    results is part of the search kernel (hkcsa_count_batch_peers) and no collective follows it."""

    def __init__(self, P: int, device, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.P = int(P)
        self.buf = symm.empty((2, max(self.P, 1)), dtype=torch.int64, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.world = self.hdl.world_size
        self.rank = self.hdl.rank
        base = self.buf.data_ptr() - int(self.hdl.buffer_ptrs[self.rank])     # offset of the tensor in its allocation
        ptrs = [int(p) + base for p in self.hdl.buffer_ptrs]
        self.peer_lo = (C.c_uint64 * self.world)(*ptrs)
        self.peer_hi = (C.c_uint64 * self.world)(*[p + 8 * max(self.P, 1) for p in ptrs])

    @property
    def lo(self) -> torch.Tensor:
        return self.buf[0, : self.P]

    @property
    def hi(self) -> torch.Tensor:
        return self.buf[1, : self.P]

    def barrier(self) -> None:
        """Device-side barrier over the symmetric-memory signal pads, on the current stream: after it every
        rank's writes into this rank's arrays have landed."""
        self.hdl.barrier()


def sharded_count_fused(index, pat: torch.Tensor, off: torch.Tensor, out: PeerRanges, bounds=None,
                        use_kmer_table=None):
    """Every rank holds the full batch and searches slice `rank`; the kernel stores the answers into `out` on
    every rank.  Returns (lo, hi) of the whole batch (views of `out`): consume or clone them before the next call
    with the same `out` -- the call starts with a barrier, after which peers overwrite the arrays."""
    bounds = bounds if bounds is not None else shard_bounds(off, out.world)
    b, e = bounds[out.rank]
    lp, lo_ = local_slice(pat, off, b, e)
    out.barrier()                      # every rank is done with the previous batch's answers in `out`
    index.count_batch_peers(lp, lo_, b, out.peer_lo, out.peer_hi, use_kmer_table=use_kmer_table)
    out.barrier()
    return out.lo, out.hi


class PeerGather:
    """Packed answers {lo: low 32 bits, count: high 32 bits} of a GLOBAL pattern batch in torch symmetric memory.
    Every rank searches its slice into local (lo, hi) arrays and a store kernel (hkcsa_ranges_push_peers) writes the
    packed slice into the array of every rank -- 8 bytes per pattern over NVLink, one multimem.st per 16 bytes when
    the switch offers a multicast mapping.  The gather is pipelined against the search in chunks."""

    def __init__(self, P: int, device, group=None, use_multicast: bool | None = None):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.P = int(P)
        self.buf = symm.empty(max(self.P, 2), dtype=torch.int64, device=device)
        self.hdl = symm.rendezvous(self.buf, group)
        self.world = self.hdl.world_size
        self.rank = self.hdl.rank
        base = self.buf.data_ptr() - int(self.hdl.buffer_ptrs[self.rank])        # offset of the tensor in its allocation
        self.peer_out = (C.c_uint64 * self.world)(*[int(p) + base for p in self.hdl.buffer_ptrs])
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self.multicast = (mc + base) if (mc and use_multicast is not False) else 0
        if use_multicast and not self.multicast:
            raise RuntimeError("no multicast mapping for the symmetric buffer on this system")
        self._side = torch.cuda.Stream(device=device)

    @property
    def packed(self) -> torch.Tensor:
        return self.buf[: self.P]

    def unpack(self):
        """(lo, hi) int64 of the whole batch, the format count_batch returns."""
        from . import _lib
        lo = torch.empty(self.P, dtype=torch.int64, device=self.buf.device)
        hi = torch.empty_like(lo)
        _lib.check(_lib.load().hkcsa_ranges_unpack(self.buf.data_ptr(), self.P, lo.data_ptr(), hi.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream))
        return lo, hi


def chunked_slices(pat: torch.Tensor, off: torch.Tensor, begin: int, end: int, chunks: int):
    """[begin, end) of a CSR batch cut into `chunks` consecutive sub-batches: [(pat, off, first pattern)], computed
    once per batch layout (the cut points come to the host)."""
    out = []
    P = end - begin
    for c in range(chunks):
        b, e = begin + P * c // chunks, begin + P * (c + 1) // chunks
        if e > b:
            p_, o_ = local_slice(pat, off, b, e)
            out.append((p_, o_, b))
    return out


def sharded_count_packed(index, slices, out: PeerGather, **count_kw) -> torch.Tensor:
    """`slices`: this rank's part of the batch as chunked_slices() cut it.  Chunk k is searched on the current stream
    while chunk k-1 is pushed to the peers on a side stream.  Returns the packed answers of the whole batch (a view
    of `out`: consume before the next call with the same `out`)."""
    from . import _lib
    L = _lib.load()
    main = torch.cuda.current_stream()
    side = out._side
    out.hdl.barrier()                  # every rank is done with the previous batch's answers in `out`
    side.wait_stream(main)
    for p_, o_, base in slices:
        lo, hi = index.count_batch(p_, o_, **count_kw)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _lib.check(L.hkcsa_ranges_push_peers(lo.data_ptr(), hi.data_ptr(), lo.numel(), int(base), out.world,
                                                 out.peer_out, out.multicast, side.cuda_stream))
        lo.record_stream(side)
        hi.record_stream(side)
    main.wait_stream(side)
    out.hdl.barrier()                  # every rank's stores have landed here
    return out.packed
