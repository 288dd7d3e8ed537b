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


def shard_bounds(offsets: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Split P patterns (CSR offsets int64[P+1]) into `world` contiguous slices with near-equal
    total length (the work of a backward search is one step per symbol)."""
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
    bounds = shard_bounds(off.cpu().numpy(), world)
    b, e = bounds[rank]
    lp, lo_ = local_slice(pat, off, b, e)
    lo, hi = count_fn(lp, lo_)
    P = off.numel() - 1
    width = max(x[1] - x[0] for x in bounds) if bounds else 0
    send = torch.full((2, width), -2, dtype=torch.int64, device=lo.device)
    send[0, : e - b] = lo
    send[1, : e - b] = hi
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    out_lo = torch.empty(P, dtype=torch.int64, device=lo.device)
    out_hi = torch.empty(P, dtype=torch.int64, device=lo.device)
    for r, (rb, re_) in enumerate(bounds):
        out_lo[rb:re_] = recv[r][0, : re_ - rb]
        out_hi[rb:re_] = recv[r][1, : re_ - rb]
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
