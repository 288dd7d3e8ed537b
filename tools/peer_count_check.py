"""torchrun check (N GPUs) of the count kernel with the result gather fused in (hkcsa_count_batch_peers):
identical to count + NCCL all-gather, and the time of both.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_count_check.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch, torch.distributed as dist
from hkcsa import engine as E, dist as hdist

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("N", 200_000_000)); P = int(os.environ.get("P", 10_000_000)); kind = int(os.environ.get("KIND", 0))
text = torch.cat([E.gen_text(kind, 42 + kind, n), torch.tensor([0x24], dtype=torch.uint8, device=dev)])
idx = E.DeviceIndex(text, sa_sample_rate=32) if rank == 0 else None
idx = hdist.broadcast_index(idx, src=0, device=dev, with_bwt=True)
alpha = torch.from_numpy(np.frombuffer(idx.wt.alphabet, dtype=np.uint8).copy()).to(dev); alpha = alpha[alpha != 0x24]
pats, off = E.gen_patterns(44, P, text[:n], alpha)
idx.build_kmer_table(); idx.build_occ_table(5, layout=1)
bounds = hdist.shard_bounds(off, world)
b, e = bounds[rank]
lp, lo_ = hdist.local_slice(pats, off, b, e)
out = hdist.PeerRanges(P, dev)

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize(); dist.barrier()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item()) * 1e3, r

def plain():
    lo, hi = idx.count_batch(lp, lo_, use_kmer_table=True)
    return hdist.gather_ranges(lo, hi, bounds)

def search_only():
    return idx.count_batch(lp, lo_, use_kmer_table=True)

ms_search, _ = timed(search_only)
ms_plain, (glo, ghi) = timed(plain)
ms_fused, (flo, fhi) = timed(lambda: hdist.sharded_count_fused(idx, pats, off, out, bounds=bounds, use_kmer_table=True))
ok = bool(torch.equal(glo, flo)) and bool(torch.equal(ghi, fhi))
okt = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"check": "count_with_fused_gather", "world": world, "text_bytes": n, "patterns": P,
                      "identical_on_every_rank": bool(okt.item()), "search_only_ms": ms_search,
                      "search_plus_nccl_allgather_ms": ms_plain, "search_with_fused_peer_stores_ms": ms_fused,
                      "patterns_per_s_fused_all_results_on_all_ranks": P / (ms_fused / 1e3),
                      "patterns_per_s_nccl": P / (ms_plain / 1e3)}))
dist.destroy_process_group()
