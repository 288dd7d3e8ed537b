"""torchrun check of the distributed suffix-array build (hkcsa.dist_sa) on real GPUs:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_sa_check.py [--size BYTES] [--kind 0|1] [--verify]
Every rank contributes a contiguous block of the synthetic text; the slices are gathered on rank 0 and, with
--verify, compared with the single-GPU builder.  Prints one JSON line with the timing (max over ranks)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch, torch.distributed as dist
from hkcsa import engine as E, dist_sa

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=50_000_000)
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--verify", action="store_true")
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
n = args.size
full = E.gen_text(args.kind, 42 + args.kind, n)                 # same bytes on every rank; keep only my block
lo, hi = n * rank // world, n * (rank + 1) // world
block = full[lo:hi].clone()
del full
torch.cuda.synchronize(); dist.barrier()
best = None
for _ in range(args.reps):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    sl = dist_sa.distributed_suffix_array(block)
    torch.cuda.synchronize(); dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    best = float(dt.item()) if best is None else min(best, float(dt.item()))
counts = torch.zeros(world, dtype=torch.int64, device=dev); counts[rank] = sl.sa.numel()
dist.all_reduce(counts)
ok = None
if args.verify:
    cmax = int(counts.max().item())
    pad = torch.zeros(cmax, dtype=torch.int32, device=dev); pad[: sl.sa.numel()] = sl.sa
    bpad = torch.zeros(cmax, dtype=torch.uint8, device=dev); bpad[: sl.bwt.numel()] = sl.bwt
    gs = [torch.empty_like(pad) for _ in range(world)]; gb = [torch.empty_like(bpad) for _ in range(world)]
    dist.all_gather(gs, pad); dist.all_gather(gb, bpad)
    if rank == 0:
        sa = torch.cat([gs[r][: int(counts[r])] for r in range(world)])
        bw = torch.cat([gb[r][: int(counts[r])] for r in range(world)])
        ref = E.suffix_array(sl.text)
        ok = bool(torch.equal(sa, ref)) and bool(torch.equal(bw, E.bwt(sl.text, ref)))
if rank == 0:
    print(json.dumps({"check": "distributed_suffix_array", "world": world, "text_bytes": n, "kind": args.kind,
                      "seconds": best, "MB_per_s": n / 1e6 / best, "slice_sizes": counts.cpu().tolist(),
                      "rounds": int(sl.stats.rounds), "verified_against_single_gpu": ok}), flush=True)
dist.destroy_process_group()
