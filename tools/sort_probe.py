"""Probe: per-pass time of the onesweep radix sort for different digit distributions (GPU only).
Answers: how much of a pass is the warp match (cost grows with the number of distinct digits per warp)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch
from hkcsa import engine as E

n = int(float(os.environ.get("N", "1e8")))
g = torch.Generator(device="cuda"); g.manual_seed(1)
cases = {
    "random 64-bit": torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g),
    "all zero": torch.zeros(n, dtype=torch.int64, device="cuda"),
    "4 values per digit": (torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g) & 0x0303030303030303),
    "16 values per digit": (torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g) & 0x0F0F0F0F0F0F0F0F),
}
only = os.environ.get("CASES")
if only:
    cases = {k: v for k, v in cases.items() if any(w in k for w in only.split(","))}
for name, keys in cases.items():
    vals = torch.arange(n, dtype=torch.int32, device="cuda")
    for rep in range(2):
        k, v = keys.clone(), vals.clone()
        E.prof_enable(True)
        torch.cuda.synchronize()
        E.sort_pairs_u64(k, v, 64)
        torch.cuda.synchronize()
        p = E.prof_read()
        E.prof_enable(False)
    o = p["onesweep_u64"]
    print(f"{name:22s} passes {o['launches']} avg {o['ms']/o['launches']:.3f} ms/pass  {o['alg_bytes']/o['ms']/1e6:.0f} GB/s", flush=True)
    if name == "random 64-bit":
        assert bool((k[1:].view(torch.int64) != k[:-1]).any())
