"""Probe (GPU box): EnhancedFMIndex(str) on the C2 text for different numbers of staging threads (HKCSA_STAGE_THREADS)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch
from hkcsa import engine as E
from csa.enhanced_fm_index import EnhancedFMIndex
text = E.gen_text(E.DNA4, 43, 100_000_000).cpu().numpy().tobytes().decode("latin-1")
EnhancedFMIndex(text[:1 << 20])
for t in sys.argv[1:]:
    os.environ["HKCSA_STAGE_THREADS"] = t
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fm = EnhancedFMIndex(text)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        del fm
    print(f"threads {t}: min {1e3 * min(ts):.2f} ms  median {1e3 * sorted(ts)[2]:.2f} ms", flush=True)
