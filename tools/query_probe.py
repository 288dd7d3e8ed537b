"""Count / locate / build timings of one index (N, KIND, P from the environment): wavelet tree vs sampled Occ table.
(cudaLimitMaxL2FetchGranularity = 32 / 64 / 128 was probed with this script: no measurable effect on B200.)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch
from hkcsa import engine as E, _lib
L = _lib.load()
n = int(os.environ.get("N", 200_000_000)); P = int(os.environ.get("P", 4_000_000)); kind = int(os.environ.get("KIND", 0))
text = torch.cat([E.gen_text(kind, 42 + kind, n), torch.tensor([0x24], dtype=torch.uint8, device="cuda")])
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
idx = E.DeviceIndex(text, sa_sample_rate=32)
alpha = torch.from_numpy(np.frombuffer(idx.wt.alphabet, dtype=np.uint8).copy()).cuda(); alpha = alpha[alpha != 0x24]
pats, off = E.gen_patterns(44, P, text[:n], alpha)
idx.build_kmer_table()
for g in (0,):
    idx._occ = None
    wt_ms = t(lambda: idx.count_batch(pats, off, use_kmer_table=True, use_occ_table=False))
    loc_ms = t(lambda: idx.locate_batch(pats, off[: P // 8 + 1], use_samples=True))
    idx.build_occ_table(5, layout=int(os.environ.get("LAYOUT", 0)))
    occ_ms = t(lambda: idx.count_batch(pats, off, use_kmer_table=True, use_occ_table=True))
    loc2_ms = t(lambda: idx.locate_batch(pats, off[: P // 8 + 1], use_samples=True))
    idx._occ = None
    build_ms = t(lambda: E.DeviceIndex(text, sa_sample_rate=32), reps=2)
    print(f"g={g}: count wt {P/wt_ms/1e6:.3f} G/s  occ {P/occ_ms/1e6:.3f} G/s  locate wt {loc_ms:.2f} ms occ {loc2_ms:.2f} ms  build {build_ms:.2f} ms")
