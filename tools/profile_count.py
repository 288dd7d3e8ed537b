"""Builds the 200 MB English-like index (BASELINE configs[2]) and runs the batched count once per rank structure
(wavelet tree, sampled Occ table) -- the target of `ncu -k regex:fm_count` captures (profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np
import torch
from hkcsa import engine as E

n = int(os.environ.get("N", 200_000_000))
P = int(os.environ.get("P", 2_000_000))
kind = int(os.environ.get("KIND", 0))
text = torch.cat([E.gen_text(kind, 42 + kind, n), torch.tensor([0x24], dtype=torch.uint8, device="cuda")])
idx = E.DeviceIndex(text, sa_sample_rate=32)
alpha = torch.from_numpy(np.frombuffer(idx.wt.alphabet, dtype=np.uint8).copy()).cuda()
alpha = alpha[alpha != 0x24]
pats, off = E.gen_patterns(44, P, text[:n], alpha)
idx.build_kmer_table()
a = idx.count_batch(pats, off, use_kmer_table=True, use_occ_table=False)
idx.build_occ_table(int(os.environ.get("SHIFT", 5)), layout=int(os.environ.get("LAYOUT", 1)))
b = idx.count_batch(pats, off, use_kmer_table=True, use_occ_table=True)
torch.cuda.synchronize()
assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
print("ok", n, P)
