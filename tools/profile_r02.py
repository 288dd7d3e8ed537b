"""Round-2 kernels in one short program, the target of `ncu -k regex:...` captures kept under profiles/:
  * the distributed build with 4 ranks emulated on this GPU (dsa_pack_kernel<0/1>, extension rounds) on 4 x 64 MB of
    English-like text, and on a repetitive text (rank doubling: dsa_keybuild_dbl / dsa_isa_publish);
  * H_k from the suffix array (hk_*), the class/offset code of every wavelet level (rrr_*), rank on the coded form;
  * the packed result gather (ranges_push_kernel) with three local "peers".
Peers are plain buffers of this process, so NVLink is not exercised here -- the instruction mix, occupancy and DRAM
traffic of the kernels are."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np
import torch
from hkcsa import _lib, dist_sa, engine as E

n = int(os.environ.get("N", 256_000_000))
parts = 4
text = torch.cat([E.gen_text(0, 42, n - 1), torch.tensor([0x24], dtype=torch.uint8, device="cuda")])
blocks = [text[n * r // parts: n * (r + 1) // parts].clone() for r in range(parts)]
slices = dist_sa.emulate_distributed_suffix_array(blocks)
print("emulated build:", [s.sa.numel() for s in slices], "rounds", slices[0].rounds, flush=True)
rep = E.gen_text(0, 7, 1 << 20).repeat(32)
rep[-1] = 0x24
rb = [rep[rep.numel() * r // parts: rep.numel() * (r + 1) // parts].clone() for r in range(parts)]
rs = dist_sa.emulate_distributed_suffix_array(rb)
print("repetitive build: doubling rounds", rs[0].dbl_rounds, flush=True)
del slices, rs, blocks, rb

m = 100_000_000
t2 = text[:m].clone()
t2[-1] = 0x24
idx = E.DeviceIndex(t2, sa_sample_rate=32)
for k in (1, 3, 8):
    print("H_%d = %.6f" % (k, idx.entropy(k)), flush=True)
sp = idx.space()
print("bits/symbol: query blob %.3f raw levels %.3f stored %.3f" % (sp["query_blob_bits_per_symbol"],
      sp["raw_level_bits_per_symbol"], sp["coded_level_bits_per_symbol"]), flush=True)
vec = idx.coded_levels()[0]
pos = torch.randint(0, m, (4_000_000,), device="cuda")
r1 = vec.rank(pos)
assert torch.equal(r1, idx.wt.bv_rank(0, pos))
alpha = torch.from_numpy(np.frombuffer(idx.wt.alphabet, dtype=np.uint8).copy()).cuda()
pats, off = E.gen_patterns(44, 4_000_000, t2[: m - 1], alpha[alpha != 0x24])
lo, hi = idx.count_batch(pats, off)
P = lo.numel()
bufs = [torch.zeros(P, dtype=torch.int64, device="cuda") for _ in range(3)]
peers = (C.c_uint64 * 3)(*[b.data_ptr() for b in bufs])
_lib.check(_lib.load().hkcsa_ranges_push_peers(lo.data_ptr(), hi.data_ptr(), P, 0, 3, peers, 0,
                                               torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok")
