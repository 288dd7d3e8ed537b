"""Timeline of one index build (GPU only): every bracketed launch with its start / end, and the idle time before it."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch
from hkcsa import engine as E, _lib

kind, seed, n = (1, 43, 100_000_000) if len(sys.argv) < 2 or sys.argv[1] == "c2" else (0, 42, 200_000_000)
L = _lib.load()
text = E.gen_text(kind, seed, n)
text = torch.cat([text, torch.tensor([0x24], dtype=torch.uint8, device=text.device)])
for _ in range(3):
    E.DeviceIndex(text, sa_sample_rate=32)
torch.cuda.synchronize()
names = [k for k in E.prof_read().keys()] if False else None
E.prof_enable(True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
idx = E.DeviceIndex(text, sa_sample_rate=32)
b.record()
torch.cuda.synchronize()
N = 512
st, en, cl = (C.c_float * N)(), (C.c_float * N)(), (C.c_int * N)()
cnt = C.c_int(0)
_lib.check(L.hkcsa_prof_timeline(st, en, cl, N, C.byref(cnt)))
cls_names = {}
for nm in ["byte_hist", "sa_pack0", "sa_keybuild", "radix_scan", "onesweep_u64", "seg_reduce", "seg_scan", "seg_apply",
           "bwt_gather", "wt_levels", "wt_pack", "wt_dir", "count", "locate", "ssa_build", "other"]:
    cls_names[L.hkcsa_prof_class_index(nm.encode())] = nm
print(f"step {a.elapsed_time(b):.3f} ms (with every launch bracketed)")
prev_end, busy, gaps = 0.0, 0.0, 0.0
for i in range(cnt.value):
    gap = st[i] - prev_end
    print(f"{i:3d} {cls_names.get(cl[i], '?'):14s} start {st[i]:8.3f} end {en[i]:8.3f} dur {en[i] - st[i]:7.3f} gap-before {gap:7.3f}")
    if gap > 0:
        gaps += gap
    prev_end = max(prev_end, en[i])
print(f"sum of gaps between bracketed scopes on the critical order: {gaps:.3f} ms")
E.prof_enable(False)
