"""Probe (GPU only): the distributed build with its ranks emulated on ONE GPU, every launch class timed.
Answers: how much of the group-local round is the ordering kernel (class sa_keybuild) and how much the seg_* refinement?
usage: python tools/group_round_probe.py [world=2] [bytes_per_rank=1e9] [wide=0]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200"))
import numpy as np, torch
from hkcsa import engine as E, dist_sa

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
wide = bool(int(sys.argv[3])) if len(sys.argv) > 3 else None
text = E.gen_text(E.ENG96, 42, world * per)
text[-1] = 0x24
blocks = [text[r * per:(r + 1) * per] for r in range(world)]
for rep in range(2):
    E.prof_enable(rep == 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = dist_sa.emulate_distributed_suffix_array(blocks, wide=wide, profile=(rep == 1))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rep == 1:
        print(f"world {world} x {per} bytes, wide {wide}: {dt * 1e3:.1f} ms for all ranks in turn")
        print("rounds rank 0:", res[0].rounds)
        print("phases rank 0 (ms):", {k: round(v * 1e3, 2) for k, v in res[0].phases.items()})
        import ctypes as C
        from hkcsa import _lib
        L = _lib.load()
        N = 2048
        st_, en_, cl_ = (C.c_float * N)(), (C.c_float * N)(), (C.c_int * N)()
        cnt = C.c_int(0)
        _lib.check(L.hkcsa_prof_timeline(st_, en_, cl_, N, C.byref(cnt)))
        kb, sap = L.hkcsa_prof_class_index(b"sa_keybuild"), L.hkcsa_prof_class_index(b"seg_apply")
        print("sa_keybuild launches (ms):", [round(en_[i] - st_[i], 3) for i in range(cnt.value) if cl_[i] == kb])
        print("seg_apply launches (ms):", [round(en_[i] - st_[i], 3) for i in range(cnt.value) if cl_[i] == sap])
        p = E.prof_read()
        E.prof_enable(False)
        for k, v in sorted(p.items(), key=lambda kv: -kv[1]["ms"]):
            if v["launches"]:
                print(f"  {k:14s} {v['ms']:9.3f} ms  {v['launches']:4d} launches")
    del res
