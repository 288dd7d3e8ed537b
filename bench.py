#!/usr/bin/env python
"""bench.py -- index build MB/s (SA + BWT + wavelet tree) and batched count / locate
throughput of the B200 hot path, with the reference's CPU path timed beside it.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # reference arm (CPU oracle port)

One step = one full index build of the workload text: byte histogram -> K1 suffix
array -> K2 BWT -> K3 wavelet tree with rank/select directories and C[] -> sampled
SA.  Default workload = BASELINE.json configs[1]: 100 MB synthetic DNA (sigma 4,
order-5 Markov, seed 43) + '$'.  N > 1 (torchrun): N independent replicas, one per
GPU ("replicas only"), value = N * text bytes / max-over-ranks time;
the count queries that follow are sharded over the ranks (index replicated).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "high-order-entropy-compressed-suffix-array_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (generator kind, seed, text bytes, description)
    "c1": (0, 42, 1 << 20, "1 MiB ENG96 order-3 Markov text + '$' (BASELINE configs[0])"),
    "c2": (1, 43, 100_000_000, "100 MB DNA4 order-5 Markov text + '$' (BASELINE configs[1])"),
    "c3": (0, 42, 200_000_000, "200 MB ENG96 order-3 Markov text + '$' (BASELINE configs[2])"),
}
METRIC = "index build MB/s (SA+BWT+WT)"
TOP_KERNEL_CLASS = "onesweep_u64"     # the dominant kernel (share of the step checked in profiles/): timed live
SA_SAMPLE_RATE = 32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override the text size in bytes (debug)")
    ap.add_argument("--patterns", type=int, default=10_000_000, help="count queries after the build (C4 shape)")
    ap.add_argument("--cpu-sample", type=int, default=32_000_000, help="bytes of the workload the CPU baseline builds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-queries", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region (NVML, every 20 ms; falls back to
    nvidia-smi polling)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
        if out.returncode == 0 and out.stdout.strip():
            r = [x.strip() for x in out.stdout.strip().split(",")]
            self.sm.append(float(r[0]))
            self.sm_max = float(r[1])
            for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
                if r[2 + i].lower() == "active":
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# ------------------------------------------------------------------ reference arm / CPU baseline
def cpu_build_sample(kind, seed, nbytes):
    """The oracle port of the reference's build path on a bounded sample of the workload:
    build_suffix_array(text+'$') -> bwt_transform -> WaveletTree(bwt) spine + rank_support + C[]."""
    from oracle import oracle as O
    text = O.gen_text(kind, seed, nbytes).tobytes() + b"$"
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    sa = O.build_suffix_array(text, threads=threads)
    bwt = O.bwt_transform(text, sa)
    O.build_count(text)
    _, levels = O.wt_spine(bwt)
    for bits in levels:
        O.rank_support(bits)
    dt = time.perf_counter() - t0
    return dt, threads, len(text)


def run_reference(args):
    kind, seed, nbytes, desc = WORKLOADS[args.workload]
    if args.size:
        nbytes = args.size
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = min(nbytes, args.cpu_sample)
    times = []
    threads = 1
    for i in range(args.warmup + args.steps):
        dt, threads, n = cpu_build_sample(kind, seed, sample)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = sample / 1e6 / (ms / 1e3)
    sample_desc = f"first {sample} bytes of the workload text (+'$'), oracle port of SA+BWT+WT, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/u32/u64 integer", "data": "synthetic",
        "config": {"workload": desc, "text_bytes": nbytes, "cpu_sample_bytes": sample},
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from hkcsa import engine as E
    from hkcsa import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("HKCSA_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1 << 20, dtype=torch.uint8, device=dev)               # set up the communicator
        dist.broadcast(warm, 0)
        dist.all_gather([torch.empty_like(warm) for _ in range(world)], warm)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    kind, seed, nbytes, desc = WORKLOADS[args.workload]
    if args.size:
        nbytes = args.size
    L = _lib.load()

    # ---- synthetic text (+ '$'), resident in HBM; pinned host copy for the e2e leg
    text = torch.empty(nbytes + 1, dtype=torch.uint8, device=dev)
    E.check(L.hkcsa_gen_text(kind, seed, nbytes, text.data_ptr(), torch.cuda.current_stream().cuda_stream))
    text[nbytes] = 0x24
    n = nbytes + 1
    h_text = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_text.copy_(text)
    h_sa = torch.empty(n, dtype=torch.int32).pin_memory()
    h_bwt = torch.empty(n, dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    torch.cuda.synchronize()

    def build_step(src):
        return E.DeviceIndex(src, sa_sample_rate=SA_SAMPLE_RATE)

    launches0 = L.hkcsa_launch_count()
    # ---- warm-up
    idx = None
    for _ in range(args.warmup):
        idx = build_step(text)
    torch.cuda.synchronize()
    launches_per_step = (L.hkcsa_launch_count() - launches0) // max(1, args.warmup)

    # ---- timed: K device-resident builds, CUDA events per step, L2 flushed between steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    E.prof_enable(True, classes=[TOP_KERNEL_CLASS])    # live durations of the dominant kernel only (roofline)
    clk = ClockSampler(local)          # samples through both timed regions (device-resident and e2e)
    clk.__enter__()
    barrier()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.zero_()
        a.record()
        idx = build_step(text)
        b.record()
    barrier()
    wall = time.perf_counter() - wall0
    prof_top = E.prof_read()
    E.prof_enable(False)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms = max_over_ranks(float(np.mean(step_ms)))
    value = world * nbytes / 1e6 / (ms / 1e3)
    stats = idx.stats.sa

    # ---- e2e: host text -> H2D -> build -> D2H of SA + BWT, every step
    def e2e_step():
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        d.copy_(h_text, non_blocking=True)
        ix = E.DeviceIndex(d, sa_sample_rate=SA_SAMPLE_RATE, host_sa=h_sa, host_bwt=h_bwt)
        torch.cuda.synchronize()      # build + both device->host copies (side stream) done
        return ix

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
    e2e_value = world * nbytes / 1e6 / (e2e_ms / 1e3)
    clk.__exit__(None, None, None)

    # ---- per-kernel breakdown: two more builds with every kernel class timed (outside the timed region: the
    #      event pairs around ~130 launches per build are not free)
    E.prof_enable(True)
    for _ in range(2):
        flush.zero_()
        build_step(text)
    prof = E.prof_read()
    E.prof_enable(False)
    for v in prof.values():
        v["steps"] = 2
    prof[TOP_KERNEL_CLASS] = prof_top[TOP_KERNEL_CLASS]     # the roofline kernel: from the timed region itself
    prof[TOP_KERNEL_CLASS]["steps"] = args.steps

    # ---- dominant kernel roofline: onesweep radix pass, 24 B per (key, value) pair per launch
    peak, peak_src = measured_peak_gbs()
    top = max(prof.items(), key=lambda kv: kv[1]["ms"] / kv[1]["steps"]) if prof else (None, None)
    one = prof.get("onesweep_u64")
    roofline = None
    traffic = None
    try:   # DRAM bytes per pair from the committed ncu --set full capture, scaled to this run's average launch
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            tr = json.load(f)["onesweep64_kernel"]
        per_pair = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["pairs_in_profiled_launch"]
        if one and one["launches"]:
            traffic = per_pair * (one["alg_bytes"] / 24.0) / one["launches"]
    except Exception:
        traffic = None
    if one and one["ms"] > 0:
        achieved = one["alg_bytes"] / (one["ms"] / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "onesweep64_kernel (one 8-bit LSD radix pass, 24 B per (u64,u32) pair)",
                    "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic,
                    "alg_bytes_per_launch": one["alg_bytes"] / one["launches"],
                    "launches": one["launches"], "avg_launch_ms": one["ms"] / one["launches"],
                    "share_of_step": one["ms"] / max(1e-9, sum(step_ms))}

    # ---- batched count / locate (BASELINE config 4 shape).  N > 1: rank 0's index is broadcast over
    #      NCCL, the SAME global batch is cut into contiguous slices balanced by total symbols
    #      (hkcsa.dist.shard_bounds), every rank searches its slice, (lo, hi) are all-gathered.
    queries = None
    if not args.no_queries and args.patterns > 0:
        from hkcsa import dist as hdist
        P_total = args.patterns
        alpha = torch.from_numpy(np.frombuffer(idx.wt.alphabet, dtype=np.uint8).copy()).to(dev)
        alpha = alpha[alpha != 0x24]
        pats, off = E.gen_patterns(44, P_total, text[:nbytes], alpha)       # identical on every rank
        bcast_ms = None
        q_idx = idx
        if world > 1:
            barrier()
            t0 = time.perf_counter()
            q_idx = hdist.broadcast_index(idx if rank == 0 else None, src=0, device=dev, with_bwt=True)
            barrier()
            bcast_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            bounds = hdist.shard_bounds(off, world)
            pb, pe = bounds[rank]
            my_pats, my_off = hdist.local_slice(pats, off, pb, pe)
        else:
            my_pats, my_off = pats, off
        P = my_off.numel() - 1
        torch.cuda.synchronize()
        reps = 3

        def time_count(use_table, use_occ=False):
            for _ in range(2):
                q_idx.count_batch(my_pats, my_off, use_kmer_table=use_table, use_occ_table=use_occ)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                r_ = q_idx.count_batch(my_pats, my_off, use_kmer_table=use_table, use_occ_table=use_occ)
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b) / reps), r_

        c_ms_plain, (lo0, hi0) = time_count(False)          # every symbol walks the wavelet tree
        t0 = time.perf_counter()
        q_idx.build_kmer_table()
        torch.cuda.synchronize()
        kmer_ms = (time.perf_counter() - t0) * 1e3
        c_ms, (lo, hi) = time_count(True)                   # last k symbols from the k-mer jump table
        assert torch.equal(lo, lo0) and torch.equal(hi, hi0)
        occ_info = None
        if q_idx.bwt is not None:          # same queries ranked on the sampled Occ table (identical ranges)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            plan_o, _ = q_idx.build_occ_table(5, layout=1)
            torch.cuda.synchronize()
            occ_build_ms = (time.perf_counter() - t0) * 1e3
            o_ms, (lo_o, hi_o) = time_count(True, True)
            assert torch.equal(lo, lo_o) and torch.equal(hi, hi_o)
            occ_info = {"count_patterns_per_s": P_total / (o_ms / 1e3), "count_ms": o_ms, "build_ms": occ_build_ms,
                        "bytes": int(plan_o.blob_bytes), "layout": "per-symbol bitmaps, one 8-byte entry per 32 rows"}
            q_idx._occ = None
        gather_ms = None
        if world > 1:                      # results to every rank, timed apart from the search
            hdist.gather_ranges(lo, hi, bounds)                      # first call: NCCL channel set-up
            barrier()
            t0 = time.perf_counter()
            glo, ghi = hdist.gather_ranges(lo, hi, bounds)
            barrier()
            gather_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            hits = float((glo >= 0).sum().item())
        else:
            hits = float((lo >= 0).sum().item())
        # locate a slice of this rank's patterns through the sampled SA (LF walks)
        PL = min(P, 1_000_000 // world)
        offL = my_off[: PL + 1]
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o_off, o_pos = q_idx.locate_batch(my_pats, offL, use_samples=True)
        torch.cuda.synchronize()
        a2.record()
        o_off, o_pos = q_idx.locate_batch(my_pats, offL, use_samples=True)
        b2.record()
        barrier()
        l_ms = max_over_ranks(a2.elapsed_time(b2))
        occ_total = sum_over_ranks(float(o_pos.numel()))
        best_ms = min(c_ms, occ_info["count_ms"]) if occ_info else c_ms
        queries = {"count_patterns_per_s": P_total / (best_ms / 1e3), "count_patterns": P_total,
                   "count_ms": best_ms,
                   "rank_structure": ("sampled Occ table (per-symbol bitmaps)" if occ_info and occ_info["count_ms"] < c_ms
                                      else "wavelet tree") + " + k-mer jump table",
                   "count_patterns_per_s_wavelet_tree": P_total / (c_ms / 1e3),
                   "hit_fraction": hits / P_total, "pattern_len": "uniform 8-64",
                   "count_patterns_per_s_no_jump_table": P_total / (c_ms_plain / 1e3),
                   "occ_table": occ_info,
                   "kmer_jump_table": {"k": int(q_idx._kmer[1]), "build_ms": kmer_ms,
                                       "bytes": int(q_idx._kmer[0].numel() * 4) if q_idx._kmer[0] is not None else 0},
                   "index_broadcast_ms": bcast_ms, "result_allgather_ms": gather_ms,
                   "locate_occurrences_per_s": occ_total / (l_ms / 1e3), "locate_patterns": PL * world,
                   "locate_occurrences": occ_total, "locate_ms": l_ms, "sa_sample_rate": SA_SAMPLE_RATE,
                   "scaling": "strong (fixed global batch)",
                   "sharding": "index broadcast from rank 0, patterns split into contiguous slices balanced by symbols"}

    # ---- BASELINE config 4 proper: 10 M count queries on the 200 MB English-like index (C3 text), index built on
    #      rank 0 and broadcast, patterns sharded.  Reported beside the headline; skipped with --no-queries.
    c4 = None
    if not args.no_queries and args.patterns > 0 and args.workload == "c2" and not args.size:
        from hkcsa import dist as hdist
        k3, s3, n3, _ = WORKLOADS["c3"]
        del idx, q_idx
        torch.cuda.empty_cache()
        t3 = torch.empty(n3 + 1, dtype=torch.uint8, device=dev)
        E.check(L.hkcsa_gen_text(k3, s3, n3, t3.data_ptr(), torch.cuda.current_stream().cuda_stream))
        t3[n3] = 0x24
        idx3 = E.DeviceIndex(t3, sa_sample_rate=SA_SAMPLE_RATE) if rank == 0 or world == 1 else None
        if world > 1:
            idx3 = hdist.broadcast_index(idx3, src=0, device=dev, with_bwt=True)
        alpha3 = torch.from_numpy(np.frombuffer(idx3.wt.alphabet, dtype=np.uint8).copy()).to(dev)
        alpha3 = alpha3[alpha3 != 0x24]
        pats3, off3 = E.gen_patterns(44, args.patterns, t3[:n3], alpha3)
        pats3_full, off3_full = pats3, off3
        if world > 1:
            pb, pe = hdist.shard_bounds(off3, world)[rank]
            pats3, off3 = hdist.local_slice(pats3, off3, pb, pe)
        idx3.build_kmer_table()
        for _ in range(2):
            idx3.count_batch(pats3, off3, use_kmer_table=True, use_occ_table=False)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            lo3, hi3 = idx3.count_batch(pats3, off3, use_kmer_table=True, use_occ_table=False)
        b.record()
        barrier()
        c4_ms = max_over_ranks(a.elapsed_time(b) / 3)
        c4 = {"workload": "10 M count queries (len 8-64) on the 200 MB ENG96 index (BASELINE configs[3])",
              "count_patterns_per_s": args.patterns / (c4_ms / 1e3), "count_ms": c4_ms,
              "hit_fraction": sum_over_ranks(float((lo3 >= 0).sum().item())) / args.patterns,
              "rank_structure": "wavelet tree + k-mer jump table",
              "wavelet_levels": idx3.wt.levels, "kmer_k": int(idx3._kmer[1])}
        # locate of the first 1 M patterns of this rank through the sampled SA (LF walks)
        PL3 = min(off3.numel() - 1, 1_000_000 // world)

        def time_locate3():
            idx3.locate_batch(pats3, off3[: PL3 + 1], use_samples=True)
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            _, pos_ = idx3.locate_batch(pats3, off3[: PL3 + 1], use_samples=True)
            b_.record()
            barrier()
            return max_over_ranks(a_.elapsed_time(b_)), sum_over_ranks(float(pos_.numel()))

        l3_ms, l3_occ = time_locate3()
        c4["locate_occurrences_per_s"] = l3_occ / (l3_ms / 1e3)
        c4["locate_ms"] = l3_ms
        c4["locate_occurrences"] = l3_occ
        # the same queries ranked on the sampled Occ table (optional second rank structure, identical ranges)
        if idx3.bwt is not None:
            for shift, layout in ((5, 0), (6, 0), (5, 1)):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                plan_o, blob_o = idx3.build_occ_table(shift, layout=layout)
                torch.cuda.synchronize()
                occ_ms = (time.perf_counter() - t0) * 1e3
                for _ in range(2):
                    idx3.count_batch(pats3, off3, use_kmer_table=True, use_occ_table=True)
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3):
                    lo4, hi4 = idx3.count_batch(pats3, off3, use_kmer_table=True, use_occ_table=True)
                b.record()
                barrier()
                o_ms = max_over_ranks(a.elapsed_time(b) / 3)
                assert torch.equal(lo4, lo3) and torch.equal(hi4, hi3)
                l_ms, l_occ = time_locate3()
                peers = None
                if world > 1 and layout == 1:
                    # every rank needs every answer: NCCL all-gather after the search, against the search kernel
                    # storing its slice into all ranks' arrays itself (peer-mapped symmetric memory over NVLink)
                    bounds3 = hdist.shard_bounds(off3_full, world)
                    try:
                        peers = hdist.PeerRanges(args.patterns, dev)
                    except Exception as exc:       # symmetric memory unavailable on this box: report, do not fail
                        peers = None
                        c4["all_answers_on_all_ranks"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
                    ok_all = sum_over_ranks(1.0 if peers is not None else 0.0) == world
                    if not ok_all:
                        peers = None
                if peers is not None:

                    def timed3(fn):
                        fn()
                        torch.cuda.synchronize()
                        barrier()
                        t0_ = time.perf_counter()
                        for _ in range(3):
                            r_ = fn()
                        torch.cuda.synchronize()
                        barrier()
                        return max_over_ranks((time.perf_counter() - t0_) / 3 * 1e3), r_

                    def nccl3():
                        l_, h_ = idx3.count_batch(pats3, off3, use_kmer_table=True, use_occ_table=True)
                        return hdist.gather_ranges(l_, h_, bounds3)

                    g_ms, (glo3, ghi3) = timed3(nccl3)
                    f_ms, (flo3, fhi3) = timed3(lambda: hdist.sharded_count_fused(idx3, pats3_full, off3_full, peers,
                                                                                 bounds=bounds3, use_kmer_table=True))
                    assert torch.equal(glo3, flo3) and torch.equal(ghi3, fhi3)
                    c4["all_answers_on_all_ranks"] = {
                        "search_plus_nccl_allgather": {"ms": g_ms, "patterns_per_s": args.patterns / (g_ms / 1e3)},
                        "search_with_fused_peer_stores": {"ms": f_ms, "patterns_per_s": args.patterns / (f_ms / 1e3)},
                        "note": "wall clock incl. launch and the cross-rank barrier; hkcsa_count_batch_peers writes every "
                                "range into the result arrays of all ranks from inside the search kernel"}
                    del peers
                c4["occ_table_bitmaps" if layout else f"occ_table_rows_{1 << shift}"] = {"count_patterns_per_s": args.patterns / (o_ms / 1e3), "count_ms": o_ms,
                                                       "build_ms": occ_ms, "bytes": int(plan_o.blob_bytes),
                                                       "locate_occurrences_per_s": l_occ / (l_ms / 1e3), "locate_ms": l_ms}
                del blob_o
                idx3._occ = None
            c4["count_patterns_per_s_wavelet_tree"] = c4["count_patterns_per_s"]
            best = min(("occ_table_rows_32", "occ_table_rows_64", "occ_table_bitmaps"), key=lambda k_: c4[k_]["count_ms"])
            if c4[best]["count_ms"] < c4["count_ms"]:
                c4["count_patterns_per_s"] = c4[best]["count_patterns_per_s"]
                c4["count_ms"] = c4[best]["count_ms"]
                c4["rank_structure"] = "sampled Occ table (" + best + ") + k-mer jump table"

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on a bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = min(nbytes, args.cpu_sample)
        dt, threads, _ = cpu_build_sample(kind, seed, sample)
        cpu = {"value": sample / 1e6 / dt, "unit": "MB/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} bytes of the workload text (+'$'): oracle port of SA+BWT+WT spine, "
                         f"{dt:.2f} s on {threads} threads"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 symbols / u32 ranks / u64 keys (integer)", "data": "synthetic",
            "config": {"workload": desc, "text_bytes": nbytes, "index_symbols": n, "sa_sample_rate": SA_SAMPLE_RATE,
                       "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "l2": "256 MiB flush buffer written between timed steps; working set >> L2"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": "MB/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": n,
                    "d2h_bytes_per_step": 5 * n,
                    "what": "pinned host text -> H2D -> build -> D2H of SA (4n) + BWT (n) on a side stream "
                            "overlapping the rest of the build, wall clock incl. final sync"},
            "gpu_launches": int(launches_per_step) * args.steps,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "kernels": {k: {"ms_per_step": v["ms"] / v["steps"], "launches_per_step": v["launches"] / v["steps"],
                            "alg_GBps": (v["alg_bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 and v["alg_bytes"] else None}
                        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"] / kv[1]["steps"])},
            "kernels_note": TOP_KERNEL_CLASS + ": CUDA events inside the timed region; the other classes: two extra "
                            "builds after it with every launch bracketed",
            "top_kernel": top[0],
            "sa": {"rounds": int(stats.rounds), "k0": int(stats.k0), "bits_per_symbol": int(stats.bits_per_symbol),
                   "round_elems": [int(stats.round_elems[i]) for i in range(int(stats.rounds))],
                   "round_passes": [int(stats.round_passes[i]) for i in range(int(stats.rounds))],
                   "alg_bytes": int(stats.alg_bytes)},
            "queries": queries,
            "queries_c4": c4,
            "wall_s_timed_region": wall,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
