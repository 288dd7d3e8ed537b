"""The reference's benchmark harness (tests/benchmark.py) restated for this package: same entry points and result
fields -- ``benchmark_construction`` (:25-38), ``benchmark_pattern_search`` (:40-52), ``run_full_benchmark`` (:54-89),
``print_benchmark_summary`` (:91-107), ``BenchmarkResults`` (:16-23) -- around ``csa.csa.CompressedSuffixArray``.

The reference file itself cannot run anywhere: it imports ``memory_profiler`` (not installed, and only used as a
decorator) and a ``CompressedSuffixArray`` its csa/csa.py never defines.  tests/test_gpu_dropin_api.py runs the
ORIGINAL file unchanged against this package when the reference checkout is present (shimming memory_profiler), and this
restatement otherwise (the GPU box has no reference checkout).

Memory is reported twice: host RSS growth, as the reference measures it, and device bytes held by the index.
"""
import gc
import os
import time

import psutil

from benchmarks.patterns import generate_random_patterns
from csa.csa import CompressedSuffixArray


def get_process_memory():
    """Resident set size of this process in MB."""
    return psutil.Process(os.getpid()).memory_info().rss / (1024 * 1024)


class BenchmarkResults:
    def __init__(self):
        self.construction_time = 0
        self.construction_memory = 0
        self.pattern_times = {}
        self.pattern_memory = {}
        self.total_time = 0
        self.peak_memory = 0
        self.index_device_bytes = 0


def benchmark_construction(text, epsilon=0.5):
    """(csa, seconds, host MB) of one index build."""
    gc.collect()
    before = get_process_memory()
    t0 = time.time()
    csa = CompressedSuffixArray(text, epsilon=epsilon)
    return csa, time.time() - t0, get_process_memory() - before


def benchmark_pattern_search(csa, pattern):
    """(locations, seconds, host MB) of one locate."""
    gc.collect()
    before = get_process_memory()
    t0 = time.time()
    locations = csa.locate(pattern)
    return locations, time.time() - t0, get_process_memory() - before


def run_full_benchmark(text, pattern_lengths=(5, 10, 50, 100, 500, 1000), iterations=3, verbose=True):
    say = print if verbose else (lambda *a, **k: None)
    results = BenchmarkResults()
    say("\nBenchmarking CSA Construction...")
    csa, results.construction_time, results.construction_memory = benchmark_construction(text)
    results.index_device_bytes = csa.index_bytes()
    say(f"Construction Time: {results.construction_time:.4f} seconds")
    say(f"Construction Memory: {results.construction_memory:.2f} MB (host), {results.index_device_bytes / 1e6:.2f} MB (device index)")
    say("\nBenchmarking Pattern Searches...")
    for pattern in generate_random_patterns(text, list(pattern_lengths)):
        times, mems = [], []
        say(f"\nPattern length: {len(pattern)}")
        for it in range(iterations):
            locations, dt, dm = benchmark_pattern_search(csa, pattern)
            times.append(dt)
            mems.append(dm)
            say(f"Iteration {it + 1}: Time={dt:.4f}s, Memory={dm:.2f}MB")
            say(f"Found {len(locations)} occurrences")
        results.pattern_times[len(pattern)] = sum(times) / iterations
        results.pattern_memory[len(pattern)] = sum(mems) / iterations
    results.total_time = results.construction_time + sum(results.pattern_times.values())
    results.peak_memory = max([results.construction_memory] + list(results.pattern_memory.values()))
    return results


def print_benchmark_summary(results):
    print("\n=== Benchmark Summary ===")
    print(f"\nConstruction:\nTime: {results.construction_time:.4f} seconds\nMemory: {results.construction_memory:.2f} MB")
    print("\nPattern Search (averages):\nPattern Length | Time (s) | Memory (MB)\n" + "-" * 40)
    for length in sorted(results.pattern_times):
        print(f"{length:>13} | {results.pattern_times[length]:>8.4f} | {results.pattern_memory[length]:>10.2f}")
    print(f"\nOverall:\nTotal Time: {results.total_time:.4f} seconds\nPeak Memory: {results.peak_memory:.2f} MB")


if __name__ == "__main__":
    print_benchmark_summary(run_full_benchmark("mississippi$" * 1000))      # the reference's default workload (:110)
