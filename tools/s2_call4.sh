#!/bin/bash
# session-2 call 4: one-call index build, sampled cut points, coalesced seg_scan: tests + C2 / C3 bench
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2c4_gpu_tests.log
cat gpurun_out/s2c4_gpu_tests.log
python bench.py --no-cpu-baseline --no-queries --no-c3 > gpurun_out/s2c4_bench_c2.json 2> gpurun_out/s2c4_bench_c2.err
python tools/bench_summary.py gpurun_out/s2c4_bench_c2.json
python bench.py --workload c3 --no-cpu-baseline --no-queries --steps 8 > gpurun_out/s2c4_bench_c3.json 2> gpurun_out/s2c4_bench_c3.err
python tools/bench_summary.py gpurun_out/s2c4_bench_c3.json
python bench.py --workload c1 --no-cpu-baseline --no-queries > gpurun_out/s2c4_bench_c1.json 2> gpurun_out/s2c4_bench_c1.err
python tools/bench_summary.py gpurun_out/s2c4_bench_c1.json
