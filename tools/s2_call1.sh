#!/bin/bash
# session-2 call 1: new one-call SA+BWT path and one-pass sampled SA: tests, then C2 / C3 bench with 56- vs 64-bit keys
set -x
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "one_call or sampled_sa_one_pass or sa_bwt" 2>&1 | tail -8 > gpurun_out/s2c1_newtests.log
cat gpurun_out/s2c1_newtests.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2c1_gpu_tests.log
cat gpurun_out/s2c1_gpu_tests.log
python bench.py --no-cpu-baseline --no-queries > gpurun_out/s2c1_bench_c2.json 2> gpurun_out/s2c1_bench_c2.err
python tools/bench_summary.py gpurun_out/s2c1_bench_c2.json
python bench.py --workload c3 --no-cpu-baseline --no-queries --steps 8 > gpurun_out/s2c1_bench_c3_carry56.json 2> gpurun_out/s2c1_bench_c3.err
python tools/bench_summary.py gpurun_out/s2c1_bench_c3_carry56.json
HKCSA_CARRY56=0 python bench.py --workload c3 --no-cpu-baseline --no-queries --steps 8 > gpurun_out/s2c1_bench_c3_gather64.json 2>> gpurun_out/s2c1_bench_c3.err
python tools/bench_summary.py gpurun_out/s2c1_bench_c3_gather64.json
