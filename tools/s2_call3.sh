#!/bin/bash
# session-2 call 3 (2 GPUs): GPU tests, distributed-build checks (32- and 64-bit ids, repetitive texts), N=2 bench
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s2c3_gpu_tests.log
cat gpurun_out/s2c3_gpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_sa_check.py --size 50000000 --kind 0 --verify --profile > gpurun_out/s2c3_dist_n2_eng50M.log 2>&1; echo rc=$?; tail -3 gpurun_out/s2c3_dist_n2_eng50M.log
timeout 300 $TR tools/dist_sa_check.py --size 12000000 --kind 2 --verify --profile > gpurun_out/s2c3_dist_n2_miss12M.log 2>&1; echo rc=$?; tail -3 gpurun_out/s2c3_dist_n2_miss12M.log
timeout 300 $TR tools/dist_sa_check.py --size 100000000 --kind 1 --verify --wide --profile > gpurun_out/s2c3_dist_n2_dna100M_wide.log 2>&1; echo rc=$?; tail -3 gpurun_out/s2c3_dist_n2_dna100M_wide.log
timeout 300 $TR tools/dist_sa_check.py --size 1000000000 --kind 0 --wide --profile > gpurun_out/s2c3_dist_n2_eng1G_wide.log 2>&1; echo rc=$?; tail -5 gpurun_out/s2c3_dist_n2_eng1G_wide.log
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s2c3_bench_n2.json 2> gpurun_out/s2c3_bench_n2.err; echo bench rc=$?; tail -c 1500 gpurun_out/s2c3_bench_n2.err | grep -v NCCL | tail -20
python tools/bench_summary.py gpurun_out/s2c3_bench_n2.json
