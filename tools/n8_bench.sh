# bench.py on N GPUs of one node (default 8): the line incl. the dist_build record lands in gpurun_out/r2_bench_n$N.json
set -x
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo bench rc=$?
grep -v "NCCL INFO" gpurun_out/r2_bench_n$N.err | tail -15
grep -c "NCCL INFO" gpurun_out/r2_bench_n$N.err; grep "NVLS\|nranks" gpurun_out/r2_bench_n$N.err | head -5
head -c 600 gpurun_out/r2_bench_n$N.json
