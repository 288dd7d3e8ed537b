set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_sa_check.py --size 50000000 --kind 0 --verify --profile > gpurun_out/r2_dist_n2_eng50M.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_dist_n2_eng50M.log
timeout 300 $TR tools/dist_sa_check.py --size 12000000 --kind 2 --verify --profile > gpurun_out/r2_dist_n2_miss12M.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_dist_n2_miss12M.log
timeout 300 $TR tools/dist_sa_check.py --size 64000000 --kind 3 --verify --profile --sa-only > gpurun_out/r2_dist_n2_repeat64M.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_dist_n2_repeat64M.log
timeout 300 $TR tools/dist_sa_check.py --size 100000000 --kind 1 --verify --wide --profile > gpurun_out/r2_dist_n2_dna100M_wide.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_dist_n2_dna100M_wide.log
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err; echo bench rc=$?; tail -c 1500 gpurun_out/r2_bench_n2_a.err | grep -v NCCL | tail -20
