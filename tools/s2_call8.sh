#!/bin/bash
# session-2 call 8 (2 GPUs): N=2 bench with the sampled cut points, one-call build, new gather
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/dist_sa_check.py --size 100000000 --kind 0 --verify --profile > gpurun_out/s2c8_dist_n2_eng100M.log 2>&1; echo rc=$?; tail -1 gpurun_out/s2c8_dist_n2_eng100M.log | cut -c1-900
timeout 300 $TR tools/dist_sa_check.py --size 100000000 --kind 1 --verify --wide --profile > gpurun_out/s2c8_dist_n2_dna100M_wide.log 2>&1; echo rc=$?; tail -1 gpurun_out/s2c8_dist_n2_dna100M_wide.log | cut -c1-900
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s2c8_bench_n2.json 2> gpurun_out/s2c8_bench_n2.err; echo bench rc=$?
python tools/bench_summary.py gpurun_out/s2c8_bench_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2c8_bench_n2.json').read().strip().splitlines()[-1])
db=d['dist_build']
for k in ['parity','ms','MBps','single_gpu','fraction_of_linear','phases_ms_max_over_ranks','rounds_rank0','properties_ok']:
    print(k, db.get(k))
print(d['queries_c4']['all_answers_on_all_ranks'])
PY
