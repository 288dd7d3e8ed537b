N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for early in 0 1; do
  HKCSA_DSA_TEXT_EARLY=$early timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-c3 --no-queries --no-cpu-baseline > gpurun_out/ab_text_$early.json 2> gpurun_out/ab_text_$early.err; echo rc=$?
  python - <<PY
import json
d=json.loads(open('gpurun_out/ab_text_$early.json').read().strip().splitlines()[-1])
db=d['dist_build']
print('early=$early', {k:db[k] for k in ('ms','ms_runs','fraction_of_linear','properties_ok')}); print(db['phases_ms_max_over_ranks']); print(db['parity'])
PY
done
