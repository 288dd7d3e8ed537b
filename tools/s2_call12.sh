#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/s2c12_gpu_tests.log
cat gpurun_out/s2c12_gpu_tests.log
for b in default 40 48; do
  if [ $b = default ]; then unset HKCSA_BITS0; else export HKCSA_BITS0=$b; fi
  python bench.py --no-cpu-baseline --no-queries --no-c3 --steps 8 > gpurun_out/s2c12_bench_c2_$b.json 2> gpurun_out/s2c12_bench_c2.err
  python tools/bench_summary.py gpurun_out/s2c12_bench_c2_$b.json
  python -c "
import json;d=json.loads(open('gpurun_out/s2c12_bench_c2_$b.json').read().strip().splitlines()[-1]);print(d['sa'])"
done
unset HKCSA_BITS0
python bench.py --workload c3 --no-cpu-baseline --no-queries --steps 6 > gpurun_out/s2c12_bench_c3.json 2> gpurun_out/s2c12_bench_c3.err
python tools/bench_summary.py gpurun_out/s2c12_bench_c3.json
python tools/build_timeline.py c2 2>&1 | head -8
