set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; echo bench rc=$?
python bench.py --steps 20 --warmup 5 --workload c3 --no-c3 --no-cpu-baseline > gpurun_out/r2_bench_c3_c.json 2> gpurun_out/r2_bench_c3_c.err; echo bench rc=$?
