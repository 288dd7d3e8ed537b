# Round-end check on one GPU: the GPU test suite, smoke, the C2 / C1 / C3 bench lines and the ncu launch list of a C2 build
# (outputs under gpurun_out/; the lines kept for the record are copied to profiles/).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench_c2.json 2> gpurun_out/r02f_bench_c2.err; echo bench rc=$?
python bench.py --steps 20 --warmup 5 --workload c1 --no-c3 > gpurun_out/r02f_bench_c1.json 2> gpurun_out/r02f_bench_c1.err; echo bench rc=$?
python bench.py --steps 20 --warmup 5 --workload c3 --no-c3 --no-cpu-baseline > gpurun_out/r02f_bench_c3.json 2> gpurun_out/r02f_bench_c3.err; echo bench rc=$?
CMD="python bench.py --steps 2 --warmup 1 --no-queries --no-cpu-baseline --no-c3"
$CMD > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02f_launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches rc=$?
for f in c2 c1 c3; do python tools/bench_summary.py gpurun_out/r02f_bench_$f.json; done
