set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
bash tools/sort_shapes.sh > gpurun_out/r02_onesweep_shapes.txt 2>&1; cat gpurun_out/r02_onesweep_shapes.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; echo bench rc=$?
N=30000000 python tools/stress_check.py > gpurun_out/r02_stress_check.txt 2>&1; tail -12 gpurun_out/r02_stress_check.txt
python tools/profile_r02.py > gpurun_out/r02_profile_plain.log 2>&1; echo profile rc=$?; tail -8 gpurun_out/r02_profile_plain.log
