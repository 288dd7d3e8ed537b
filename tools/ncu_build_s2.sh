#!/bin/bash
# session-2 profiling pass (one GPU): launch list of the C2 bench, --set full of the dominant kernel and of the build kernels
set -x
CMD="python bench.py --steps 2 --warmup 1 --no-queries --no-cpu-baseline --no-c3"
$CMD > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02s2_launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:onesweep64 -s 8 -c 2 -o gpurun_out/r02s2_onesweep $CMD > gpurun_out/ncu_onesweep.log 2>&1
echo onesweep rc=$?
ncu --set full --clock-control none -k regex:"sa_pack0|gram_|seg_apply|seg_reduce|seg_scan|wt_levels|wt_tile_hist|wt_count_all|wt_dir_fix_all|ssa_mark_pack|ssa_fix_sample|group_sort" -s 0 -c 18 -o gpurun_out/r02s2_build_others $CMD > gpurun_out/ncu_others.log 2>&1
echo others rc=$?
CMD3="python bench.py --workload c3 --steps 1 --warmup 1 --no-queries --no-cpu-baseline"
ncu --set full --clock-control none -k regex:"wt_levels|wt_tile_hist|bwt_gather|sa_keybuild|seg_apply|sa_pack0" -s 0 -c 10 -o gpurun_out/r02s2_build_c3 $CMD3 > gpurun_out/ncu_c3.log 2>&1
echo c3 rc=$?
rm -f gpurun_out/r02s2_*.txt
for r in r02s2_build_others r02s2_build_c3; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r.txt && rm -f gpurun_out/$r.ncu-rep
done
python tools/ncu_summary.py gpurun_out/r02s2_onesweep.ncu-rep gpurun_out/r02s2_onesweep.txt
ls -la gpurun_out/ | tail -12
