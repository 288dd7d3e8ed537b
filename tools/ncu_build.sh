set -x
CMD="python bench.py --steps 2 --warmup 1 --no-queries --no-cpu-baseline --no-c3"
$CMD > gpurun_out/ncu_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:onesweep64 -s 8 -c 2 -o gpurun_out/r02_onesweep $CMD > gpurun_out/ncu_onesweep.log 2>&1
echo onesweep rc=$?
ncu --set full --clock-control none -k regex:"sa_pack0|seg_apply|seg_reduce|bwt_gather|wt_levels|ssa_|sa_keybuild" -s 0 -c 12 -o gpurun_out/r02_build_others $CMD > gpurun_out/ncu_others.log 2>&1
echo others rc=$?
python tools/profile_r02.py > gpurun_out/ncu_plain_r02.log 2>&1 &&
ncu --set full --clock-control none -k regex:"dsa_pack|dsa_keybuild_ext" -c 10 -o gpurun_out/r02_dsa_pack python tools/profile_r02.py > gpurun_out/ncu_new1.log 2>&1
echo new1 rc=$?
ncu --set full --clock-control none -k regex:"dsa_keybuild_dbl|dsa_isa_publish" -s 8 -c 4 -o gpurun_out/r02_dsa_dbl python tools/profile_r02.py > gpurun_out/ncu_new2.log 2>&1
echo new2 rc=$?
ncu --set full --clock-control none -k regex:"hk_count|hk_heads|hk_sum|rrr_|ranges_push" -c 14 -o gpurun_out/r02_hk_rrr_push python tools/profile_r02.py > gpurun_out/ncu_new3.log 2>&1
echo new3 rc=$?
for r in r02_build_others r02_dsa_pack r02_dsa_dbl r02_hk_rrr_push; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep gpurun_out/$r.txt && rm -f gpurun_out/$r.ncu-rep
done
python tools/ncu_summary.py gpurun_out/r02_onesweep.ncu-rep gpurun_out/r02_onesweep.txt
ncu -i gpurun_out/r02_onesweep.ncu-rep --page source --csv > gpurun_out/r02_onesweep_source.csv 2>/dev/null
ls -la gpurun_out/ | head -40; du -sh gpurun_out
