mkdir -p gpurun_out /tmp/prof
timeout 600 python scripts/debug_elect.py 14 17 20 > gpurun_out/debug_elect.log 2>&1; cat gpurun_out/debug_elect.log | tail -30
# ncu: warp-elect kernel (bin 5 of the 2nd call) and block-elect kernel (bin 7 of the 2nd call), scale 20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spgemm_warp_elect_kernel" -s 9 -c 1 -o /tmp/prof/mxm_warp python scripts/prof_driver.py mxm 20 2 > gpurun_out/prof_mxm_warp.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spgemm_block_elect_kernel" -s 7 -c 1 -o /tmp/prof/mxm_block python scripts/prof_driver.py mxm 20 2 > gpurun_out/prof_mxm_block.log 2>&1
for k in warp block; do
  ncu -i /tmp/prof/mxm_$k.ncu-rep --page raw --csv > gpurun_out/mxm_${k}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/mxm_$k.ncu-rep --page details > gpurun_out/mxm_${k}_details.txt 2>/dev/null
  ncu -i /tmp/prof/mxm_$k.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/mxm_${k}_source.csv.gz
done
tail -2 gpurun_out/prof_mxm_warp.log gpurun_out/prof_mxm_block.log
