mkdir -p gpurun_out /tmp/prof
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spgemm_group_elect_kernel" -s 1 -c 1 -o /tmp/prof/mxm_group python scripts/prof_driver.py mxm 20 2 > gpurun_out/prof_mxm_group.log 2>&1
ncu -i /tmp/prof/mxm_group.ncu-rep --page raw --csv > gpurun_out/mxm_group_raw.csv 2>/dev/null
ncu -i /tmp/prof/mxm_group.ncu-rep --page details > gpurun_out/mxm_group_details.txt 2>/dev/null
ncu -i /tmp/prof/mxm_group.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/mxm_group_source.csv.gz
tail -n 2 gpurun_out/prof_mxm_group.log
