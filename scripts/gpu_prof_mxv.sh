mkdir -p gpurun_out /tmp/prof
for hot in 0 1; do
  GRB_HOT_KB=196 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_seg_kernel" -s 2 -c 1 -o /tmp/prof/seg_hot$hot python scripts/prof_mxv.py $hot > gpurun_out/prof_hot$hot.log 2>&1
  ncu -i /tmp/prof/seg_hot$hot.ncu-rep --page raw --csv > gpurun_out/seg2_hot${hot}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/seg_hot$hot.ncu-rep --page details > gpurun_out/seg2_hot${hot}_details.txt 2>/dev/null
  ncu -i /tmp/prof/seg_hot$hot.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/seg2_hot${hot}_source.csv.gz
done
