#!/bin/bash
mkdir -p gpurun_out /tmp/prof
export GRB_CUDA_SPMV_TRIAL=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spgemm_rows_kernel -s 5 -c 1 -o /tmp/prof/full_rows python scripts/prof_driver.py mxm 22 1 > gpurun_out/p_rows.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof/full_rows.ncu-rep --page details > gpurun_out/full_rows_details.txt 2>/dev/null
ncu -i /tmp/prof/full_rows.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/full_rows_source.csv.gz
ncu -i /tmp/prof/full_rows.ncu-rep --page raw --csv > gpurun_out/full_rows_raw.csv 2>/dev/null
python scripts/ncu_hot.py gpurun_out/full_rows_source.csv.gz 50 > gpurun_out/full_rows_hot.txt 2>&1
grep -E "Duration|Registers Per|Achieved Occ|Theoretical Occ|Warp Cycles Per Issued|Eligible Warps|L1/TEX Hit|L2 Hit|DRAM Throughput|Issue Slots Busy|Block Size|Grid Size|Dynamic Shared|Executed Ipc Active" gpurun_out/full_rows_details.txt | head -20
