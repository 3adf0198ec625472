#!/bin/bash
# split heavy rows: parity + A/B; full ncu capture of the heaviest CTA bin at scale 22
mkdir -p gpurun_out /tmp/prof
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullscale.py -m gpu -x -q -k "all_bins or rmat_parity or mxm_vs_oracle or row_end or headline or fullscale or mxm" > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - S )) s)"
tail -4 gpurun_out/pytest_l.log
timeout 600 python scripts/mxm_ab.py 22 '{}' '{"spgemm_split": "0"}' '{}' > gpurun_out/mxm_ab_l.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab_l.log
export GRB_CUDA_SPMV_TRIAL=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spgemm_block_kernel -s 5 -c 1 -o /tmp/prof/full_bin6 python scripts/prof_driver.py mxm 22 1 > gpurun_out/p_bin6.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/p_bin6.log
ncu -i /tmp/prof/full_bin6.ncu-rep --page details > gpurun_out/full_bin6_details.txt 2>/dev/null
ncu -i /tmp/prof/full_bin6.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/full_bin6_source.csv.gz
ncu -i /tmp/prof/full_bin6.ncu-rep --page raw --csv > gpurun_out/full_bin6_raw.csv 2>/dev/null
python scripts/ncu_hot.py gpurun_out/full_bin6_source.csv.gz 45 > gpurun_out/full_bin6_hot.txt 2>&1
grep -E "Duration|Registers Per|Achieved Occ|Theoretical Occ|Warp Cycles Per Issued|Eligible Warps|L1/TEX Hit|L2 Hit|DRAM Throughput|Issue Slots Busy|Block Size|Grid Size|Dynamic Shared" gpurun_out/full_bin6_details.txt | head -20
