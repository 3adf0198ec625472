mkdir -p gpurun_out
export GRB_CUDA_SPGEMM_TABLE_FACTOR8=20
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spgemm_block_kernel -s 7 -c 7 -o gpurun_out/prof_spgemm_r03 python scripts/prof_driver.py mxm 20 2 > gpurun_out/prof_mxm2.log 2>&1; tail -2 gpurun_out/prof_mxm2.log
