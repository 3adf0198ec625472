mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -k regex:"spgemm|compact_rows|row_flops|bin_" --csv --log-file gpurun_out/ncu_mxm22_r01.csv python scripts/prof_driver.py mxm 22 2 > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
timeout 900 ncu --metrics $M --clock-control none -k regex:"spmv|merge_search" --csv --log-file gpurun_out/ncu_mxv22_r01.csv python scripts/prof_driver.py mxv 22 4 > gpurun_out/p2.log 2>&1; tail -1 gpurun_out/p2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_kernel -s 2 -c 1 -o gpurun_out/prof_spmv_merge_r01b python scripts/prof_driver.py mxv 22 4 > gpurun_out/p3.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench22.log 2>&1; tail -1 gpurun_out/bench22.log | cut -c1-300
