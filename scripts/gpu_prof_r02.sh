# round-2 profile pass: per-kernel time + DRAM traffic (mxm, mxv), launch list of the bench command, full captures of the top kernels
mkdir -p gpurun_out /tmp/prof
export GRB_CUDA_SPMV_TRIAL=0
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active"
timeout 900 ncu --metrics $M --clock-control none -k regex:"spgemm|compact_rows|row_flops|bin_|row_end|split_parts" --csv --log-file gpurun_out/ncu_mxm22_r02.csv python scripts/prof_driver.py mxm 22 2 > gpurun_out/p1.log 2>&1; tail -n 1 gpurun_out/p1.log
timeout 900 ncu --metrics $M --clock-control none -k regex:"spmv|merge_search|seg_" --csv --log-file gpurun_out/ncu_mxv22_r02.csv python scripts/prof_driver.py mxv 22 4 > gpurun_out/p2.log 2>&1; tail -n 1 gpurun_out/p2.log
unset GRB_CUDA_SPMV_TRIAL
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_r02.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-workloads --no-scale25 > gpurun_out/bench_under_ncu.log 2>&1; tail -n 1 gpurun_out/bench_under_ncu.log | cut -c1-200
export GRB_CUDA_SPMV_TRIAL=0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spgemm_block_kernel -s 5 -c 1 -o /tmp/prof/full_spgemm python scripts/prof_driver.py mxm 22 1 > gpurun_out/p3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_kernel -s 2 -c 1 -o /tmp/prof/full_spmv python scripts/prof_driver.py mxv 22 4 > gpurun_out/p4.log 2>&1
for k in full_spgemm full_spmv; do
  ncu -i /tmp/prof/$k.ncu-rep --page raw --csv > gpurun_out/${k}_raw.csv 2>/dev/null
  ncu -i /tmp/prof/$k.ncu-rep --page details > gpurun_out/${k}_details.txt 2>/dev/null
  ncu -i /tmp/prof/$k.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${k}_source.csv.gz
done
ls -la gpurun_out | tail -n 12
