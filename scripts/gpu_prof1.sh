set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r01.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
# full capture of the top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_kernel -s 2 -c 1 -o gpurun_out/prof_spmv_merge_r01 python scripts/prof_driver.py mxv 22 4 > gpurun_out/prof_mxv.log 2>&1; tail -2 gpurun_out/prof_mxv.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spgemm_hash_kernel -s 12 -c 6 -o gpurun_out/prof_spgemm_r01 python scripts/prof_driver.py mxm 20 2 > gpurun_out/prof_mxm.log 2>&1; tail -2 gpurun_out/prof_mxm.log
ls -la gpurun_out
