mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "mxm_all_bins or optional_kernels or inner_outer or aliasing or io_roundtrips" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|exit" gpurun_out/sanitize_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "optional_kernels or nonsquare_max_plus" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck.log
grep -E "RACECHECK SUMMARY|hazard|passed|failed|exit" gpurun_out/sanitize_racecheck.log | tail -12
