mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
# memcheck: every SpGEMM bin incl. split rows (rows kernel, split kernel), row-end operands and consumers, tile kernel, masks in the hash,
# banded SpMV, aggregators (mxv / vxm recipes, pow / float unary ops), diag + broadcast, mask algebra
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_agg.py -m gpu -q -x -p no:cacheprovider -k "mxm_all_bins or row_end or complemented_mask or (tiled_kernel and float32 and opts0) or banded or (aggregators and (hypot or mean or count) and fp64) or diag or mask_algebra or inner_outer or aliasing" > gpurun_out/sanitize_memcheck_r02.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck_r02.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|exit" gpurun_out/sanitize_memcheck_r02.log | tail -8
# racecheck (shared-memory hazards) on the default SpGEMM kernels: rows kernel with its barrier-free row hand-over, split kernel
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "mxm_all_bins or nonsquare_max_plus" > gpurun_out/sanitize_racecheck_r02.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitize_racecheck_r02.log
grep -E "RACECHECK SUMMARY|hazard|passed|failed|exit" gpurun_out/sanitize_racecheck_r02.log | tail -12
