#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullscale.py -m gpu -x -q -k "mxm or rmat_parity or headline or power or row_end or all_bins" > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - S )) s)"
tail -4 gpurun_out/pytest_r.log
timeout 600 python scripts/mxm_ab.py 22 '{}' '{"spgemm_rows": "0"}' '{"spgemm_rows_waves": "2"}' '{"spgemm_rows_waves": "4"}' > gpurun_out/mxm_ab_r.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab_r.log
