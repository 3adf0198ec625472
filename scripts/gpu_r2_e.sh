#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tile_check.py 10 14 > gpurun_out/tile_check_small.log 2>&1; echo "tile_check rc=$?"; tail -2 gpurun_out/tile_check_small.log
timeout 300 python scripts/tile_dbg.py 20 '{"spgemm_tile":"0"}' '{"spgemm_tile":"0","spgemm_cas_first":"0"}' '{}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512"}' '{"spgemm_tile_tf8":"20"}' '{"spgemm_tile_tf8":"12"}' > gpurun_out/tile_dbg.log 2>&1; echo "tile_dbg rc=$?"
cat gpurun_out/tile_dbg.log
timeout 600 python scripts/mxm_ab.py 22 '{"spgemm_tile":"0"}' '{"spgemm_tile":"0","spgemm_table_factor8":"16"}' '{}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512"}' '{"spgemm_tile_threads":"384"}' '{"spgemm_tile_tf8":"12"}' '{"spgemm_tile_tf8":"20"}' > gpurun_out/mxm_ab.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab.log
timeout 900 python -m pytest tests/test_ffi_adapter.py tests/test_gpu_boundary.py -m gpu -q > gpurun_out/pytest_boundary.log 2>&1; echo "pytest boundary rc=$?"
tail -40 gpurun_out/pytest_boundary.log
