#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tile_check.py 10 14 > gpurun_out/tile_check_small.log 2>&1; echo "tile_check rc=$?"; tail -2 gpurun_out/tile_check_small.log
timeout 300 python scripts/tile_dbg.py 20 '{"spgemm_tile":"0"}' '{}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512"}' '{"spgemm_tile_ctas":"3"}' > gpurun_out/tile_dbg.log 2>&1; echo "tile_dbg rc=$?"
cat gpurun_out/tile_dbg.log
timeout 600 python scripts/mxm_ab.py 22 '{"spgemm_tile":"0"}' '{}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512"}' '{"spgemm_tile_ctas":"3"}' '{"spgemm_tile_scap":"1024"}' '{"spgemm_tile_scap":"2688"}' '{"spgemm_tile_threads":"384"}' '{"spgemm_tile_threads":"128","spgemm_tile_ctas":"3"}' > gpurun_out/mxm_ab.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
