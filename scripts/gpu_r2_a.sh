#!/bin/bash
# round 2, first GPU pass: tile-kernel correctness (small -> large), parity suite, A/B timing at scale 22
mkdir -p gpurun_out
timeout 300 python scripts/tile_check.py 8 10 13 > gpurun_out/tile_check_small.log 2>&1; echo "tile_check_small rc=$?"
tail -5 gpurun_out/tile_check_small.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 300 python scripts/tile_check.py 16 18 > gpurun_out/tile_check_big.log 2>&1; echo "tile_check_big rc=$?"
tail -8 gpurun_out/tile_check_big.log
timeout 600 python scripts/mxm_ab.py 22 '{"spgemm_tile":"0"}' '{}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512"}' '{"spgemm_tile_ctas":"3"}' '{"spgemm_tile_scap":"1024"}' '{"spgemm_tile_scap":"3072"}' '{"spgemm_tile_tf8":"10"}' '{"spgemm_tile_threads":"384"}' > gpurun_out/mxm_ab.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab.log
