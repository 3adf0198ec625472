#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tile_dbg.py 20 '{"spgemm_tile":"0"}' '{}' '{"spgemm_tile_poll":"1"}' '{"spgemm_cas_first":"0"}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512"}' '{"spgemm_tile_ctas":"1","spgemm_tile_threads":"512","spgemm_tile_poll":"1"}' > gpurun_out/tile_dbg.log 2>&1; echo "tile_dbg rc=$?"
cat gpurun_out/tile_dbg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spgemm_tile -c 1 -o gpurun_out/tile18 python scripts/tile_dbg.py 18 > gpurun_out/ncu_tile18.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_tile18.log
