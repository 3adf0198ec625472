#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_band_kernel -s 2 -c 1 -o gpurun_out/band22 python scripts/band_prof.py 22 band > gpurun_out/ncu_band22.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_band22.log
