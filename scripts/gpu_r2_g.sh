#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "banded or vector_multiplies or rmat_parity or complemented or tiled or float_noninteger" > gpurun_out/pytest_band.log 2>&1; echo "pytest band rc=$?"
tail -15 gpurun_out/pytest_band.log
timeout 600 python scripts/band_ab.py 22 > gpurun_out/band_ab.log 2>&1; echo "band_ab rc=$?"
cat gpurun_out/band_ab.log
