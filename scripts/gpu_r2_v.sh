#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullscale.py -m gpu -x -q -k "mxm or rmat_parity or headline or power or row_end or all_bins" > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_r.log
timeout 600 python scripts/mxm_ab.py 22 '{}' '{}' 2>&1 | grep -v host
for v in v40; do
  echo "variant $v"
  GRB_CUDA_LIB=$PWD/python-graphblas_b200/graphblas_b200/variants/libgrb_cuda_$v.so timeout 600 python scripts/mxm_ab.py 22 '{}' '{}' 2>&1 | grep -v host
done
