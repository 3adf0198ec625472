#!/bin/bash
# balanced per-warp ranges + pipelined loads (main) vs no pipeline (variant); parity of the mxm tests
mkdir -p gpurun_out
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullscale.py -m gpu -x -q -k "all_bins or rmat_parity or mxm_vs_oracle or row_end or headline or mxm" > gpurun_out/pytest_m.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - S )) s)"
tail -4 gpurun_out/pytest_m.log
timeout 600 python scripts/mxm_ab.py 22 '{}' '{"spgemm_thr_6": "256"}' '{"spgemm_thr_5": "64", "spgemm_thr_6": "64"}' '{"spgemm_thr_7": "128", "spgemm_thr_8": "128"}' > gpurun_out/mxm_ab_m.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab_m.log
GRB_CUDA_LIB=$PWD/python-graphblas_b200/graphblas_b200/variants/libgrb_cuda_nopipe.so timeout 600 python scripts/mxm_ab.py 22 '{}' > gpurun_out/mxm_ab_m2.log 2>&1; echo "mxm_ab nopipe rc=$?"
cat gpurun_out/mxm_ab_m2.log
