mkdir -p gpurun_out
timeout 600 python scripts/mxv_ab.py 22 > gpurun_out/mxv_ab_1024.log 2>&1; grep -E "natural|permuted|Error|error" gpurun_out/mxv_ab_1024.log | tail -50
for th in 768; do
  GRB_CUDA_LIBRARY=$PWD/build/variants/libgrb_cuda_$th.so timeout 600 python scripts/mxv_ab.py 22 natural > gpurun_out/mxv_ab_$th.log 2>&1; echo "== $th threads"; grep -E "seg|rror" gpurun_out/mxv_ab_$th.log | tail -24
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullscale.py -m gpu -q -x -p no:cacheprovider -k "vector_multiplies or rmat_parity or degree_identities or goldens or aliasing" > gpurun_out/pytest_mxv.log 2>&1; tail -5 gpurun_out/pytest_mxv.log
