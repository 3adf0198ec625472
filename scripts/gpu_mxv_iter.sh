mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "goldens or vector_multiplies or rmat or bfs" -p no:cacheprovider 2>&1 | tail -2
timeout 600 python scripts/mxv_ab.py 22 2>&1 | tail -8
