mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mxm or rmat or goldens or power or aliasing or inner" -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fullscale.py -m gpu -q -x -k "mxm" -p no:cacheprovider 2>&1 | tail -3
timeout 600 python scripts/mxm_ab.py 22 > gpurun_out/mxm_ab.log 2>&1; grep -v host: gpurun_out/mxm_ab.log | cut -c1-140
