timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mxm or rmat or goldens" -p no:cacheprovider 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullscale.py -m gpu -q -x -k "masked" -p no:cacheprovider --durations=3 2>&1 | tail -8
