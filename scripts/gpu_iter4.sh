mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python scripts/mxv_ab2.py 22 > gpurun_out/mxv_ab2.log 2>&1; cat gpurun_out/mxv_ab2.log
