mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "all_bins" -p no:cacheprovider 2>&1 | grep -E "Panic|passed|failed|Error" | head
