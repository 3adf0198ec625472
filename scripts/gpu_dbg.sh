timeout 600 python scripts/debug_mxv.py 2>&1 | tail -40
