#!/bin/bash
# single GPU: whole parity suite + the new bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench22.log 2> gpurun_out/bench22.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench22.log; tail -5 gpurun_out/bench22.err
