#!/bin/bash
# 2 GPUs: NCCL parity test of the partitioned path, then the bench at N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -25 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench22_n2.log 2> gpurun_out/bench22_n2.err; echo "bench n2 rc=$?"
tail -c 5000 gpurun_out/bench22_n2.log; tail -8 gpurun_out/bench22_n2.err
