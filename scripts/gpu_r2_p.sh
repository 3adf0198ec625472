#!/bin/bash
# 2 GPUs: aggregator tests, NCCL parity test of the partitioned path, bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_agg.py -m gpu -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -5 gpurun_out/pytest_multi.log
S=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench22_n2.log 2> gpurun_out/bench22_n2.err; echo "bench n2 rc=$? ($(( $(date +%s) - S )) s)"
tail -5 gpurun_out/bench22_n2.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/bench22_n2.log').read().strip().splitlines()[-1])
    print({k: l[k] for k in ('value', 'ms_per_step', 'n_gpus')}, l['e2e']['ms_per_step'] if l['e2e'] else None)
    print('mxv', l['mxv']['ms_per_iter'], l['mxv']['partition'])
    print('workloads', json.dumps(l['workloads'])[:1500])
    print('scale25', json.dumps(l['scale25'])[:2500])
except Exception as e:
    print('parse failed', e)
PY
