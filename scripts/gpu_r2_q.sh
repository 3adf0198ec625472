#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus ${NG:-2} --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench22_n${NG:-2}.log 2> gpurun_out/bench22_n${NG:-2}.err; echo "bench n2 rc=$? ($(( $(date +%s) - S )) s)"
tail -5 gpurun_out/bench22_n${NG:-2}.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/bench22_n' + __import__('os').environ.get('NG', '2') + '.log').read().strip().splitlines()[-1])
    print({k: l[k] for k in ('value', 'ms_per_step', 'n_gpus')})
    print('mxv', l['mxv']['ms_per_iter'], l['mxv']['ms_per_iter_by_exchange'])
    print('workloads', json.dumps(l['workloads'])[:1800])
    print('scale25', json.dumps(l['scale25'])[:2500])
except Exception as e:
    print('parse failed', e)
PY
