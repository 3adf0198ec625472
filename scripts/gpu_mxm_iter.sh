mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mxm or rmat or goldens or power or aliasing" -p no:cacheprovider 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-mxv > gpurun_out/bench22_mxm.log 2>&1; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench22_mxm.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('ms_per_step',d['ms_per_step'],'value',d['value']); print(d['phases_ms'])
else: print(open('gpurun_out/bench22_mxm.log').read()[-2000:])
PY
