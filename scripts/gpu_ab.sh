mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mxm or rmat or goldens or power or aliasing" -p no:cacheprovider 2>&1 | tail -2
run() { python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-mxv --no-workloads 2>&1 | python -c "
import json,sys
l=[x for x in sys.stdin if x.startswith('{')]
d=json.loads(l[-1]); k=d['phases_ms']['kernels']; print('$1', 'ms', round(d['ms_per_step'],2), {a:round(b,2) for a,b in k.items() if b>0.3})"; }
run t1024
run t1024_again
