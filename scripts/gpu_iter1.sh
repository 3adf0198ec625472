mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "mxm or rmat or goldens or power or aliasing" -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_fullscale.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
timeout 600 python scripts/mxm_ab.py 22 > gpurun_out/mxm_ab.log 2>&1; cat gpurun_out/mxm_ab.log | cut -c1-600
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-workloads > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('mxm ms', d['ms_per_step'], 'mxv', d['mxv']['ms_per_iter'], d['mxv']['roofline']['frac'])"
