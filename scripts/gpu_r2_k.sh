#!/bin/bash
# row-end CSR result: parity + A/B + bench
mkdir -p gpurun_out
S=$(date +%s)
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_boundary.py tests/test_gpu_fullscale.py -m gpu -x -q > gpurun_out/pytest_k.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - S )) s)"
tail -5 gpurun_out/pytest_k.log
timeout 600 python scripts/mxm_ab.py 22 '{}' '{"spgemm_row_end": "0"}' '{"spgemm_table_factor8": "24"}' '{"spgemm_table_factor8": "18"}' > gpurun_out/mxm_ab_k.log 2>&1; echo "mxm_ab rc=$?"
cat gpurun_out/mxm_ab_k.log
S=$(date +%s)
timeout 900 python bench.py --no-scale25 > gpurun_out/bench_k.log 2> gpurun_out/bench_k.err; echo "bench rc=$? ($(( $(date +%s) - S )) s)"
tail -c 1500 gpurun_out/bench_k.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/bench_k.log').read().strip().splitlines()[-1])
    print({k: l[k] for k in ('value', 'ms_per_step', 'value_compact')}, l['roofline']['frac'])
    e = l['e2e']; print(e['ms_per_step'], e['single_call']['ms_per_step'], e['same_result'], l['e2e_reference_interface'])
    print(l['mxv']['ms_per_iter'], l['mxv']['roofline']['kernel'], l['mxv']['roofline']['kernel_us'])
    print(l['phases_ms'])
except Exception as e:
    print('parse failed', e)
PY
