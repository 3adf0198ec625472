#!/bin/bash
# full GPU suite + the default bench line (what the driver runs at round end)
mkdir -p gpurun_out
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - S )) s)"
tail -5 gpurun_out/pytest_full.log
S=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_j.log 2> gpurun_out/bench_j.err; echo "bench rc=$? ($(( $(date +%s) - S )) s)"
tail -c 1500 gpurun_out/bench_j.err
python - <<'PY'
import json
try:
    l = json.loads(open('gpurun_out/bench_j.log').read().strip().splitlines()[-1])
    print({k: l[k] for k in ('value', 'ms_per_step')}, l['roofline']['frac'], l['e2e'], l['e2e_reference_interface'])
    print(l['mxv']['ms_per_iter'], l['mxv']['roofline']['kernel'], l['mxv']['roofline']['kernel_us'])
    print(l['phases_ms'])
except Exception as e:
    print('parse failed', e)
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
