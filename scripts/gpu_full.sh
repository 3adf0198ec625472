set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest.log | tail -20
timeout 900 python bench.py > gpurun_out/bench22.log 2>&1; tail -1 gpurun_out/bench22.log | cut -c1-1500
