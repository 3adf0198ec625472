set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
tail -40 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "goldens and auto-auto or nonsquare or all_bins" -p no:cacheprovider > gpurun_out/sanitizer.log 2>&1; echo "sanitizer exit $?" >> gpurun_out/sanitizer.log; tail -15 gpurun_out/sanitizer.log
timeout 400 python bench.py --scale 18 --steps 3 --warmup 3 > gpurun_out/bench18.log 2>&1; tail -3 gpurun_out/bench18.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench22.log 2>&1; tail -3 gpurun_out/bench22.log
