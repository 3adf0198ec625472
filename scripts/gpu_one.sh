python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-mxv --no-workloads 2>&1 | tail -5 | cut -c1-600
