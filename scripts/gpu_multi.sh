mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N1 value',d['value'],'ms',d['ms_per_step']); print(json.dumps(d.get('workloads'))[:1500]); print('e2e',d['e2e']); print('mxv',d['mxv']['ms_per_iter'], d['mxv']['roofline']['frac']); print('cpu', d['cpu_baseline'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N2 value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']); print('mxv',d['mxv']['ms_per_iter'], d['mxv']['GB_per_s'])" || tail -20 gpurun_out/bench_n2.log
