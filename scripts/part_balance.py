"""How well does the cost-weighted row partition of the mxm balance the ranks?  One GPU times every block of an N-way partition
in turn (python scripts/part_balance.py [scale] [N]) for several weights of the split rows."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
from graphblas_b200 import distributed as D
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
N = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2A, 42, device=dev)
v = bench.values_torch(c.numel(), 43, torch.float32, device=dev)
B = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
deg = ip[1:] - ip[:-1]
rowflops = torch.zeros(n, dtype=torch.int64, device=dev)
rowflops.index_add_(0, torch.repeat_interleave(torch.arange(n, device=dev), deg), deg[c.long()])
print("rows above the split threshold:", int((rowflops > 10440).sum()), "their flops:", int(rowflops[rowflops > 10440].sum()), "of", int(rowflops.sum()))
sr = gb.semiring.plus_times
for reread in (0.0, 0.25, 0.5, 1.0, 2.0):
    cost = torch.zeros(n + 1, dtype=torch.float64, device=dev)
    cost[1:] = D.mxm_row_costs(rowflops, reread=reread)
    b = D.row_blocks_by_prefix(torch.cumsum(cost, 0).cpu().numpy(), N)
    times = []
    for g in range(N):
        A = D.local_block(gb, ip, c, v, n, b[g], b[g + 1])
        C = A.mxm(B, sr).new(); C = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(3):
            C = None
            C = A.mxm(B, sr).new()
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / 3)
        C = None
    print(f"reread={reread:4.2f}  max {max(times):6.2f} ms  mean {sum(times)/N:6.2f} ms  balance {sum(times)/N/max(times):.3f}  blocks: " + " ".join(f"{t:.2f}" for t in times), flush=True)
