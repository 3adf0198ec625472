"""Per-kernel time of the iterative workloads (BFS as the reference notebook writes it), library profile mode."""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
nnz = c.numel()
deg = ip[1:] - ip[:-1]
src = int(torch.nonzero(deg > 0)[0])
A = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(nnz, dtype=torch.bool, device=dev), n, n)
gb.cuda.matrix_build_transpose(A)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def bfs(trace=False):
    q = gb.Vector.from_coo([src], [True], size=n)
    v = gb.Vector(gb.dtypes.INT64, n)
    out = []
    for level in range(1, n):
        if trace:
            torch.cuda.synchronize(); e0.record()
        v(mask=q.V)[:] = level
        q(~v.S, replace=True) << q.vxm(A, gb.semiring.any_pair)
        nq = q.nvals
        if trace:
            e1.record(); torch.cuda.synchronize(); out.append((level, nq, round(e0.elapsed_time(e1) * 1e3)))
        if nq == 0:
            break
    return out

bfs(); bfs()
print("BFS per level (level, next frontier, us):", bfs(trace=True))
gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
bfs()
kt = gb.cuda.kernel_times(reset=True); gb.cuda.set_option("profile", "0")
print("BFS kernels (us total, launches):", {k: (round(v[0] * 1e3), v[1]) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][0])})
torch.cuda.synchronize(); t0 = time.perf_counter(); bfs(); torch.cuda.synchronize(); print("BFS wall ms", (time.perf_counter() - t0) * 1e3)
