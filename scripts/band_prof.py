import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
method = sys.argv[2] if len(sys.argv) > 2 else "band"
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
v = bench.values_torch(c.numel(), 45, torch.float32, device=dev)
A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
x = gb.cuda.vector_from_torch(bench.values_torch(n, 46, torch.float32, device=dev))
gb.cuda.set_option("spmv", method)
for _ in range(4):
    y = A.mxv(x, gb.semiring.plus_times).new()
torch.cuda.synchronize()
