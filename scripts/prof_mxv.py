import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
import os
gb.cuda.set_option("spmv_hot_kb", os.environ.get("GRB_HOT_KB", "132"))
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
ip, c, n = bench.rmat_csr_torch(22, bench.RMAT_2B, 42, device=dev)
v = bench.values_torch(c.numel(), 45, torch.float32, device=dev)
A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
x = gb.cuda.vector_from_torch(bench.values_torch(n, 46, torch.float32, device=dev))
for hot in sys.argv[1:] or ["0", "1"]:
    gb.cuda.set_option("spmv_hot", hot)
    for _ in range(3):
        y = A.mxv(x, gb.semiring.plus_times).new()
torch.cuda.synchronize()
print("ok", y.nvals)
