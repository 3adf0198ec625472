"""fp32 plus_times A.mxv(x) on the bench matrix with the merge-path kernel: kernel time (profile mode) -- used to A/B library variants via GRB_CUDA_LIBRARY."""
import sys, pathlib, os
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
ip, c, n = bench.rmat_csr_torch(22, bench.RMAT_2B, 42, device=dev)
A = gb.cuda.matrix_from_device_csr(ip, c, bench.values_torch(c.numel(), 45, torch.float32, device=dev), n, n)
x = gb.cuda.vector_from_torch(bench.values_torch(n, 46, torch.float32, device=dev))
gb.cuda.set_option("spmv", "merge")
for _ in range(5):
    y = A.mxv(x, gb.semiring.plus_times).new()
gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
for _ in range(20):
    y = A.mxv(x, gb.semiring.plus_times).new()
kt = gb.cuda.kernel_times(reset=True)
print(os.environ.get("GRB_CUDA_LIBRARY", "default"), {k: round(v[0] / v[1] * 1e3, 1) for k, v in kt.items()}, "checksum", float(gb.cuda.vector_as_torch(y)[0].double().sum()))
