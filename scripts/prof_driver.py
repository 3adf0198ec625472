"""Small driver for ncu: python scripts/prof_driver.py {mxv|mxm|bfs|sssp} [scale] -- builds the R-MAT input on the GPU and runs the
hot call a few times (first calls are warm-up; profile with -s / -k)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch
import bench
import graphblas_b200 as gb

what = sys.argv[1]
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 22
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
gb.init()
torch.cuda.set_stream(torch.cuda.Stream())
gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
if what == "mxm":
    ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2A, 42, device=dev)
    v = bench.values_torch(c.numel(), 43, torch.float32, device=dev)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    for _ in range(reps):
        C = None
        C = A.mxm(A, gb.semiring.plus_times).new()
    torch.cuda.synchronize()
    print("nvals", C.nvals)
else:
    ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
    if what == "mxv":
        v = bench.values_torch(c.numel(), 45, torch.float32, device=dev)
        A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
        x = gb.cuda.vector_from_torch(bench.values_torch(n, 46, torch.float32, device=dev))
        for _ in range(reps):
            y = A.mxv(x, gb.semiring.plus_times).new()
        torch.cuda.synchronize()
        print("nvals", y.nvals)
    elif what == "sssp":
        w = (bench.values_torch(c.numel(), 43, torch.float32, device=dev) * 255).to(torch.int64) + 1
        A = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
        gb.cuda.matrix_build_transpose(A)
        d = gb.Vector.from_coo([0], [0], size=n, dtype=gb.dtypes.INT64)
        for _ in range(reps):
            d(gb.binary.min) << d.vxm(A, gb.semiring.min_plus)
        torch.cuda.synchronize()
        print("nvals", d.nvals)
