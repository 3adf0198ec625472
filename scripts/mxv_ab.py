import sys, pathlib, os
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
for name, tdt, sr in (("fp32 plus_times", torch.float32, gb.semiring.plus_times), ("fp64 plus_second", torch.float64, gb.semiring.plus_second),
                      ("int64 min_plus", torch.int64, gb.semiring.min_plus)):
    v = (bench.values_torch(nnz, 45, torch.float32, device=dev) * 255 + 1).to(tdt)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    x = gb.cuda.vector_from_torch((bench.values_torch(n, 46, torch.float32, device=dev) * 100).to(tdt))
    es = v.element_size()
    rho = 0 if "second" in name else 1
    algo_bytes = nnz * (4 + rho * es) + (n + 1) * 8 + n * es + n * (es + 1)
    for tex in ("0", "1"):
        gb.cuda.set_option("spmv_tex", tex)
        for _ in range(3):
            y = A.mxv(x, sr).new()
        gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
        for _ in range(10):
            y = A.mxv(x, sr).new()
        kt = gb.cuda.kernel_times(reset=True); gb.cuda.set_option("profile", "0")
        ms = kt["spmv_merge"][0] / kt["spmv_merge"][1]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = A.mxv(x, sr).new()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:18s} tex={tex} merge_kernel={ms*1e3:7.1f} us  ({algo_bytes/ms/1e6:7.1f} GB/s, {algo_bytes/ms/1e6/6579*100:4.1f}% of measured HBM peak)  whole call={e0.elapsed_time(e1)/20*1e3:7.1f} us  other kernels={ {k: round(v[0]/v[1]*1e3,1) for k,v in kt.items() if k!='spmv_merge'} }")
    del A, x, v
