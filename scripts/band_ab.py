"""A/B of the pull SpMV kernels incl. the column-banded one on the Graph500-skew R-MAT: kernel time (library profile mode),
whole-call time, exactness against the merge-path result.  python scripts/band_ab.py [scale]"""
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
PEAK = 6550.4
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
nnz = c.numel()
for name, tdt, sr in (("fp32 plus_times", torch.float32, gb.semiring.plus_times), ("int64 min_plus", torch.int64, gb.semiring.min_plus),
                      ("fp64 plus_second", torch.float64, gb.semiring.plus_second), ("fp32 plus_times sparse x", torch.float32, gb.semiring.plus_times)):
    v = (bench.values_torch(nnz, 45, torch.float32, device=dev) * 255 + 1).floor().to(tdt)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    xv = (bench.values_torch(n, 46, torch.float32, device=dev) * 100).floor().to(tdt)
    xp = (bench.values_torch(n, 47, torch.float32, device=dev) < 0.5).to(torch.uint8) if "sparse" in name else None
    x = gb.cuda.vector_from_torch(xv, xp)
    es = v.element_size()
    rho = 0 if "second" in name else 1
    algo_bytes = nnz * (4 + rho * es) + (n + 1) * 8 + n * es + n * (es + 1)
    ref = None
    for method in ("merge", "seg", "band"):
        gb.cuda.set_option("spmv", method)
        gb.cuda.set_option("spmv_hot", "0")
        for _ in range(3):
            y = A.mxv(x, sr).new()
        yv, yp = (t.clone() for t in gb.cuda.vector_as_torch(y, sync=True))
        if ref is None:
            ref = (yv, yp)
        same = bool(torch.equal(ref[1], yp)) and bool(torch.equal(ref[0][yp.bool()], yv[yp.bool()]))
        gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
        for _ in range(10):
            y = A.mxv(x, sr).new()
        kt = gb.cuda.kernel_times(reset=True); gb.cuda.set_option("profile", "0")
        main = [k for k in ("spmv_band", "spmv_seg", "spmv_merge") if k in kt][0]
        ms = kt[main][0] / kt[main][1]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            y = A.mxv(x, sr).new()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:26s} {method:5s} {main}={ms*1e3:7.1f} us ({algo_bytes/ms/1e6:7.1f} GB/s = {algo_bytes/ms/1e6/PEAK*100:4.1f}% of measured HBM)  call={e0.elapsed_time(e1)/20*1e3:7.1f} us  exact={same}  others={ {k: round(v[0]/v[1]*1e3,1) for k,v in kt.items() if k!=main} }", flush=True)
    gb.cuda.set_option("spmv", "auto"); gb.cuda.set_option("spmv_hot", "auto")
    for _ in range(12):
        y = A.mxv(x, sr).new()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = A.mxv(x, sr).new()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:26s} auto  call={e0.elapsed_time(e1)/20*1e3:7.1f} us", flush=True)
    del A, x, v
