"""A/B of the pull SpMV: plain merge kernel vs hot-column cache, per type/semiring, on the Graph500-skew R-MAT
(natural labels and randomly permuted labels) -- kernel time (library profile mode) and whole-call time."""
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
PEAK = 6550.4
ip0, c0, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
for labels in (("natural", "permuted") if len(sys.argv) < 3 else (sys.argv[2],)):
    if labels == "permuted":
        g = torch.Generator(device=dev); g.manual_seed(3)
        perm = torch.randperm(n, device=dev, generator=g)
        deg = ip0[1:] - ip0[:-1]
        rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
        key = torch.sort(perm[rows] * n + perm[c0.long()]).values
        rows2, c = key // n, (key % n).to(torch.int32)
        ip = torch.zeros(n + 1, dtype=torch.int64, device=dev); ip[1:] = torch.cumsum(torch.bincount(rows2, minlength=n), 0)
        del perm, rows, key, rows2
    else:
        ip, c = ip0, c0
    nnz = c.numel()
    for name, tdt, sr in (("fp32 plus_times", torch.float32, gb.semiring.plus_times), ("fp64 plus_second", torch.float64, gb.semiring.plus_second),
                          ("int64 min_plus", torch.int64, gb.semiring.min_plus)):
        v = (bench.values_torch(nnz, 45, torch.float32, device=dev) * 255 + 1).to(tdt)
        A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
        x = gb.cuda.vector_from_torch((bench.values_torch(n, 46, torch.float32, device=dev) * 100).to(tdt))
        es = v.element_size()
        rho = 0 if "second" in name else 1
        algo_bytes = nnz * (4 + rho * es) + (n + 1) * 8 + n * es + n * (es + 1)
        ref = None
        for method, hot, kb in (("merge", "0", 132), ("seg", "0", 132), ("seg", "1", 100), ("seg", "1", 132), ("seg", "1", 164), ("seg", "1", 196), ("seg", "1", 227)):
            gb.cuda.set_option("spmv", method)
            gb.cuda.set_option("spmv_hot", hot)
            gb.cuda.set_option("spmv_hot_kb", kb)
            for _ in range(3):
                y = A.mxv(x, sr).new()
            yv = gb.cuda.vector_as_torch(y, sync=True)[0].clone()
            if ref is None:
                ref = yv
            same = bool(torch.equal(ref, yv)) if tdt != torch.float32 else bool(torch.allclose(ref, yv, rtol=1e-4))
            gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
            for _ in range(10):
                y = A.mxv(x, sr).new()
            kt = gb.cuda.kernel_times(reset=True); gb.cuda.set_option("profile", "0")
            main = [k for k in ("spmv_seg_hot", "spmv_seg", "spmv_merge") if k in kt][0]
            ms = kt[main][0] / kt[main][1]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                y = A.mxv(x, sr).new()
            e1.record(); torch.cuda.synchronize()
            print(f"{labels:9s} {name:18s} {method:5s} hot={hot} kb={kb:3d} {main}={ms*1e3:7.1f} us ({algo_bytes/ms/1e6:7.1f} GB/s, {algo_bytes/ms/1e6/PEAK*100:4.1f}% of measured HBM)  call={e0.elapsed_time(e1)/20*1e3:7.1f} us  same={same}  others={ {k: round(v[0]/v[1]*1e3,1) for k,v in kt.items() if k!=main} }", flush=True)
        gb.cuda.set_option("spmv_hot", "auto")
        gb.cuda.set_option("spmv", "auto")
        del A, x, v
