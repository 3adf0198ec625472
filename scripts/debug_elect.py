"""elect vs atomic SpGEMM kernels: exact comparison (int32 plus_times, int64 any_pair) on R-MAT 2a, first differing row."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
for scale in [int(a) for a in sys.argv[1:]] or [16, 20]:
    ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2A, 42, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(11)
    v = torch.randint(1, 3, (c.numel(),), device=dev, generator=g, dtype=torch.int32)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    deg = ip[1:] - ip[:-1]
    rows_of = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    flops = torch.zeros(n, dtype=torch.int64, device=dev).index_add_(0, rows_of, deg[c.long()])
    for srname in ("plus_times", "any_pair"):
        sr = getattr(gb.semiring, srname)
        res = {}
        for elect in ("0", "1"):  # group kernel off / on
            gb.cuda.set_option("spgemm_group", elect)
            C = A.mxm(A, sr).new()
            gb.cuda.matrix_sort(C)
            p_, j_, x_ = gb.cuda.matrix_as_torch(C) if hasattr(gb.cuda, "matrix_as_torch") else (None, None, None)
            if p_ is None:
                I, J, X = C.to_coo()
                res[elect] = (torch.as_tensor(I.astype("int64")), torch.as_tensor(J.astype("int64")), torch.as_tensor(X))
            else:
                res[elect] = (p_.clone(), j_.clone(), x_.clone())
            del C
        a, b = res["0"], res["1"]
        same = all(x.shape == y.shape and bool(torch.equal(x, y)) for x, y in zip(a, b))
        print(f"scale {scale} {srname}: nvals {a[1].numel()} vs {b[1].numel()} identical={same}", flush=True)
        if not same and a[1].numel() == b[1].numel():
            if hasattr(gb.cuda, "matrix_as_torch"):
                bad = torch.nonzero((a[1] != b[1]) | (a[2] != b[2])).flatten()
                k = int(bad[0]); row = int(torch.searchsorted(a[0], torch.tensor([k], device=a[0].device), right=True)[0]) - 1
                print("  first diff at entry", k, "row", row, "row flops", int(flops[row]), "deg", int(deg[row]), "n bad", bad.numel())
                s, e = int(a[0][row]), int(a[0][row + 1])
                db = torch.nonzero((a[1][s:e] != b[1][s:e]) | (a[2][s:e] != b[2][s:e])).flatten()[:8]
                print("  cols a", a[1][s:e][db].tolist(), "b", b[1][s:e][db].tolist(), "vals a", a[2][s:e][db].tolist(), "b", b[2][s:e][db].tolist())
                badrows = torch.unique(torch.searchsorted(a[0], bad, right=True) - 1)
                print("  bad rows", badrows.numel(), "flops of bad rows: min", int(flops[badrows].min()), "max", int(flops[badrows].max()), "sample", flops[badrows][:12].tolist())
    gb.cuda.set_option("spgemm_group", None)
