import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
dev = torch.device("cuda", 0)
for scale in (16, 18, 20, 22):
    ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
    deg = ip[1:] - ip[:-1]
    A = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(c.numel(), dtype=torch.float32, device=dev), n, n)
    x = gb.cuda.vector_from_torch(torch.ones(n, dtype=torch.float32, device=dev))
    for method in ("rowwarp", "merge"):
        gb.cuda.set_option("spmv", method)
        y = A.mxv(x, gb.semiring.plus_times).new()
        vals, pres = gb.cuda.vector_as_torch(y)
        vals, pres = vals.clone(), pres.clone().bool()
        badp = torch.nonzero(pres != (deg > 0)).flatten()
        badv = torch.nonzero((vals != deg.float()) & (deg > 0)).flatten()
        print(f"scale {scale} {method}: n={n} nnz={c.numel()} presence mismatches={badp.numel()} value mismatches={badv.numel()} nvals={y.nvals}")
        for r in badv[:6].tolist() + badp[:4].tolist():
            print("   row", r, "deg", int(deg[r]), "ptr", int(ip[r]), "got", float(vals[r]), "pres", bool(pres[r]), "tile(1792)", (int(ip[r]) + r) // 1792, "->", (int(ip[r + 1]) + r + 1) // 1792)
