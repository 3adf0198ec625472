"""Debug: in-hash masked mxm vs unmasked-then-mask on G500-skew R-MAT at several scales; prints rows whose counts differ."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import numpy as np, torch
import graphblas_b200 as gb
import bench
gb.init()
dev = torch.device("cuda", 0)
for scale in [int(s) for s in sys.argv[1:]] or [14, 16, 18]:
    ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
    g = torch.Generator(device=dev); g.manual_seed(5)
    v = torch.randint(1, 4, (c.numel(),), device=dev, generator=g, dtype=torch.int32)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    fast = A.mxm(A, gb.semiring.plus_times).new(mask=A.S)
    print(f"scale {scale}: nnz(A)={A.nvals} fast nvals={fast.nvals}", flush=True)
    if scale <= 18:
        gb.cuda.set_option("spgemm_mask", "0")
        slow = A.mxm(A, gb.semiring.plus_times).new(mask=A.S)
        gb.cuda.set_option("spgemm_mask", "1")
        print(f"   slow nvals={slow.nvals} isequal={fast.isequal(slow)}", flush=True)
        Fp, Fj, Fx = fast.to_csr(); Sp, Sj, Sx = slow.to_csr()
        fc, sc = np.diff(Fp.astype(np.int64)), np.diff(Sp.astype(np.int64))
        bad = np.flatnonzero(fc != sc)
        deg = (ip[1:] - ip[:-1]).cpu().numpy()
        print("   rows with differing counts:", bad.size, "first:", [(int(r), int(deg[r]), int(fc[r]), int(sc[r])) for r in bad[:12]])
        if bad.size == 0 and not np.array_equal(Fx, Sx):
            print("   values differ at", np.flatnonzero(Fx != Sx)[:10])
    else:
        Fp, Fj, Fx = fast.to_csr()
        fc = np.diff(Fp.astype(np.int64)); deg = (ip[1:] - ip[:-1]).cpu().numpy()
        bad = np.flatnonzero(fc > deg)
        print("   rows with count > deg:", bad.size, [(int(r), int(deg[r]), int(fc[r])) for r in bad[:12]])
