"""Exact comparison of the tiled one-pass SpGEMM (spgemm_tile.cuh) against the staged binned path on R-MAT inputs:
row pointers, sorted columns and values must be identical (integer-valued fp32 / int64 inputs, so sums are exact).
python scripts/tile_check.py [scale ...]"""
import sys, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb

gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
scales = [int(a) for a in sys.argv[1:]] or [10, 14, 17]
ok_all = True
for scale in scales:
    for params, name in ((bench.RMAT_2A, "2a"), (bench.RMAT_2B, "2b")):
        if name == "2b" and scale > 16:
            continue
        ip, c, n = bench.rmat_csr_torch(scale, params, 42, device=dev)
        for dt, sr in ((torch.float32, gb.semiring.plus_times), (torch.int64, gb.semiring.min_plus), (torch.float64, gb.semiring.plus_second)):
            g = torch.Generator(device=dev); g.manual_seed(7)
            v = torch.randint(1, 5, (c.numel(),), device=dev, generator=g).to(dt)
            A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
            outs = []
            for opts in ({"spgemm_tile": "0"}, {"spgemm_tile": "1"}, {"spgemm_tile": "1", "spgemm_tile_ctas": "1", "spgemm_tile_threads": "512"},
                         {"spgemm_tile": "1", "spgemm_tile_scap": "512", "spgemm_tile_tcap": "2048"}):
                for k, val in opts.items():
                    gb.cuda.set_option(k, val)
                C = A.mxm(A, sr).new()
                gb.cuda.matrix_sort(C)
                p_, j_, x_ = gb.cuda.matrix_as_torch(C)
                outs.append((p_.clone(), j_.clone(), x_.clone(), C.nvals))
                for k in opts:
                    gb.cuda.set_option(k, None)
                del C
            ref = outs[0]
            for i, o in enumerate(outs[1:], 1):
                same = o[3] == ref[3] and torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]) and torch.equal(o[2], ref[2])
                ok_all &= bool(same)
                print(f"scale {scale} {name} {str(dt):14s} variant {i}: nvals {o[3]} vs {ref[3]}  identical={bool(same)}", flush=True)
print("TILE_CHECK", "OK" if ok_all else "FAILED")
sys.exit(0 if ok_all else 1)
