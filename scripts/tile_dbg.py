"""Phase counters of the tiled SpGEMM kernel (option spgemm_tile_dbg) for a few configurations.  python scripts/tile_dbg.py [scale]"""
import sys, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2A, 42, device=dev)
v = bench.values_torch(c.numel(), 43, torch.float32, device=dev)
A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
sr = gb.semiring.plus_times
SETS = [json.loads(a) for a in sys.argv[2:]] or [{}]
for opts in SETS:
    for k, val in opts.items():
        gb.cuda.set_option(k, val)
    C = A.mxm(A, sr).new(); C = None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    C = A.mxm(A, sr).new()
    e1.record(); torch.cuda.synchronize()
    print(f"{json.dumps(opts):80s} call={e0.elapsed_time(e1):9.2f} ms nvals={C.nvals}", flush=True)
    gb.cuda.set_option("spgemm_tile_dbg", "1")
    C = None
    C = A.mxm(A, sr).new(); C = None
    gb.cuda.set_option("spgemm_tile_dbg", None)
    for k in opts:
        gb.cuda.set_option(k, None)
