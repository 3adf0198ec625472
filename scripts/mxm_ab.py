"""A/B of SpGEMM tuning options on the bench workload (R-MAT 2a, plus_times fp32): whole-call device time + per-kernel
times (library profile mode) per option set.  python scripts/mxm_ab.py [scale]"""
import sys, pathlib, json
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2A, 42, device=dev)
v = bench.values_torch(c.numel(), 43, torch.float32, device=dev)
A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
sr = gb.semiring.plus_times
SETS = [
    {},
    {"spgemm_batch_cas": "0"},
    {},
    {"spgemm_batch_cas": "0"},
]
if len(sys.argv) > 2:
    SETS = [json.loads(a) for a in sys.argv[2:]]
ref_nvals = None
for opts in SETS:
    for k, val in opts.items():
        gb.cuda.set_option(k, val)
    C = None
    for _ in range(2):
        C = None
        C = A.mxm(A, sr).new()
    nv = C.nvals
    ref_nvals = ref_nvals or nv
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        C = None
        C = A.mxm(A, sr).new()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
    C = None
    C = A.mxm(A, sr).new()
    kt = gb.cuda.kernel_times(reset=True); gb.cuda.set_option("profile", "0")
    gb.cuda.set_option("trace_host", "1"); gb.cuda.kernel_times(reset=True)
    C = None
    torch.cuda.synchronize()
    C = A.mxm(A, sr).new()
    torch.cuda.synchronize()
    ht = gb.cuda.kernel_times(reset=True); gb.cuda.set_option("trace_host", "0")
    C = None
    print(f"{json.dumps(opts):90s} call={ms:7.2f} ms  nvals_ok={nv == ref_nvals}  kernels={ {k: round(t[0], 2) for k, t in kt.items() if t[0] > 0.3} }", flush=True)
    print("      host:", {k[5:]: round(t[0], 2) for k, t in ht.items() if k.startswith("host:")}, flush=True)
    for k in opts:
        gb.cuda.set_option(k, None)
