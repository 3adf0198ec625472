"""merge vs seg pull SpMV on the calls the iterative workloads make (accumulator fused, transposed operand)."""
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
ip, c, n = bench.rmat_csr_torch(scale, bench.RMAT_2B, 42, device=dev)
nnz = c.numel()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timeit(fn, reps=20):
    for _ in range(12):   # past the library's kernel-selection trial
        fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

Af = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(nnz, dtype=torch.float64, device=dev), n, n)
gb.cuda.matrix_build_transpose(Af)
wv = gb.cuda.vector_from_torch(torch.rand(n, dtype=torch.float64, device=dev))
g = torch.Generator(device=dev); g.manual_seed(43)
W = gb.cuda.matrix_from_device_csr(ip, c, torch.randint(1, 256, (nnz,), device=dev, generator=g, dtype=torch.int64), n, n)
gb.cuda.matrix_build_transpose(W)
d = gb.cuda.vector_from_torch(torch.randint(0, 1000, (n,), device=dev, dtype=torch.int64))
A32 = gb.cuda.matrix_from_device_csr(ip, c, bench.values_torch(nnz, 45, torch.float32, device=dev), n, n)
x32 = gb.cuda.vector_from_torch(bench.values_torch(n, 46, torch.float32, device=dev))

def pr_plain():
    return Af.T.mxv(wv, gb.semiring.plus_second).new()
def pr_accum():
    r = gb.Vector(gb.dtypes.FP64, n); r[:] = 0.1
    r(gb.binary.plus) << Af.T.mxv(wv, gb.semiring.plus_second)
def pr_fill_only():
    r = gb.Vector(gb.dtypes.FP64, n); r[:] = 0.1
def sssp_accum():
    d(gb.binary.min) << d.vxm(W, gb.semiring.min_plus)
def sssp_plain():
    return d.vxm(W, gb.semiring.min_plus).new()
def mxv32():
    return A32.mxv(x32, gb.semiring.plus_times).new()

for name, fn in (("fp32 plus_times mxv", mxv32), ("pagerank mxv plain", pr_plain), ("pagerank fill only", pr_fill_only), ("pagerank fill+accum mxv", pr_accum),
                 ("sssp vxm plain", sssp_plain), ("sssp vxm min-accum", sssp_accum)):
    out = []
    for method in ("merge", "seg"):
        gb.cuda.set_option("spmv", method); gb.cuda.set_option("spmv_hot", "0")
        out.append(f"{method}={timeit(fn):7.1f} us")
    gb.cuda.set_option("spmv", "auto"); gb.cuda.set_option("spmv_hot", "auto")
    out.append(f"auto={timeit(fn):7.1f} us")
    gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True); fn()
    out.append("auto ran: " + ",".join(k for k in gb.cuda.kernel_times(reset=True) if k.startswith("spmv")))
    gb.cuda.set_option("profile", "0")
    print(f"{name:26s} " + "  ".join(out), flush=True)
gb.cuda.set_option("profile", "1"); gb.cuda.kernel_times(reset=True)
for _ in range(5):
    pr_accum(); sssp_accum()
print({k: round(v[0] / v[1] * 1e3, 1) for k, v in gb.cuda.kernel_times(reset=True).items()})
