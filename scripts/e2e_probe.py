"""Where does the blocked, overlapped export lose its time?  python scripts/e2e_probe.py"""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "python-graphblas_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import torch, bench
import graphblas_b200 as gb
gb.init()
torch.cuda.set_stream(torch.cuda.Stream()); gb.cuda.use_torch_stream()
dev = torch.device("cuda", 0)
ip, c, n = bench.rmat_csr_torch(22, bench.RMAT_2A, 42, device=dev)
v = bench.values_torch(c.numel(), 43, torch.float32, device=dev)
A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
sr = gb.semiring.plus_times
C = A.mxm(A, sr).new(); nv = C.nvals; C = None
o_ptr = torch.empty(n + 1, dtype=torch.int64).pin_memory()
o_col = torch.empty(nv, dtype=torch.int32).pin_memory()
o_val = torch.empty(nv, dtype=torch.float32).pin_memory()

def wall(fn, reps=2):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3 / reps

def single():
    Ch = A.mxm(A, sr).new()
    gb.cuda.matrix_export_host_csr32(Ch, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), sort=False)
print("single mxm + export            %7.1f ms" % wall(single), flush=True)
Ch = A.mxm(A, sr).new(); gb.cuda.matrix_compact(Ch)
print("export of a ready compact C    %7.1f ms" % wall(lambda: gb.cuda.matrix_export_host_csr32(Ch, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), sort=False)), flush=True)
cp, cj, cx = gb.cuda.matrix_as_torch(Ch)
def raw_copy(chunks):
    s = torch.cuda.Stream()
    per = (nv + chunks - 1) // chunks
    with torch.cuda.stream(s):
        for k in range(chunks):
            a, b = k * per, min(nv, (k + 1) * per)
            o_col[a:b].copy_(cj[a:b], non_blocking=True)
            o_val[a:b].copy_(cx[a:b], non_blocking=True)
    s.synchronize()
for ch in (1, 12, 48):
    print("raw torch D2H in %2d chunks      %7.1f ms" % (ch, wall(lambda: raw_copy(ch))), flush=True)
Ch = None; cp = cj = cx = None
for blocks in (3, 6, 12, 24):
    print("mxm_to_host_csr32 blocks=%2d     %7.1f ms" % (blocks, wall(lambda: gb.cuda.mxm_to_host_csr32(A, A, sr, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), blocks=blocks))), flush=True)
# compute side only: the same block loop without the export
def blocks_only(blocks):
    ipx, cjx, cxx = gb.cuda.matrix_as_torch(A, sync=False)
    nnz = int(ipx[-1])
    targets = torch.tensor([nnz * k // blocks for k in range(blocks + 1)], dtype=torch.int64, device=dev)
    bounds = torch.searchsorted(ipx, targets).tolist(); bounds[0], bounds[-1] = 0, n
    keep = []
    for q0, q1 in zip(bounds[:-1], bounds[1:]):
        k0, k1 = int(ipx[q0]), int(ipx[q1])
        Ab = gb.cuda.matrix_from_device_csr((ipx[q0:q1 + 1] - k0).contiguous(), cjx[k0:k1], cxx[k0:k1], q1 - q0, n)
        Cb = Ab.mxm(A, sr).new(); gb.cuda.matrix_compact(Cb); keep.append(Cb)
print("12 blocks, multiply + compact only %5.1f ms" % wall(lambda: blocks_only(12)), flush=True)
