"""Full-size (BASELINE.json scale-22) checks.  (1) Entry-by-entry comparison with the ORACLE on a random row sample of the
headline product (the bench's own fp32 inputs: pattern exact, values within the stated tolerance) and of the SSSP / PageRank
multiplies; (2) size-independent identities that any correct implementation must satisfy on the whole result:
  * mxv with all-ones operands reproduces the degree vector exactly; merge-path and warp-per-row kernels agree bit-exactly;
  * mxm: nvals equals the symbolic count, flops equals sum_k deg_A_col(k)*deg_B_row(k), and the checksum of checksums
    C.1 == A.(A.1) holds exactly for small-integer values;
  * BFS levels / SSSP distances satisfy the edge-relaxation optimality conditions on every edge.
torch is used only to generate inputs on the device and to evaluate the properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCALE = 22


@pytest.fixture(scope="module")
def gb():
    import graphblas_b200 as gb

    gb.init()
    return gb


@pytest.fixture(scope="module")
def torch():
    import torch

    return torch


def _rmat(torch, params, seed=42):
    import bench

    return bench.rmat_csr_torch(SCALE, params, seed, device=torch.device("cuda", 0))


def _vec_to_torch(gb, torch, v):
    vals, pres = gb.cuda.vector_as_torch(v)
    return vals.clone(), pres.clone()


def test_mxv_degree_identities(gb, torch):
    import bench

    ip, c, n = _rmat(torch, bench.RMAT_2B)
    deg = (ip[1:] - ip[:-1])
    ones = torch.ones(c.numel(), dtype=torch.float32, device=c.device)
    A = gb.cuda.matrix_from_device_csr(ip, c, ones, n, n)
    x = gb.cuda.vector_from_torch(torch.ones(n, dtype=torch.float32, device=c.device))
    results = {}
    for method in ("merge", "seg", "hot", "rowwarp", "band"):
        gb.cuda.set_option("spmv", "seg" if method == "hot" else method)
        gb.cuda.set_option("spmv_hot", "1" if method == "hot" else "0")   # hot-column cache of the pull kernel forced / off
        y = A.mxv(x, gb.semiring.plus_times).new()
        vals, pres = _vec_to_torch(gb, torch, y)
        assert torch.equal(pres.bool(), deg > 0)
        assert torch.equal(vals[deg > 0], deg[deg > 0].to(torch.float32))      # max degree 97k < 2^24: exact in fp32
        assert y.nvals == int((deg > 0).sum())
        results[method] = vals
        z = A.mxv(x, gb.semiring.any_pair).new()
        zv, zp = _vec_to_torch(gb, torch, z)
        assert torch.equal(zp.bool(), deg > 0) and bool((zv[deg > 0] == 1).all())
    gb.cuda.set_option("spmv", "auto")
    gb.cuda.set_option("spmv_hot", "auto")
    # int64 min_plus with x = 0: y(i) = min weight of row i; compare with a torch segmented min; both kernels bit-exact
    g = torch.Generator(device=c.device); g.manual_seed(7)
    w = torch.randint(1, 256, (c.numel(),), device=c.device, generator=g, dtype=torch.int64)
    W = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
    x0 = gb.cuda.vector_from_torch(torch.zeros(n, dtype=torch.int64, device=c.device))
    rows = torch.repeat_interleave(torch.arange(n, device=c.device), deg)
    want = torch.full((n,), 1 << 62, dtype=torch.int64, device=c.device).scatter_reduce(0, rows, w, "amin")
    for method in ("merge", "seg", "hot", "rowwarp", "band"):
        gb.cuda.set_option("spmv", "seg" if method == "hot" else method)
        gb.cuda.set_option("spmv_hot", "1" if method == "hot" else "0")
        y = W.mxv(x0, gb.semiring.min_plus).new()
        vals, pres = _vec_to_torch(gb, torch, y)
        assert torch.equal(vals[deg > 0], want[deg > 0]), method
    gb.cuda.set_option("spmv", "auto")
    gb.cuda.set_option("spmv_hot", "auto")
    # random x (fp64, exactly summable small integers): hot-cache kernel == plain merge kernel bit-exactly, incl. sparse x
    g2 = torch.Generator(device=c.device); g2.manual_seed(9)
    xr = torch.randint(0, 8, (n,), device=c.device, generator=g2).to(torch.float64)
    pr = (torch.rand(n, device=c.device, generator=g2) < 0.7).to(torch.uint8)
    Ad = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(c.numel(), dtype=torch.float64, device=c.device), n, n)
    for present in (None, pr):
        xv = gb.cuda.vector_from_torch(xr, present)
        outs = []
        for hot in ("0", "1"):
            gb.cuda.set_option("spmv_hot", hot)
            outs.append(_vec_to_torch(gb, torch, Ad.mxv(xv, gb.semiring.plus_times).new()))
        gb.cuda.set_option("spmv_hot", "auto")
        assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][0], outs[1][0])
    # "auto": the first multiplies with a CSR are the library's timed trial (merge / segmented / segmented + hot columns,
    # two runs each), then the winner is kept -- every one of them must return the same exact result
    for M, xv, sr in ((Ad, gb.cuda.vector_from_torch(xr, pr), gb.semiring.plus_times), (W, x0, gb.semiring.min_plus)):
        gb.cuda.set_option("spmv", "merge")
        ref_v, ref_p = (t.clone() for t in _vec_to_torch(gb, torch, M.mxv(xv, sr).new()))
        gb.cuda.set_option("spmv", "auto")
        for it in range(11):
            got_v, got_p = _vec_to_torch(gb, torch, M.mxv(xv, sr).new())
            assert torch.equal(got_p, ref_p) and torch.equal(got_v[ref_p.bool()], ref_v[ref_p.bool()]), it
    # pull over the cached transpose == push: vxm(x, A) column sums == in-degree
    indeg = torch.bincount(c.long(), minlength=n)
    for vm in ("pull", "push"):
        gb.cuda.set_option("vxm_method", vm)
        y = x.vxm(A, gb.semiring.plus_times).new()
        vals, pres = _vec_to_torch(gb, torch, y)
        assert torch.equal(pres.bool(), indeg > 0), vm
        assert torch.equal(vals[indeg > 0], indeg[indeg > 0].to(torch.float32)), vm
    gb.cuda.set_option("vxm_method", "auto")


def test_mxm_checksum_of_checksums(gb, torch):
    import bench

    ip, c, n = _rmat(torch, bench.RMAT_2A)
    g = torch.Generator(device=c.device); g.manual_seed(11)
    v = torch.randint(1, 3, (c.numel(),), device=c.device, generator=g, dtype=torch.int32)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    deg = (ip[1:] - ip[:-1])
    flops_want = int(deg[c.long()].sum())
    flops, nvals_sym = gb.cuda.mxm_symbolic(A, A)
    assert flops == flops_want
    C = A.mxm(A, gb.semiring.plus_times).new()
    assert C.nvals == nvals_sym                       # one-pass numeric count == independent symbolic (two-pass) count
    assert (C.nrows, C.ncols) == (n, n)
    ones = gb.cuda.vector_from_torch(torch.ones(n, dtype=torch.int32, device=c.device))
    lhs = C.mxv(ones, gb.semiring.plus_times).new()                                  # C.1 on the (unsorted) result
    rhs = A.mxv(A.mxv(ones, gb.semiring.plus_times).new(), gb.semiring.plus_times).new()   # A.(A.1)
    lv, lp = _vec_to_torch(gb, torch, lhs)
    rv, rp = _vec_to_torch(gb, torch, rhs)
    assert torch.equal(lp, rp) and torch.equal(lv[lp.bool()], rv[rp.bool()])
    # structure-only product agrees on the pattern size
    Cb = A.mxm(A, gb.semiring.any_pair).new()
    assert Cb.nvals == C.nvals


def test_bfs_and_sssp_optimality_conditions(gb, torch):
    import bench

    ip, c, n = _rmat(torch, bench.RMAT_2B)
    deg = (ip[1:] - ip[:-1])
    dev = c.device
    rows = torch.repeat_interleave(torch.arange(n, device=dev), deg)
    cols = c.long()
    src = int(torch.nonzero(deg > 0)[0])
    A = gb.cuda.matrix_from_device_csr(ip, c, torch.ones(c.numel(), dtype=torch.bool, device=dev), n, n)
    # level BFS exactly as the reference notebook writes it
    q = gb.Vector.from_coo([src], [True], size=n)
    lv = gb.Vector(gb.dtypes.INT64, n)
    for level in range(1, 200):
        lv(mask=q.V)[:] = level
        q(~lv.S, replace=True) << q.vxm(A, gb.semiring.any_pair)
        if q.nvals == 0:
            break
    L, P = _vec_to_torch(gb, torch, lv)
    P = P.bool()
    assert bool(P[src]) and int(L[src]) == 1 and int(P.sum()) > n // 4
    INF = 1 << 40
    Lf = torch.where(P, L, torch.full_like(L, INF))
    assert bool((Lf[cols] <= Lf[rows] + 1)[P[rows]].all())                 # no edge skips a level; reached set is closed
    best = torch.full((n,), INF, dtype=torch.int64, device=dev).scatter_reduce(0, cols, Lf[rows], "amin")
    reached = P.clone(); reached[src] = False
    assert torch.equal(best[reached] + 1, L[reached])                       # every reached vertex has a parent one level up
    # SSSP: Bellman-Ford sweeps w(min) << w.vxm(A, min_plus) to a fixed point
    g = torch.Generator(device=dev); g.manual_seed(43)
    w = torch.randint(1, 256, (c.numel(),), device=dev, generator=g, dtype=torch.int64)
    W = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
    d = gb.Vector.from_coo([src], [0], size=n, dtype=gb.dtypes.INT64)
    for it in range(200):
        old = d.dup()
        d(gb.binary.min) << d.vxm(W, gb.semiring.min_plus)
        if d.isequal(old):
            break
    D, DP = _vec_to_torch(gb, torch, d)
    DP = DP.bool()
    assert torch.equal(DP, P)                                                # same reachable set as BFS
    Df = torch.where(DP, D, torch.full_like(D, INF))
    assert bool((Df[cols] <= Df[rows] + w)[DP[rows]].all())                 # no edge can be relaxed further
    bestd = torch.full((n,), INF, dtype=torch.int64, device=dev).scatter_reduce(0, cols, Df[rows] + w, "amin")
    assert torch.equal(bestd[reached], D[reached]) and int(D[src]) == 0      # every distance is realised by an in-edge


def test_masked_mxm_graph500_scale22(gb, torch):
    """BASELINE config 2b: C<A.S> = A (+).(x) A on the Graph500-skew matrix (1.46e11 multiply-adds; the unmasked product
    would be ~0.5 TB).  Checks: result pattern is a subset of the mask, and on a slice of rows the in-hash masked path
    agrees exactly with forming the unmasked product and masking afterwards."""
    import bench

    ip, c, n = _rmat(torch, bench.RMAT_2B)
    dev = c.device
    g = torch.Generator(device=dev); g.manual_seed(5)
    v = torch.randint(1, 4, (c.numel(),), device=dev, generator=g, dtype=torch.int32)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    C = A.mxm(A, gb.semiring.plus_times).new(mask=A.S)
    assert 0 < C.nvals <= A.nvals
    # slice: 20000 low-degree rows of A against the full A, both paths
    r0, r1 = n // 2, n // 2 + 20000
    k0, k1 = int(ip[r0]), int(ip[r1])
    As = gb.cuda.matrix_from_device_csr((ip[r0:r1 + 1] - k0).contiguous(), c[k0:k1].contiguous(), v[k0:k1].contiguous(), r1 - r0, n)
    fast = As.mxm(A, gb.semiring.plus_times).new(mask=As.S)
    gb.cuda.set_option("spgemm_mask", "0")
    try:
        slow = As.mxm(A, gb.semiring.plus_times).new(mask=As.S)
    finally:
        gb.cuda.set_option("spgemm_mask", "1")
    assert fast.isequal(slow)
    # the full result restricted to those rows equals the slice result
    Cp, Cj, Cx = C.to_csr()
    Fp, Fj, Fx = fast.to_csr()
    a, b = int(Cp[r0]), int(Cp[r1])
    assert np.array_equal(Cp[r0:r1 + 1] - Cp[r0], Fp) and np.array_equal(Cj[a:b], Fj) and np.array_equal(Cx[a:b], Fx)


def test_lazy_sort_all_row_classes(gb, torch):
    """The on-demand row sort (what finishes a 'jumbled' mxm result): rows of every size class -- warp (<= 32), shared-memory
    bitonic (<= 512, <= 8192) and the batched segmented sort for longer rows -- against torch.sort, values carried along."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(21)
    lens = torch.tensor([0, 1, 7, 32, 33, 500, 513, 4000, 8192, 8193, 20000, 70000, 3, 0, 100000], device=dev)
    ncols = 1 << 20
    ip = torch.zeros(lens.numel() + 1, dtype=torch.int64, device=dev); ip[1:] = torch.cumsum(lens, 0)
    cols = torch.cat([torch.randperm(ncols, device=dev, generator=g)[: int(l)] for l in lens]).to(torch.int32)
    for dt in (torch.int8, torch.float32, torch.int64):
        vals = torch.arange(cols.numel(), device=dev).to(dt)
        A = gb.cuda.matrix_from_device_csr(ip, cols, vals, lens.numel(), ncols, sorted=False)
        gb.cuda.matrix_sort(A)
        p2, c2, v2 = gb.cuda.matrix_as_torch(A)
        assert torch.equal(p2, ip)
        for r in range(lens.numel()):
            b, e = int(ip[r]), int(ip[r + 1])
            want_c, order = torch.sort(cols[b:e].long())
            assert torch.equal(c2[b:e].long(), want_c), (dt, r)
            assert torch.equal(v2[b:e], vals[b:e][order]), (dt, r)


# ------------------------------------------------------------------ oracle comparison on row samples of BASELINE's own configs
def _host_csr(ip, c, v):
    return ip.cpu().numpy(), c.cpu().numpy().astype(np.int64), v.cpu().numpy()


def _sample_rows(hp, hc, hv, rows, ncols):
    from oracle import bigref as R

    rows = np.sort(rows)
    lens = hp[rows + 1] - hp[rows]
    ptr = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    idx = np.repeat(hp[rows] - ptr[:-1], lens) + np.arange(ptr[-1])
    return rows, R.BigMat(ptr, hc[idx], hv[idx], rows.size, ncols)


def test_mxm_headline_row_sample_vs_oracle(gb, torch):
    """BASELINE configs[1] as bench.py runs it (R-MAT 2a, fp32 values U(0,1), plus_times): 50 000 random rows of C = A.A compared
    with the oracle's sorted Gustavson product of those rows (oracle/grb_oracle.c, pinned to the reference's goldens).
    Pattern: exact.  Values: |got - want| <= 1e-5 * sqrt(k) * sum|products| bound, stated as rtol = 1e-5 * sqrt(max row degree)
    against the oracle's fp32 sum in ascending-k order (the hash accumulates in arbitrary order; the reference's own tolerance
    for such comparisons is isclose's rel_tol 1e-7 on exactly representable sums, graphblas/core/matrix.py:417)."""
    import bench
    from oracle import bigref as R

    ip, c, n = _rmat(torch, bench.RMAT_2A)
    v = bench.values_torch(c.numel(), 43, torch.float32, device=c.device)
    A = gb.cuda.matrix_from_device_csr(ip, c, v, n, n)
    C = A.mxm(A, gb.semiring.plus_times).new()
    gb.cuda.matrix_sort(C)
    cp, cj, cx = gb.cuda.matrix_as_torch(C)
    hp, hc, hv = _host_csr(ip, c, v)
    rng = np.random.default_rng(3)
    rows, As = _sample_rows(hp, hc, hv, rng.choice(n, size=50_000, replace=False), n)
    want = R.mxm_T("plus_times", As, R.BigMat(hp, hc, hv, n, n))
    rt = torch.as_tensor(rows, device=c.device)
    starts, ends = cp[rt], cp[rt + 1]
    lens = (ends - starts).cpu().numpy()
    assert np.array_equal(lens, np.diff(want.indptr)), "row lengths differ from the oracle"
    gather = torch.repeat_interleave(starts - torch.as_tensor(want.indptr[:-1], device=c.device), ends - starts) + torch.arange(int(want.indptr[-1]), device=c.device)
    got_j = cj[gather].cpu().numpy().astype(np.int64)
    got_x = cx[gather].cpu().numpy()
    assert np.array_equal(got_j, want.indices), "column pattern differs from the oracle"
    kmax = int((hp[1:] - hp[:-1]).max())
    rtol = 1e-5 * np.sqrt(kmax)
    err = np.abs(got_x.astype(np.float64) - want.values.astype(np.float64))
    assert np.all(err <= rtol * np.abs(want.values.astype(np.float64))), float((err / np.abs(want.values)).max())
    # the same rows through the opt-in tile kernel (spgemm_tile.cuh)
    nvals_staged = C.nvals
    del cp, cj, cx, C
    gb.cuda.set_option("trim", "1")       # hand the library's cached 20 GB blocks back before the second product
    gb.cuda.set_option("spgemm_tile", "1")
    try:
        C2 = A.mxm(A, gb.semiring.plus_times).new()
        assert C2.nvals == nvals_staged
        gb.cuda.matrix_sort(C2)
        p2, j2, x2 = gb.cuda.matrix_as_torch(C2)
        starts2, ends2 = p2[rt], p2[rt + 1]
        assert np.array_equal((ends2 - starts2).cpu().numpy(), np.diff(want.indptr))
        gather2 = torch.repeat_interleave(starts2 - torch.as_tensor(want.indptr[:-1], device=c.device), ends2 - starts2) + torch.arange(int(want.indptr[-1]), device=c.device)
        assert np.array_equal(j2[gather2].cpu().numpy().astype(np.int64), want.indices)
        err2 = np.abs(x2[gather2].cpu().numpy().astype(np.float64) - want.values.astype(np.float64))
        assert np.all(err2 <= rtol * np.abs(want.values.astype(np.float64)))
    finally:
        gb.cuda.set_option("spgemm_tile", None)
        gb.cuda.set_option("trim", "1")


def test_sssp_and_pagerank_multiplies_row_sample_vs_oracle(gb, torch):
    """configs 4 / 5 multiplies on the Graph500-skew matrix: int64 min_plus mxv (exact) and fp64 plus_second A'.mxv (rtol 1e-12:
    fp64 sums of <= 1e5 positive terms) on 200 000 random output rows, against the oracle's row-dot mxv"""
    import bench
    from oracle import bigref as R

    ip, c, n = _rmat(torch, bench.RMAT_2B)
    g = torch.Generator(device=c.device); g.manual_seed(43)
    w = torch.randint(1, 256, (c.numel(),), device=c.device, generator=g, dtype=torch.int64)
    W = gb.cuda.matrix_from_device_csr(ip, c, w, n, n)
    xi = torch.randint(0, 1000, (n,), device=c.device, generator=g, dtype=torch.int64)
    pres = (torch.rand(n, device=c.device, generator=g) < 0.6).to(torch.uint8)
    x = gb.cuda.vector_from_torch(xi, pres)
    y = W.mxv(x, gb.semiring.min_plus).new()
    yv, yp = _vec_to_torch(gb, torch, y)
    hp, hc, hw = _host_csr(ip, c, w)
    rng = np.random.default_rng(5)
    rows, Ws = _sample_rows(hp, hc, hw, rng.choice(n, size=200_000, replace=False), n)
    want = R.mxv_T("min_plus", Ws, R.BigVec(xi.cpu().numpy(), pres.cpu().numpy()))
    rt = torch.as_tensor(rows, device=c.device)
    assert np.array_equal(yp[rt].cpu().numpy(), want.present)
    sel = want.present.astype(bool)
    assert np.array_equal(yv[rt].cpu().numpy()[sel], want.vals[sel])
    # PageRank multiply: r = A'.w with plus_second on fp64; columns of A = rows of the transposed CSR
    ones = torch.ones(c.numel(), dtype=torch.float64, device=c.device)
    Af = gb.cuda.matrix_from_device_csr(ip, c, ones, n, n)
    wv = torch.rand(n, device=c.device, generator=g, dtype=torch.float64)
    r = Af.T.mxv(gb.cuda.vector_from_torch(wv), gb.semiring.plus_second).new()
    rv, rp = _vec_to_torch(gb, torch, r)
    from graphblas_b200 import distributed as D

    tp, tc, _ = D.transpose_csr_torch(ip, c, None, n)
    htp, htc = tp.cpu().numpy(), tc.cpu().numpy().astype(np.int64)
    rows, Ts = _sample_rows(htp, htc, np.ones(htc.size, dtype=np.float64), rng.choice(n, size=200_000, replace=False), n)
    want = R.mxv_T("plus_second", Ts, R.BigVec(wv.cpu().numpy(), np.ones(n, np.uint8)))
    rt = torch.as_tensor(rows, device=c.device)
    assert np.array_equal(rp[rt].cpu().numpy(), want.present)
    sel = want.present.astype(bool)
    assert np.allclose(rv[rt].cpu().numpy()[sel], want.vals[sel], rtol=1e-12, atol=0)
