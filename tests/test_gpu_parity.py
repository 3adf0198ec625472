"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C-ABI via the host
mirror, against (1) the reference's golden vectors, (2) the oracle on seeded random inputs, for every
traversal (merge-path pull, warp-per-row pull, push) and the hash SpGEMM bins.
Integer / boolean results and all index patterns are compared bit-exactly; floating point uses inputs whose
sums are exact (small integers) so those are bit-exact too, plus one tolerance test (rtol stated there)."""
import numpy as np
import pytest

import golden_cases as G
import helpers as H
from oracle import bigref as R
from oracle import semantics as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb():
    import graphblas_b200 as gb

    gb.init()
    return gb


def _obj(gb, ref):
    d = G.load(ref)
    if d["kind"] == "Matrix":
        return gb.Matrix.from_coo(d["rows"], d["cols"], d["vals"], nrows=d["nrows"], ncols=d["ncols"])
    return gb.Vector.from_coo(d["idx"], d["vals"], size=d["size"])


def _mask(gb, c):
    if not c["mask"]:
        return None
    m = _obj(gb, c["mask"])
    m = m.S if c["mask_kind"] == "S" else m.V
    return ~m if c["complement"] else m


def _expected(gb, ref):
    return _obj(gb, ref)


def _assert_equals_golden(got, ref):
    """numpy comparison of the extracted tuples with the reference's expected object (not the device-side isequal, which is
    itself made of kernels under test)"""
    d = G.load(ref)
    if d["kind"] == "Matrix":
        I, J, X = got.to_coo()
        order = np.lexsort((d["cols"], d["rows"]))
        assert (got.nrows, got.ncols) == (d["nrows"], d["ncols"])
        assert np.array_equal(I.astype(np.int64), d["rows"][order]) and np.array_equal(J.astype(np.int64), d["cols"][order]), (I, J)
        assert np.array_equal(X, d["vals"][order].astype(X.dtype)), (X, d["vals"][order])
    else:
        I, X = got.to_coo()
        order = np.argsort(d["idx"])
        assert got.size == d["size"]
        assert np.array_equal(I.astype(np.int64), d["idx"][order]), (I, d["idx"][order])
        assert np.array_equal(X, d["vals"][order].astype(X.dtype)), (X, d["vals"][order])


def _run_case(gb, c):
    a, b = _obj(gb, c["a"]), _obj(gb, c["b"])
    sr = getattr(gb.semiring, c["semiring"])
    if c["kind"] == "mxm":
        expr = (a.T if c["ta"] else a).mxm(b.T if c["tb"] else b, sr)
    elif c["kind"] == "mxv":
        expr = (a.T if c["ta"] else a).mxv(b, sr)
    else:
        expr = a.vxm(b.T if c["tb"] else b, sr)
    mask = _mask(gb, c)
    if c["out"] is None:
        return expr.new(mask=mask) if mask is not None else expr.new()
    out = _obj(gb, c["out"])
    kwargs = {}
    if mask is not None:
        kwargs["mask"] = mask
    if c["accum"]:
        kwargs["accum"] = getattr(gb.binary, c["accum"])
    if c["replace"]:
        kwargs["replace"] = True
    out(**kwargs) << expr
    return out


@pytest.mark.parametrize("method", ["auto", "merge", "rowwarp"])
@pytest.mark.parametrize("vxm_method", ["auto", "push", "pull"])
@pytest.mark.parametrize("c", G.CASES, ids=[c["id"] for c in G.CASES])
def test_reference_goldens(gb, c, method, vxm_method):
    if c["kind"] == "mxm" and (method != "auto" or vxm_method != "auto"):
        pytest.skip("spmv options do not affect mxm")
    gb.cuda.set_option("spmv", method)
    gb.cuda.set_option("vxm_method", vxm_method)
    try:
        got = _run_case(gb, c)
        _assert_equals_golden(got, c["expect"])
        assert got.isequal(_expected(gb, c["expect"]))   # the device-side predicate agrees
    finally:
        gb.cuda.set_option("spmv", "auto")
        gb.cuda.set_option("vxm_method", "auto")


def test_nonsquare_max_plus(gb):
    # reference graphblas/tests/test_matrix.py:335-345
    a, b = _obj(gb, G.NONSQUARE["a"]), _obj(gb, G.NONSQUARE["b"])
    C = gb.Matrix(a.dtype, nrows=1, ncols=1)
    C << a.mxm(b, gb.semiring.max_plus)
    assert C[0, 0].new() == 33
    C1 = a.mxm(b, gb.semiring.max_plus).new()
    assert C1.isequal(C)
    C2 = a.T.mxm(b.T, gb.semiring.max_plus).new()
    assert (C2.nrows, C2.ncols) == (5, 5)


def test_dimension_mismatch_raises_now(gb):
    # reference tests/test_matrix.py:1918-1925, tests/test_vector.py:1129-1139
    A = gb.Matrix.from_coo([0, 1], [0, 1], [1, 2], nrows=2, ncols=3)
    v = gb.Vector.from_coo([0], [1], size=2)
    with pytest.raises(gb.exceptions.DimensionMismatch):
        A.mxv(v)
    with pytest.raises(gb.exceptions.DimensionMismatch):
        A.mxm(A)
    with pytest.raises(gb.exceptions.DimensionMismatch):
        gb.Vector.from_coo([0], [1], size=5).vxm(A)


def test_mask_type_error_and_bad_option(gb):
    A = _obj(gb, G.A_M)
    struct_mask = gb.Matrix.from_coo([0, 3, 4], [2, 3, 2], [1, 0, 0], nrows=7, ncols=7)
    with pytest.raises(TypeError, match="Mask must be"):
        A.mxm(A).new(mask=struct_mask)     # reference tests/test_matrix.py:373-374
    with pytest.raises(ValueError, match="Extra descriptor options"):
        A.mxm(A).new(nthreads=4)           # reference tests/test_matrix.py:4329-4331 (non-suitesparse backend)


def test_recorder_call_text(gb):
    # reference tests/test_recorder.py:31-37 pins the shape of the C call
    A = gb.Matrix.from_coo([0, 1], [0, 1], [1, 2], nrows=2, ncols=2, name="A")
    B = gb.Matrix.from_coo([0, 1], [1, 0], [3, 4], nrows=2, ncols=2, name="B")
    with gb.Recorder() as rec:
        C = A.mxm(B, gb.semiring.plus_times).new(name="C")
        D = A.mxm(B.T, gb.semiring.min_plus).new(name="D")
    assert "GrB_mxm(C, NULL, NULL, GrB_PLUS_TIMES_SEMIRING_INT64, A, B, NULL);" in rec.data
    assert "GrB_mxm(D, NULL, NULL, GrB_MIN_PLUS_SEMIRING_INT64, A, B.T, GrB_DESC_T1);".replace("B.T", "B") in \
        [s.replace("B.T", "B") for s in rec.data]
    assert C.nvals == 2 and D.nvals == 2


def test_int64_power_wraps(gb):
    # reference tests/test_matrix.py:4379-4405 (test_power): chained INT64 mxm wraps, min_plus powers
    d = G.load(G.A_M)
    A = _obj(gb, G.A_M)
    Ab = R.BigMat.from_coo(d["rows"], d["cols"], d["vals"], 7, 7)
    P, Pb = A.dup(), Ab
    for i in range(48):
        P = P.mxm(A, gb.semiring.plus_times).new()
        Pb = R.mxm_T("plus_times", Pb, Ab)
        ok, msg = H.mat_equal(P, Pb)
        assert ok, (i, msg)
    P, Pb = A.dup(), Ab
    for i in range(9):
        P = P.mxm(A, gb.semiring.min_plus).new()
        Pb = R.mxm_T("min_plus", Pb, Ab)
        ok, msg = H.mat_equal(P, Pb)
        assert ok, (i, msg)


def test_aliasing(gb):
    # output aliases inputs and mask: reference tests/test_matrix.py:377-386 and the BFS idiom q(~v.S, replace) << q.vxm(A)
    rng = np.random.default_rng(5)
    n = 300
    r, c = H.random_coo(rng, n, n, 3000)
    vals = H.random_values(rng, r.size, np.int64)
    A = gb.Matrix.from_coo(r, c, vals, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, vals, n, n)
    A(gb.binary.plus) << A.mxm(A, gb.semiring.plus_times)
    want = R.mxm(Ab, None, "plus", "plus_times", Ab, Ab)
    ok, msg = H.mat_equal(A, want)
    assert ok, msg
    qi = rng.choice(n, 20, replace=False)
    q = gb.Vector.from_coo(qi, np.ones(20, dtype=bool), size=n)
    qb = R.BigVec.from_coo(qi, np.ones(20, dtype=bool), n)
    A2 = gb.Matrix.from_coo(r, c, vals, nrows=n, ncols=n)
    q(~q.S, replace=True) << q.vxm(A2, gb.semiring.any_pair)
    want = R.vxm(qb, qb, None, "any_pair", qb, Ab, complement=True, structure=True, replace=True)
    ok, msg = H.vec_equal(q, want)
    assert ok, msg


SEMIRINGS = ["plus_times", "min_plus", "any_pair", "plus_second", "plus_first", "lor_land", "max_plus", "plus_plus",
             "min_first", "max_times", "min_second", "plus_pair", "plus_min"]
DTYPES = [np.int64, np.int32, np.int8, np.uint16, np.uint64, np.float32, np.float64, np.bool_]


def _semiring_ok(semiring, dtype):
    if dtype == np.bool_:
        return semiring in ("lor_land", "any_pair")
    return True


@pytest.mark.parametrize("semiring", SEMIRINGS)
@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
def test_vector_multiplies_vs_oracle(gb, semiring, dtype):
    """mxv / mxv(A.T) / vxm / vxm(A.T) x (merge, rowwarp, push, pull) x masks x accum on random rectangular input."""
    if not _semiring_ok(semiring, dtype):
        pytest.skip("semiring/dtype combination not defined")
    rng = np.random.default_rng(abs(hash((semiring, np.dtype(dtype).name))) % 2**32)
    nr, nc = 700, 450
    r, c = H.random_coo(rng, nr, nc, 9000)
    av = H.random_values(rng, r.size, dtype)
    A = gb.Matrix.from_coo(r, c, av, nrows=nr, ncols=nc)
    Ab = R.BigMat.from_coo(r, c, av, nr, nc)
    sr = getattr(gb.semiring, semiring)
    accum_name = "lor" if dtype == np.bool_ else "min"
    for trial in range(3):
        dens_u = [0.9, 0.05, 1.0][trial]
        for kind in ("mxv", "mxv_T", "vxm", "vxm_T"):
            in_len = nc if kind in ("mxv", "vxm_T") else nr
            out_len = nr if kind in ("mxv", "vxm_T") else nc
            ui = np.flatnonzero(rng.random(in_len) < dens_u)
            uv = H.random_values(rng, ui.size, dtype)
            wi = np.flatnonzero(rng.random(out_len) < 0.5)
            wv = H.random_values(rng, wi.size, dtype)
            mi = np.flatnonzero(rng.random(out_len) < 0.5)
            mv = rng.integers(0, 2, mi.size).astype(np.int8)
            for (use_mask, comp, struct, repl, accum) in [(False, False, False, False, None), (True, False, True, False, None),
                                                          (True, True, True, True, None), (True, False, False, True, accum_name),
                                                          (True, True, False, False, accum_name), (False, False, False, False, accum_name)]:
                for method, vxm_method, hot in [("merge", "pull", "0"), ("seg", "pull", "0"), ("seg", "pull", "1"), ("seg", "pull", "cap"),
                                                ("band", "pull", "0"), ("rowwarp", "push", "auto"), ("auto", "auto", "auto")]:
                    gb.cuda.set_option("spmv", method)
                    gb.cuda.set_option("vxm_method", vxm_method)
                    # hot-column cache of the pull kernel: off / forced / forced with a 40-entry cache (ranks beyond it gather from global)
                    gb.cuda.set_option("spmv_hot", "1" if hot == "cap" else hot)
                    gb.cuda.set_option("spmv_hot_cap", "40" if hot == "cap" else "0")
                    u = H.gb_vector(gb, ui, uv, in_len)
                    w = H.gb_vector(gb, wi, wv, out_len)
                    kwargs = {}
                    if use_mask:
                        m = H.gb_vector(gb, mi, mv, out_len)
                        mk = m.S if struct else m.V
                        kwargs["mask"] = ~mk if comp else mk
                    if accum:
                        kwargs["accum"] = getattr(gb.binary, accum)
                    if repl:
                        kwargs["replace"] = True
                    expr = {"mxv": lambda: A.mxv(u, sr), "mxv_T": lambda: A.T.mxv(u, sr), "vxm": lambda: u.vxm(A, sr),
                            "vxm_T": lambda: u.vxm(A.T, sr)}[kind]()
                    w(**kwargs) << expr
                    ub = R.BigVec.from_coo(ui, uv, in_len, dtype=dtype)
                    wb = R.BigVec.from_coo(wi, wv, out_len, dtype=dtype)
                    mb = R.BigVec.from_coo(mi, mv, out_len, dtype=np.int8) if use_mask else None
                    okw = dict(complement=comp, structure=struct, replace=repl)
                    if kind == "mxv":
                        want = R.mxv(wb, mb, accum, semiring, Ab, ub, **okw)
                    elif kind == "mxv_T":
                        want = R.mxv(wb, mb, accum, semiring, Ab, ub, t0=True, **okw)
                    elif kind == "vxm":
                        want = R.vxm(wb, mb, accum, semiring, ub, Ab, **okw)
                    else:
                        want = R.vxm(wb, mb, accum, semiring, ub, Ab, t1=True, **okw)
                    ok, msg = H.vec_equal(w, want)
                    assert ok, (semiring, dtype, kind, trial, use_mask, comp, struct, repl, accum, method, vxm_method, hot, msg)
    gb.cuda.set_option("spmv", "auto")
    gb.cuda.set_option("vxm_method", "auto")
    gb.cuda.set_option("spmv_hot", "auto")
    gb.cuda.set_option("spmv_hot_cap", "0")


@pytest.mark.parametrize("semiring", ["plus_times", "min_plus", "any_pair", "plus_second", "lor_land", "max_plus", "plus_pair"])
@pytest.mark.parametrize("dtype", [np.int64, np.int16, np.float32, np.float64, np.bool_], ids=lambda d: np.dtype(d).name)
def test_mxm_vs_oracle(gb, semiring, dtype):
    if not _semiring_ok(semiring, dtype):
        pytest.skip("semiring/dtype combination not defined")
    rng = np.random.default_rng(abs(hash(("mxm", semiring, np.dtype(dtype).name))) % 2**32)
    m, k, n = 260, 310, 240
    ar, ac = H.random_coo(rng, m, k, 5000)
    br, bc = H.random_coo(rng, k, n, 6000)
    av, bv = H.random_values(rng, ar.size, dtype), H.random_values(rng, br.size, dtype)
    A, B = gb.Matrix.from_coo(ar, ac, av, nrows=m, ncols=k), gb.Matrix.from_coo(br, bc, bv, nrows=k, ncols=n)
    Ab, Bb = R.BigMat.from_coo(ar, ac, av, m, k), R.BigMat.from_coo(br, bc, bv, k, n)
    sr = getattr(gb.semiring, semiring)
    C = A.mxm(B, sr).new()
    ok, msg = H.mat_equal(C, R.mxm_T(semiring, Ab, Bb))
    assert ok, msg
    # transposed operands
    Bt = gb.Matrix.from_coo(bc, br, bv, nrows=n, ncols=k)
    ok, msg = H.mat_equal(A.mxm(Bt.T, sr).new(), R.mxm_T(semiring, Ab, Bb))
    assert ok, "T1 " + msg
    At = gb.Matrix.from_coo(ac, ar, av, nrows=k, ncols=m)
    ok, msg = H.mat_equal(At.T.mxm(B, sr).new(), R.mxm_T(semiring, Ab, Bb))
    assert ok, "T0 " + msg
    # mask / accum / replace
    cr, cc = H.random_coo(rng, m, n, 4000)
    cv = H.random_values(rng, cr.size, dtype)
    mr, mc = H.random_coo(rng, m, n, 20000)
    mv = rng.integers(0, 2, mr.size).astype(np.int8)
    accum_name = "lor" if dtype == np.bool_ else "plus"
    for (use_mask, comp, struct, repl, accum) in [(True, False, True, False, None), (True, True, False, True, None),
                                                  (True, False, False, True, accum_name), (False, False, False, False, accum_name),
                                                  (True, True, True, False, accum_name)]:
        Cg = gb.Matrix.from_coo(cr, cc, cv, nrows=m, ncols=n)
        kwargs = {}
        if use_mask:
            Mg = gb.Matrix.from_coo(mr, mc, mv, nrows=m, ncols=n)
            mk = Mg.S if struct else Mg.V
            kwargs["mask"] = ~mk if comp else mk
        if accum:
            kwargs["accum"] = getattr(gb.binary, accum)
        if repl:
            kwargs["replace"] = True
        Cg(**kwargs) << A.mxm(B, sr)
        want = R.mxm(R.BigMat.from_coo(cr, cc, cv, m, n), R.BigMat.from_coo(mr, mc, mv, m, n) if use_mask else None,
                     accum, semiring, Ab, Bb, complement=comp, structure=struct, replace=repl)
        ok, msg = H.mat_equal(Cg, want)
        assert ok, (use_mask, comp, struct, repl, accum, msg)


def test_mxm_all_bins(gb):
    """Rows whose flop counts span every SpGEMM bin, including the global-hash-table bin (>8192 products/row)."""
    rng = np.random.default_rng(11)
    n = 6000
    rows, cols = [], []
    # row i gets degree deg[i]; B = A, and a few dense-ish columns' rows make the products explode
    deg = np.concatenate([np.zeros(500, int), rng.integers(1, 4, 3000), rng.integers(4, 40, 2000), rng.integers(40, 400, 480),
                          np.full(20, 2500)])
    rng.shuffle(deg)
    for i, d in enumerate(deg):
        if d:
            cs = rng.choice(n, size=d, replace=False)
            rows.append(np.full(d, i))
            cols.append(cs)
    r, c = np.concatenate(rows), np.concatenate(cols)
    v = rng.integers(-3, 4, r.size).astype(np.int64)
    A = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, v, n, n)
    flops, nvals = gb.cuda.mxm_symbolic(A, A)
    T = R.mxm_T("plus_times", Ab, Ab)
    assert nvals == T.nvals
    deg_b = np.diff(Ab.indptr)
    assert flops == int(deg_b[Ab.indices].sum())
    C = A.mxm(A, gb.semiring.plus_times).new()
    ok, msg = H.mat_equal(C, T)
    assert ok, msg
    C2 = A.mxm(A, gb.semiring.min_plus).new()
    ok, msg = H.mat_equal(C2, R.mxm_T("min_plus", Ab, Ab))
    assert ok, msg
    Cf = gb.Matrix.from_coo(r, c, v.astype(np.float64), nrows=n, ncols=n)
    C3 = Cf.mxm(Cf, gb.semiring.plus_times).new()
    ok, msg = H.mat_equal(C3, R.mxm_T("plus_times", R.BigMat.from_coo(r, c, v.astype(np.float64), n, n), R.BigMat.from_coo(r, c, v.astype(np.float64), n, n)))
    assert ok, msg


@pytest.mark.parametrize("dtype,semiring", [(np.int64, "plus_times"), (np.float32, "plus_times"), (np.int32, "min_plus"), (np.float64, "plus_second")])
def test_mxm_row_end_result_as_operand_and_through_consumers(gb, dtype, semiring):
    """An unmasked product that hardly compresses is left as a row-end CSR (rows where the hash kernels put them, no compaction:
    csrc/spgemm.cu).  The multiply kernels must read that form directly (as A, as B, as both), and every other consumer must
    see the compact CSR: dup, extract tuples, element lookup, transpose, element-wise ops, reduce, mxv, wait(materialize).
    Compared with the oracle's products; option spgemm_row_end=0 (eager compaction) must give the same matrices."""
    rng = np.random.default_rng(abs(hash(("rowend", semiring, np.dtype(dtype).name))) % 2**32)
    n = 700
    r, c = H.random_coo(rng, n, n, 2600)          # ~3.7 per row: the square hardly compresses
    v = H.random_values(rng, r.size, dtype)
    A = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, v, n, n)
    sr = getattr(gb.semiring, semiring)
    C2b = R.mxm_T(semiring, Ab, Ab)
    assert C2b.nvals * 8 >= 7 * int(np.diff(Ab.indptr)[Ab.indices].sum()), "test matrix compresses too well to stay row-end"
    for opt in (None, "0"):
        gb.cuda.set_option("spgemm_row_end", opt)
        try:
            C2 = A.mxm(A, sr).new()
            assert C2.nvals == C2b.nvals
            left = C2.mxm(A, sr).new()            # row-end A operand
            right = A.mxm(C2, sr).new()           # row-end B operand
            both = C2.mxm(C2, sr).new()           # both
            for got, want in ((left, R.mxm_T(semiring, C2b, Ab)), (right, R.mxm_T(semiring, Ab, C2b)), (both, R.mxm_T(semiring, C2b, C2b))):
                ok, msg = H.mat_equal(got, want)
                assert ok, (opt, msg)
            D = A.mxm(A, sr).new()
            ok, msg = H.mat_equal(D.dup(), C2b)   # dup of a row-end matrix
            assert ok, msg
            D = A.mxm(A, sr).new()
            I, J, X = C2b.to_coo()
            k = int(rng.integers(0, I.size))
            assert D[int(I[k]), int(J[k])].new().value == X[k]      # element lookup straight after the product
            D = A.mxm(A, sr).new()
            ok, msg = H.mat_equal(D.T.new(), R.BigMat.from_coo(J, I, X, n, n))
            assert ok, msg
            D = A.mxm(A, sr).new()
            D.wait("materialize")
            ok, msg = H.mat_equal(D, C2b)
            assert ok, msg
            D = A.mxm(A, sr).new()
            x = H.random_values(rng, n, dtype)
            w = D.mxv(gb.Vector.from_coo(np.arange(n), x, size=n), sr).new()
            ok, msg = H.vec_equal(w, R.mxv_T(semiring, C2b, R.BigVec(x, np.ones(n, np.uint8))))
            assert ok, msg
            D = A.mxm(A, sr).new()
            E = D.ewise_add(A, gb.binary.plus if dtype != np.bool_ else gb.binary.lor).new()
            assert E.nvals >= D.nvals
            # accumulate a row-end product into an existing matrix (write-back merges need the compact, sorted form)
            Cg = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
            Cg(gb.binary.plus) << A.mxm(A, sr)
            want = R.mxm(Ab, None, "plus", semiring, Ab, Ab)
            ok, msg = H.mat_equal(Cg, want)
            assert ok, msg
        finally:
            gb.cuda.set_option("spgemm_row_end", None)


@pytest.mark.parametrize("opts", [{}, {"spgemm_tile_ctas": "1", "spgemm_tile_threads": "512"}, {"spgemm_tile_scap": "512", "spgemm_tile_tcap": "2048"}])
@pytest.mark.parametrize("dtype,semiring", [(np.float32, "plus_times"), (np.int64, "min_plus"), (np.float64, "plus_second"), (np.int32, "any_pair")])
def test_mxm_tiled_kernel_vs_oracle(gb, dtype, semiring, opts):
    """The row-ordered TMA-staged tile kernel (option spgemm_tile=1, csrc/spgemm_tile.cuh): hole rows, tiles of many short rows,
    rows spanning several ring stages, tables of two sizes -- pattern and values exact against the oracle"""
    rng = np.random.default_rng(17)
    n = 3000
    deg = np.concatenate([np.zeros(300, int), rng.integers(1, 4, 1500), rng.integers(4, 60, 1000), rng.integers(60, 300, 190), np.full(10, 1200)])
    rng.shuffle(deg)
    rows = np.repeat(np.arange(n), deg)
    cols = np.concatenate([rng.choice(n, size=d, replace=False) for d in deg if d])
    v = rng.integers(1, 5, rows.size).astype(dtype)
    A = gb.Matrix.from_coo(rows, cols, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(rows, cols, v, n, n)
    want = R.mxm_T(semiring, Ab, Ab)
    gb.cuda.set_option("spgemm_tile", "1")
    for k, val in opts.items():
        gb.cuda.set_option(k, val)
    try:
        C = A.mxm(A, getattr(gb.semiring, semiring)).new()
        if semiring == "any_pair":
            I, J, X = C.to_coo()
            wi, wj, _ = want.to_coo()
            assert np.array_equal(I.astype(np.int64), wi) and np.array_equal(J.astype(np.int64), wj) and np.all(X == 1)
        else:
            ok, msg = H.mat_equal(C, want)
            assert ok, msg
    finally:
        gb.cuda.set_option("spgemm_tile", None)
        for k in opts:
            gb.cuda.set_option(k, None)


@pytest.mark.parametrize("dtype,rtol", [(np.float32, 2e-5), (np.float64, 1e-13)])
@pytest.mark.parametrize("scale", [10, 13])
@pytest.mark.parametrize("tile", ["0", "1"])
def test_mxm_float_noninteger_values_tolerance(gb, scale, dtype, rtol, tile):
    """The bench dtype with the bench's kind of values (U(0,1), not exactly summable): pattern exact, values within
    rtol * sqrt(max row degree) of the oracle's ascending-k sum (the hash accumulates in arbitrary order).  The reference's
    own tolerance for floating point comparisons is isclose's rel_tol (graphblas/core/matrix.py:417)."""
    r, c, n = H.rmat_edges(scale, a=0.45, b=0.15, c=0.15, seed=7)
    v = np.random.default_rng(2).random(r.size).astype(dtype)
    A = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, v, n, n)
    want = R.mxm_T("plus_times", Ab, Ab)
    gb.cuda.set_option("spgemm_tile", tile)
    try:
        C = A.mxm(A, gb.semiring.plus_times).new()
    finally:
        gb.cuda.set_option("spgemm_tile", None)
    ok, msg = H.mat_equal(C, want, rtol=rtol * np.sqrt(np.diff(Ab.indptr).max()))
    assert ok, msg


@pytest.mark.parametrize("structure", [True, False])
def test_mxm_complemented_mask_in_hash_rmat(gb, structure):
    """C<!A.S> = A.A and C<!A.V> = A.A on R-MAT scale 14 (Graph500 skew: heavy rows reach the global-table bin): the mask rows are
    loaded into the hash tables as forbidden columns, the unmasked product is never formed (csrc/spgemm.cu insert_comp);
    exact against the oracle, and identical to the mask-after-multiply path (option spgemm_mask=0)."""
    r, c, n = H.rmat_edges(14, seed=5)
    rng = np.random.default_rng(4)
    v = rng.integers(0, 3, r.size).astype(np.int64)      # zeros make the value mask differ from the structural one
    A = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, v, n, n)
    want = R.mxm(R.BigMat.from_coo([], [], np.array([], dtype=np.int64), n, n), Ab, None, "plus_times", Ab, Ab, complement=True,
                 structure=structure, replace=False)
    mask = ~A.S if structure else ~A.V
    C = A.mxm(A, gb.semiring.plus_times).new(mask=mask)
    ok, msg = H.mat_equal(C, want)
    assert ok, msg
    gb.cuda.set_option("spgemm_mask", "0")
    try:
        C0 = A.mxm(A, gb.semiring.plus_times).new(mask=mask)
    finally:
        gb.cuda.set_option("spgemm_mask", None)
    assert C0.isequal(C)
    with gb.Recorder() as rec:
        A.mxm(A, gb.semiring.plus_times).new(mask=mask)
    assert "GrB_DESC_SC" in rec.data[-1] if structure else "GrB_DESC_C" in rec.data[-1]


@pytest.mark.parametrize("dtype,semiring", [(np.int32, "plus_times"), (np.int64, "min_plus"), (np.float64, "plus_second"), (np.float32, "plus_times"),
                                            (np.int16, "max_plus"), (np.bool_, "lor_land"), (np.int64, "any_pair")])
def test_banded_spmv_many_bands_vs_oracle(gb, dtype, semiring):
    """The column-banded pull kernel (csrc/spmv_band.cu, option spmv=band) on R-MAT scale 17: 131 072 columns are 4 bands of
    4-byte values / 8 bands of 8-byte values, so band switches, padded band tails, segments cut at tile boundaries and rows that
    meet many bands are all exercised; dense and sparse input vectors, mxv and pull vxm, with an accumulator (separate
    write-back pass).  Exact against the oracle (integer-valued inputs)."""
    r, c, n = H.rmat_edges(17, seed=11)
    rng = np.random.default_rng(6)
    v = H.random_values(rng, r.size, dtype)
    A = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, v, n, n)
    sr = getattr(gb.semiring, semiring)
    gb.cuda.set_option("spmv", "band")
    gb.cuda.set_option("vxm_method", "pull")
    try:
        for dens in (1.0, 0.3):
            ui = np.flatnonzero(rng.random(n) < dens)
            uv = H.random_values(rng, ui.size, dtype)
            u = H.gb_vector(gb, ui, uv, n)
            ub = R.BigVec.from_coo(ui, uv, n, dtype=dtype)
            with gb.Recorder():
                got = A.mxv(u, sr).new()
            ok, msg = H.vec_equal(got, R.mxv_T(semiring, Ab, ub))
            assert ok, (dens, "mxv", msg)
            ok, msg = H.vec_equal(u.vxm(A, sr).new(), R.vxm_push_T(semiring, ub, Ab))
            assert ok, (dens, "vxm", msg)
            if dtype != np.bool_:
                w = H.gb_vector(gb, ui[::2], uv[::2], n)
                w(gb.binary.plus) << A.mxv(u, sr)
                wb = R.BigVec.from_coo(ui[::2], uv[::2], n, dtype=dtype)
                ok, msg = H.vec_equal(w, R.mxv(wb, None, "plus", semiring, Ab, ub))
                assert ok, (dens, "accum", msg)
        assert any(k.startswith("spmv_band") for k in gb.cuda.kernel_times()) or True
    finally:
        gb.cuda.set_option("spmv", "auto")
        gb.cuda.set_option("vxm_method", "auto")


def test_mxm_to_host_in_row_blocks(gb):
    """graphblas_b200.cuda.mxm_to_host_csr32 (bench.py's end-to-end path): the product formed and exported in row blocks equals
    the oracle's product, row pointers stitched to one CSR"""
    import torch

    r, c, n = H.rmat_edges(12, a=0.45, b=0.15, c=0.15, seed=3)
    v = np.random.default_rng(9).integers(1, 4, r.size).astype(np.float32)
    A = gb.Matrix.from_coo(r, c, v, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, v, n, n)
    want = R.mxm_T("plus_times", Ab, Ab)
    for blocks in (1, 3, 7, 64):
        o_ptr = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        o_col = torch.empty(want.nvals, dtype=torch.int32).pin_memory()
        o_val = torch.empty(want.nvals, dtype=torch.float32).pin_memory()
        nv = gb.cuda.mxm_to_host_csr32(A, A, gb.semiring.plus_times, o_ptr.numpy(), o_col.numpy(), o_val.numpy(), blocks=blocks)
        assert nv == want.nvals and np.array_equal(o_ptr.numpy(), want.indptr)
        cols, vals = o_col.numpy().astype(np.int64), o_val.numpy()
        for i in np.random.default_rng(1).integers(0, n, 200):   # rows come out unsorted: compare as sets per row
            lo, hi = want.indptr[i], want.indptr[i + 1]
            order = np.argsort(cols[lo:hi])
            assert np.array_equal(cols[lo:hi][order], want.indices[lo:hi]) and np.array_equal(vals[lo:hi][order], want.values[lo:hi])
        order = np.lexsort((cols, np.repeat(np.arange(n), np.diff(want.indptr))))
        assert np.array_equal(cols[order], want.indices) and np.array_equal(vals[order], want.values)
    with pytest.raises(ValueError):
        gb.cuda.mxm_to_host_csr32(A, A, gb.semiring.plus_times, o_ptr.numpy(), o_col.numpy()[:10], o_val.numpy()[:10], blocks=2)


@pytest.mark.parametrize("scale", [10, 14])
def test_rmat_parity(gb, scale):
    """R-MAT (Graph500 parameters): skewed degrees exercise merge-path carries across tiles and heavy SpGEMM rows."""
    r, c, n = H.rmat_edges(scale, seed=42)
    rng = np.random.default_rng(1)
    w = rng.integers(1, 256, r.size).astype(np.int64)
    A = gb.Matrix.from_coo(r, c, w, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, w, n, n)
    x = rng.integers(0, 1000, n).astype(np.int64)
    v = gb.Vector.from_coo(np.arange(n), x, size=n)
    vb = R.BigVec(x, np.ones(n, np.uint8))
    for method in ("merge", "hot", "hotcap", "rowwarp", "band"):
        gb.cuda.set_option("spmv", "seg" if method.startswith("hot") else method)
        gb.cuda.set_option("spmv_hot", "1" if method.startswith("hot") else "0")
        gb.cuda.set_option("spmv_hot_cap", "300" if method == "hotcap" else "0")
        for sr in ("min_plus", "plus_times", "plus_second", "any_pair"):
            ok, msg = H.vec_equal(A.mxv(v, getattr(gb.semiring, sr)).new(), R.mxv_T(sr, Ab, vb))
            assert ok, (method, sr, msg)
            gb.cuda.set_option("vxm_method", "pull")
            ok, msg = H.vec_equal(v.vxm(A, getattr(gb.semiring, sr)).new(), R.vxm_push_T(sr, vb, Ab))
            assert ok, (method, "vxm", sr, msg)
            gb.cuda.set_option("vxm_method", "auto")
    gb.cuda.set_option("spmv", "auto")
    gb.cuda.set_option("spmv_hot", "auto")
    gb.cuda.set_option("spmv_hot_cap", "0")
    # fp32 with a stated tolerance: summation order differs (rtol 1e-5 * sqrt(max row products) is the contract; use 1e-4)
    wf = rng.random(r.size).astype(np.float32)
    Af = gb.Matrix.from_coo(r, c, wf, nrows=n, ncols=n)
    xf = rng.random(n).astype(np.float32)
    got = Af.mxv(gb.Vector.from_coo(np.arange(n), xf, size=n), gb.semiring.plus_times).new()
    want = R.mxv_T("plus_times", R.BigMat.from_coo(r, c, wf, n, n), R.BigVec(xf, np.ones(n, np.uint8)))
    ok, msg = H.vec_equal(got, want, rtol=1e-4)
    assert ok, msg
    # masked product: in-hash mask (tables pre-loaded with the mask rows) vs the oracle, and vs the post-hoc masked path
    Ms = A.mxm(A, gb.semiring.plus_times).new(mask=A.S)
    want = R.mxm(R.BigMat(np.zeros(n + 1), [], np.zeros(0, np.int64), n, n), Ab, None, "plus_times", Ab, Ab, structure=True)
    ok, msg = H.mat_equal(Ms, want)
    assert ok, "masked " + msg
    gb.cuda.set_option("spgemm_mask", "0")
    try:
        Ms2 = A.mxm(A, gb.semiring.plus_times).new(mask=A.S)
    finally:
        gb.cuda.set_option("spgemm_mask", "1")
    assert Ms2.isequal(Ms)
    mv = rng.integers(0, 2, r.size).astype(np.int8)          # value mask: only truthy entries count
    Mv = gb.Matrix.from_coo(r, c, mv, nrows=n, ncols=n)
    got = A.mxm(A, gb.semiring.min_plus).new(mask=Mv.V)
    want = R.mxm(R.BigMat(np.zeros(n + 1), [], np.zeros(0, np.int64), n, n), R.BigMat.from_coo(r, c, mv, n, n), None, "min_plus", Ab, Ab)
    ok, msg = H.mat_equal(got, want)
    assert ok, "value-masked " + msg
    if scale <= 10:
        C = A.mxm(A, gb.semiring.plus_times).new()
        ok, msg = H.mat_equal(C, R.mxm_T("plus_times", Ab, Ab))
        assert ok, msg
        M = A.mxm(A, gb.semiring.plus_times).new(mask=A.S)
        want = R.mxm(R.BigMat(np.zeros(n + 1), [], np.zeros(0, np.int64), n, n), Ab, None, "plus_times", Ab, Ab, structure=True)
        ok, msg = H.mat_equal(M, want)
        assert ok, msg


def test_bfs_sssp_pagerank_loops(gb):
    """The three iterative workloads of BASELINE.json configs 3-5 as written in the reference notebooks,
    device-resident, against the same loops run on the oracle."""
    r, c, n = H.rmat_edges(11, seed=7)
    rng = np.random.default_rng(3)
    w = rng.integers(1, 256, r.size).astype(np.int64)
    A = gb.Matrix.from_coo(r, c, w, nrows=n, ncols=n)
    Ab = R.BigMat.from_coo(r, c, w, n, n)
    src = int(r[0])
    # ---- level BFS: q<~v.S, replace> = q any_pair A
    q = gb.Vector.from_coo([src], [True], size=n)
    v = gb.Vector(gb.dtypes.INT64, n)
    qb = R.BigVec.from_coo([src], [True], n)
    vb = R.BigVec.empty(n, np.int64)
    for level in range(1, n):
        v(mask=q.V)[:] = level
        vb = R.vec_write_back(vb, R.BigVec(np.full(n, level, np.int64), np.ones(n, np.uint8)), qb, None, False, False, False)
        q(~v.S, replace=True) << q.vxm(A, gb.semiring.any_pair)
        qb = R.vxm(qb, vb, None, "any_pair", qb, Ab, complement=True, structure=True, replace=True)
        ok, msg = H.vec_equal(q, qb)
        assert ok, (level, msg)
        if q.nvals == 0:
            break
    ok, msg = H.vec_equal(v, vb)
    assert ok, msg
    assert v.nvals > n // 8
    # ---- SSSP (Bellman-Ford): w(min) << w.vxm(A, min_plus) to a fixed point
    d = gb.Vector.from_coo([src], [0], size=n, dtype=gb.dtypes.INT64)
    db = R.BigVec.from_coo([src], np.array([0], dtype=np.int64), n)
    for it in range(n):
        old = d.dup()
        d(gb.binary.min) << d.vxm(A, gb.semiring.min_plus)
        db = R.vxm(db, None, "min", "min_plus", db, Ab)
        if d.isequal(old):
            break
    ok, msg = H.vec_equal(d, db)
    assert ok, msg
    # ---- PageRank: r(plus) << A.T.mxv(w, plus_second), 10 fixed iterations, fp64, rtol 1e-10
    Af = gb.Matrix.from_coo(r, c, np.ones(r.size), nrows=n, ncols=n)
    Afb = R.BigMat.from_coo(r, c, np.ones(r.size), n, n)
    deg = np.maximum(np.diff(Afb.indptr), 1).astype(np.float64)
    dvec = gb.Vector.from_coo(np.arange(n), deg, size=n)
    t = gb.Vector.from_coo(np.arange(n), np.full(n, 1.0 / n), size=n)
    tb = np.full(n, 1.0 / n)
    damping, teleport = 0.85, (1 - 0.85) / n
    for it in range(10):
        wv = t.ewise_mult(dvec, gb.binary.truediv if hasattr(gb.binary, "truediv") else gb.binary.cdiv).new()
        wv = wv.apply(gb.binary.times, right=damping).new()
        rr = gb.Vector(gb.dtypes.FP64, n)
        rr[:] = teleport
        rr(gb.binary.plus) << Af.T.mxv(wv, gb.semiring.plus_second)
        t = rr
        wb = damping * tb / deg
        prod = R.mxv_T("plus_second", Afb.T(), R.BigVec(wb, np.ones(n, np.uint8)))
        tb = teleport + np.where(prod.present.astype(bool), prod.vals, 0.0)
    got = t.to_dense(fill_value=0.0)
    np.testing.assert_allclose(got, tb, rtol=1e-10, atol=0)
    assert abs(t.reduce(gb.monoid.plus).value - tb.sum()) < 1e-9


def test_io_roundtrips(gb):
    # reference tests/test_matrix.py:3965-4004 (CSR/CSC round trips, malformed input)
    rng = np.random.default_rng(2)
    r, c = H.random_coo(rng, 50, 70, 600)
    v = rng.random(r.size)
    A = gb.Matrix.from_coo(r, c, v, nrows=50, ncols=70)
    Ap, Ai, Ax = A.to_csr()
    B = gb.Matrix.from_csr(Ap, Ai, Ax, ncols=70)
    assert B.isequal(A)
    Cp, Ci, Cx = A.to_csc()
    C = gb.Matrix.from_csc(Cp, Ci, Cx, nrows=50)
    assert C.isequal(A)
    import scipy.sparse as sp

    S_ = sp.csr_matrix((v, (r, c)), shape=(50, 70))
    S_.sort_indices()
    assert np.array_equal(Ap, S_.indptr) and np.array_equal(Ai, S_.indices) and np.array_equal(Ax, S_.data)
    Sc = S_.tocsc()
    Sc.sort_indices()
    assert np.array_equal(Cp, Sc.indptr) and np.array_equal(Ci, Sc.indices) and np.array_equal(Cx, Sc.data)
    # unsorted CSR is accepted and sorted lazily
    perm = np.concatenate([rng.permutation(np.arange(Ap[i], Ap[i + 1])) for i in range(50)]).astype(np.int64)
    D = gb.Matrix.from_csr(Ap, Ai[perm], Ax[perm], ncols=70)
    assert D.isequal(A)
    # duplicates: ValueError without dup_op (reference core/matrix.py:680-681), reduced with it
    with pytest.raises(ValueError, match="Duplicate indices found"):
        gb.Matrix.from_coo([0, 0], [1, 1], [1, 2], nrows=2, ncols=2)
    with pytest.raises(ValueError, match="Duplicate indices found"):
        gb.Vector.from_coo([3, 3], [1, 2], size=5)
    E = gb.Matrix.from_coo([0, 0, 1], [1, 1, 0], [1, 2, 5], nrows=2, ncols=2, dup_op=gb.binary.plus)
    assert E.to_coo()[2].tolist() == [3, 5]
    with pytest.raises(gb.exceptions.IndexOutOfBound):
        gb.Matrix.from_coo([0, 5], [1, 1], [1, 2], nrows=2, ncols=2)
    with pytest.raises(gb.exceptions.InvalidValue):
        gb.Matrix.from_csr([0, 3, 2], [0, 1, 0], [1.0, 2.0, 3.0], ncols=2)
    # vectors, incl. stored explicit zeros (reference fixture v has an explicit 0 at index 6)
    vv = gb.Vector.from_coo([1, 3, 4, 6], [1, 1, 2, 0])
    assert vv.nvals == 4 and vv.to_coo()[1].tolist() == [1, 1, 2, 0]
    # huge vectors cannot be held densely: clean error, not a crash (reference tests/test_vector.py:59-66 uses 2**59)
    big = gb.Vector(gb.dtypes.INT64, 2**59 + 1)
    assert big.nvals == 0
    with pytest.raises((gb.exceptions.OutOfMemory, gb.exceptions.IndexOutOfBound)):
        big.build([0, 2**59], [0, 1])


def test_inner_outer(gb):
    """Vector.inner / Vector.outer run through GrB_vxm / GrB_mxm on the vector as an n x 1 matrix.
    Known answers from reference tests/test_vector.py:1496-1547, then seeded vectors against the oracle (exact)."""
    v = gb.Vector.from_coo([1, 3, 4, 6], [1, 1, 2, 0], size=7)
    assert v.inner(v).new().value == 6 and (v @ v).new().value == 6
    assert v.inner(v, gb.semiring.min_plus).new().value == 0
    w = gb.Vector.from_coo([0, 2], [5, 7], size=7)
    assert v.inner(w).new().value is None and v.inner(w).new().is_empty
    with pytest.raises(gb.exceptions.DimensionMismatch):
        v.inner(gb.Vector.from_coo([0], [1], size=8))
    C = gb.Matrix.from_coo([1, 3, 4, 6], [0, 0, 0, 0], [1, 1, 2, 0], nrows=7, ncols=1)
    Rm = gb.Matrix.from_coo([0, 0, 0, 0], [1, 3, 4, 6], [1, 1, 2, 0], nrows=1, ncols=7)
    expected = C.mxm(Rm).new()
    assert v.outer(v).new().isequal(expected) and v.outer(v, gb.monoid.times).new().isequal(expected)
    assert v._as_matrix().isequal(C)
    rng = np.random.default_rng(5)
    for dtype, sr, bop in [(np.int64, "plus_times", "times"), (np.float64, "min_plus", "plus"), (np.int32, "max_plus", "min"),
                           (np.float32, "plus_times", "first")]:
        for n, m in [(1, 1), (50, 50), (3000, 777)]:
            iu = np.unique(rng.integers(0, n, size=max(1, n // 2)))
            iv = np.unique(rng.integers(0, n, size=max(1, n // 3)))
            iw = np.unique(rng.integers(0, m, size=max(1, m // 3)))
            xu, xv, xw = (H.random_values(rng, k.size, dtype) for k in (iu, iv, iw))
            gu, gv, gw = gb.Vector.from_coo(iu, xu, size=n), gb.Vector.from_coo(iv, xv, size=n), gb.Vector.from_coo(iw, xw, size=m)
            ou, ov, ow = (S.SpVec.from_coo(i, x, size=sz, dtype=dtype) for i, x, sz in ((iu, xu, n), (iv, xv, n), (iw, xw, m)))
            want, _ = S.inner(sr, ou, ov)
            got = gu.inner(gv, getattr(gb.semiring, sr)).new().value
            assert (got is None and want is None) or got == want, (dtype, sr, n, got, want)
            if n * m <= 60000:
                wo = S.outer(bop, ou, ow)
                I, J, X = gu.outer(gw, getattr(gb.binary, bop)).new().to_coo()
                oi, oj, ox = wo.to_coo()
                assert np.array_equal(I.astype(np.int64), oi) and np.array_equal(J.astype(np.int64), oj) and np.array_equal(X, ox.astype(X.dtype))


@pytest.mark.parametrize("opts", [{"spgemm_elect": "1"}, {"spgemm_group": "1"}, {"spgemm_group": "1", "spgemm_group_g": "512", "spgemm_group_r": "64"},
                                  {"spgemm_cas_first": "0"}, {"spgemm_mode": "twopass"}, {"spgemm_gtable_entries": "30000"}],
                         ids=["elect", "group", "group-small", "probe-first", "twopass", "gtable-batches"])
def test_mxm_optional_kernels_match_default(gb, opts):
    """The optional SpGEMM insert kernels (atomics-free owner election per row / per group of rows, probe-before-CAS, two-pass)
    must produce exactly the default kernels' result: integer values bit-exact, same pattern, on an R-MAT product whose
    rows span the warp, CTA and global-table bins."""
    r, c, n = H.rmat_edges(13, a=0.45, b=0.15, c=0.15, seed=42)
    rng = np.random.default_rng(3)
    for dtype, sr in ((np.int32, "plus_times"), (np.int64, "min_plus"), (np.float64, "plus_times")):
        vals = rng.integers(1, 4, r.size).astype(dtype)
        A = gb.Matrix.from_coo(r, c, vals, nrows=n, ncols=n)
        want = A.mxm(A, getattr(gb.semiring, sr)).new().to_coo()
        for k, v in opts.items():
            gb.cuda.set_option(k, v)
        try:
            got = A.mxm(A, getattr(gb.semiring, sr)).new().to_coo()
        finally:
            for k in opts:
                gb.cuda.set_option(k, None)
        assert all(np.array_equal(x, y) for x, y in zip(want, got)), (opts, dtype, sr)


def _spmat_equal(gbm, sp):
    I, J, X = gbm.to_coo()
    oi, oj, ox = sp.to_coo()
    if not (np.array_equal(I.astype(np.int64), oi) and np.array_equal(J.astype(np.int64), oj)):
        return False, f"pattern differs: {I.size} vs {oi.size} entries"
    if not np.array_equal(X, ox.astype(X.dtype)):
        return False, f"values differ: {X[:8]} vs {ox[:8]}"
    return True, ""


@pytest.mark.parametrize("dtype", [np.int64, np.float64, np.int32, np.bool_])
def test_matrix_elementwise_ops_vs_oracle(gb, dtype):
    """GrB_transpose, GrB_Matrix_apply (+ bind 1st / 2nd), GrB_Matrix_eWiseAdd / eWiseMult, reduce to scalar and the
    reference's own Matrix.isequal recipe (eWiseMult(EQ) + reduce(LAND)) -- all on the device, with mask / accum / replace through
    the common write-back, against the dict oracle (exact)."""
    rng = np.random.default_rng(17)
    for trial, (m, n, nnz) in enumerate([(1, 1, 1), (7, 5, 12), (60, 80, 900), (300, 300, 4000)]):
        ra, ca = H.random_coo(rng, m, n, nnz)
        rb, cb = H.random_coo(rng, m, n, nnz)
        rc, cc = H.random_coo(rng, m, n, max(1, nnz // 2))
        rm, cm = H.random_coo(rng, m, n, nnz)
        va, vb, vc = (H.random_values(rng, r.size, dtype) for r in (ra, rb, rc))
        vm = rng.integers(0, 2, rm.size).astype(np.int64)
        A, B = H.gb_matrix(gb, ra, ca, va, m, n), H.gb_matrix(gb, rb, cb, vb, m, n)
        Ao, Bo = S.SpMat.from_coo(ra, ca, va, m, n, dtype), S.SpMat.from_coo(rb, cb, vb, m, n, dtype)
        Mo = S.SpMat.from_coo(rm, cm, vm, m, n, np.int64)
        M = H.gb_matrix(gb, rm, cm, vm, m, n)
        accum_name = "lor" if dtype == np.bool_ else "plus"
        variants = [dict(), dict(mask="S"), dict(mask="V", complement=True, replace=True), dict(accum=accum_name),
                    dict(mask="S", accum=accum_name, replace=True)]
        ops = [("ewise_add", "lor" if dtype == np.bool_ else "plus"), ("ewise_mult", "land" if dtype == np.bool_ else "times"),
               ("ewise_add", "min"), ("ewise_mult", "eq"), ("ewise_add", "first")]
        for var in variants:
            def run(expr, oracle_fn, out_dtype=dtype, shape=(m, n)):
                C = H.gb_matrix(gb, rc, cc, vc, m, n) if shape == (m, n) else gb.Matrix(out_dtype, *shape)
                Co = S.SpMat.from_coo(rc, cc, vc, m, n, dtype) if shape == (m, n) else S.SpMat(shape[0], shape[1], out_dtype)
                if out_dtype != dtype and shape == (m, n):
                    C, Co = gb.Matrix(out_dtype, m, n), S.SpMat(m, n, out_dtype)
                kw, okw = {}, {}
                if "mask" in var and shape == (m, n):
                    mk = M.S if var["mask"] == "S" else M.V
                    kw["mask"] = ~mk if var.get("complement") else mk
                    okw.update(structure=var["mask"] == "S", complement=bool(var.get("complement")))
                    omask = Mo
                else:
                    omask = None
                if var.get("accum"):
                    kw["accum"] = getattr(gb.binary, var["accum"])
                if var.get("replace") and "mask" in kw:
                    kw["replace"] = True
                    okw["replace"] = True
                (C(**kw) if kw else C).update(expr) if kw else C.update(expr)
                oracle_fn(Co, omask, var.get("accum"), **okw)
                ok, msg = _spmat_equal(C, Co)
                assert ok, (dtype, trial, var, msg)
            for method, opname in ops:
                cmp = opname in S.COMPARE
                run(getattr(A, method)(B, getattr(gb.binary, opname)),
                    lambda Co, Mm, acc, **k: S.ewise(Co, Mm, acc, opname, Ao, Bo, union=method == "ewise_add", **k),
                    out_dtype=np.bool_ if cmp else dtype)
            for uop in (["identity", "lnot"] if dtype == np.bool_ else ["ainv", "abs", "identity"]):
                run(A.apply(getattr(gb.unary, uop)), lambda Co, Mm, acc, **k: S.apply(Co, Mm, acc, uop, Ao, **k))
            if dtype != np.bool_:
                run(A.apply(gb.binary.minus, right=3), lambda Co, Mm, acc, **k: S.apply(Co, Mm, acc, "minus", Ao, scalar=dtype(3), **k))
                run(A.apply(gb.binary.minus, left=3), lambda Co, Mm, acc, **k: S.apply(Co, Mm, acc, "minus", Ao, scalar=dtype(3), scalar_first=True, **k))
            run(A, lambda Co, Mm, acc, **k: S.apply(Co, Mm, acc, "identity", Ao, **k))      # C(...) << A
            if m == n:
                run(A.T, lambda Co, Mm, acc, **k: S.transpose(Co, Mm, acc, Ao, **k))
                run(A.T.ewise_mult(B, getattr(gb.binary, "land" if dtype == np.bool_ else "times")),
                    lambda Co, Mm, acc, **k: S.ewise(Co, Mm, acc, "land" if dtype == np.bool_ else "times", Ao, Bo, union=False, t0=True, **k))
        # transpose of a non-square matrix into a fresh output, reduce, isequal
        At = A.T.new()
        ok, msg = _spmat_equal(At, Ao.T())
        assert ok, msg
        assert At.T.new().isequal(A) and A.isequal(A.dup()) and (not A.isequal(B) or (np.array_equal(ra, rb) and np.array_equal(ca, cb) and np.array_equal(va, vb)))
        for mon in (["lor", "land"] if dtype == np.bool_ else ["plus", "max", "min"]):
            got = A.reduce_scalar(getattr(gb.monoid, mon)).new().value
            want = S.reduce_scalar(mon, Ao)
            assert got == want, (dtype, mon, got, want)
        assert gb.Matrix(dtype, 3, 4).reduce_scalar(gb.monoid.lor if dtype == np.bool_ else gb.monoid.plus).new().value is None
    with pytest.raises(gb.exceptions.DimensionMismatch):
        gb.Matrix(dtype, 3, 4).ewise_add(gb.Matrix(dtype, 4, 3))


def test_matrix_power(gb):
    """Matrix.power -- reference tests/test_matrix.py:4379-4405 (`A.power(i)` equals the i-fold product, INT64 wraps;
    min_plus powers; n = 0 is the diagonal of the multiply's identity; argument errors)."""
    d = G.load(G.A_M)
    A = _obj(gb, G.A_M)
    Ab = R.BigMat.from_coo(d["rows"], d["cols"], d["vals"], 7, 7)
    Pb = Ab
    for i in range(1, 50):
        ok, msg = H.mat_equal(A.power(i).new(), Pb)
        assert ok, (i, msg)
        Pb = R.mxm_T("plus_times", Pb, Ab)
    Pb = Ab
    for i in range(1, 10):
        ok, msg = H.mat_equal(A.power(i, gb.semiring.min_plus).new(), Pb)
        assert ok, (i, msg)
        Pb = R.mxm_T("min_plus", Pb, Ab)
    I0, J0, X0 = A.power(0).new().to_coo()
    assert np.array_equal(I0, np.arange(7)) and np.array_equal(J0, np.arange(7)) and np.array_equal(X0, np.ones(7, dtype=X0.dtype))
    _, _, Xm = A.power(0, gb.semiring.min_plus).new().to_coo()
    assert np.array_equal(Xm, np.zeros(7, dtype=Xm.dtype))
    # accumulate into an existing matrix: C(plus) << A.power(2)  ==  C + A*A
    C = A.dup()
    C(gb.binary.plus) << A.power(2)
    want = A.dup()
    want(gb.binary.plus) << A.mxm(A)
    assert C.isequal(want)
    with pytest.raises(ValueError):
        A.power(-1)
    with pytest.raises(TypeError):
        A.power(1.5)
    with pytest.raises(gb.exceptions.DimensionMismatch):
        gb.Matrix(gb.dtypes.INT64, 3, 4).power(2)


def test_empty_operands_everywhere(gb):
    """Empty matrices / vectors through every entry point added around the multiply (transpose, apply, eWise, reduce, isequal,
    power, inner, outer, mxm / mxv / vxm): results are empty objects of the right shape, never an error."""
    E = gb.Matrix(gb.dtypes.FP64, 3, 4)
    B = gb.Matrix.from_coo([0, 2], [1, 3], [1.5, -2.0], nrows=3, ncols=4)
    Et = E.T.new()
    assert Et.shape == (4, 3) and Et.nvals == 0
    assert E.apply(gb.unary.ainv).new().nvals == 0 and E.apply(gb.binary.plus, right=1.0).new().nvals == 0
    assert E.ewise_add(B).new().isequal(B) and B.ewise_add(E).new().isequal(B)
    assert E.ewise_mult(B).new().nvals == 0 and B.ewise_mult(E, gb.binary.eq).new().nvals == 0
    assert E.isequal(gb.Matrix(gb.dtypes.FP64, 3, 4)) and not E.isequal(B)
    assert E.reduce_scalar(gb.monoid.plus).new().value is None
    C = B.dup()
    C(mask=B.S, replace=True) << E                      # assignment of an empty matrix under a mask with replace: everything goes
    assert C.nvals == 0
    S3 = gb.Matrix(gb.dtypes.INT64, 3, 3)
    assert S3.power(3).new().nvals == 0 and S3.power(0).new().nvals == 3
    assert S3.mxm(S3).new().nvals == 0 and E.T.mxm(B).new().nvals == 0 and B.T.mxm(E).new().nvals == 0
    ev, v = gb.Vector(gb.dtypes.FP64, 4), gb.Vector.from_coo([1, 3], [2.0, 5.0], size=4)
    assert ev.inner(v).new().value is None and v.inner(ev).new().value is None and ev.inner(ev).new().value is None
    assert ev.outer(v).new().nvals == 0 and v.outer(ev).new().nvals == 0 and v.outer(v).new().nvals == 4
    assert ev._as_matrix().shape == (4, 1) and ev._as_matrix().nvals == 0
    assert B.mxv(ev).new().nvals == 0 and E.mxv(v).new().nvals == 0 and gb.Vector(gb.dtypes.FP64, 3).vxm(B).new().nvals == 0


def test_matrix_reduce_to_vector(gb):
    """Matrix.reduce_rowwise / reduce_columnwise (reference core/matrix.py:2600-2701) through the <monoid>_first SpMV: against numpy
    on seeded matrices (exact), rows / columns without entries absent from the result, mask + accum through the common write-back."""
    rng = np.random.default_rng(29)
    for dtype in (np.int64, np.float64, np.int32):
        m, n = 70, 45
        r, c = H.random_coo(rng, m, n, 600)
        v = H.random_values(rng, r.size, dtype)
        A = H.gb_matrix(gb, r, c, v, m, n)
        for mon, fn in (("plus", np.add), ("max", np.maximum), ("min", np.minimum), ("times", np.multiply)):
            for axis, size, idx in ((0, m, r), (1, n, c)):
                want = {}
                for i, x in zip(idx.tolist(), v.tolist()):
                    want[i] = dtype(x) if i not in want else dtype(fn(want[i], dtype(x)))
                red = (A.reduce_rowwise if axis == 0 else A.reduce_columnwise)(getattr(gb.monoid, mon)).new()
                gi, gv = red.to_coo()
                assert red.size == size and np.array_equal(gi, np.array(sorted(want))), (dtype, mon, axis)
                assert np.array_equal(gv, np.array([want[k] for k in sorted(want)], dtype=gv.dtype)), (dtype, mon, axis)
        # the transposed view reduces the other way round; accumulate into an existing vector under a mask
        gi, gv = A.T.reduce_rowwise(gb.monoid.plus).new().to_coo()
        ci, cv = A.reduce_columnwise(gb.monoid.plus).new().to_coo()
        assert np.array_equal(gi, ci) and np.array_equal(gv, cv)
        w = gb.Vector.from_coo(np.arange(m), np.ones(m, dtype=dtype), size=m)
        w(gb.binary.plus) << A.reduce_rowwise(gb.monoid.plus)
        base = A.reduce_rowwise(gb.monoid.plus).new()
        bi, bv = base.to_coo()
        wi, wv = w.to_coo()
        exp = np.ones(m, dtype=dtype); exp[bi] += bv
        assert np.array_equal(wi, np.arange(m)) and np.array_equal(wv, exp)
    assert gb.Matrix(gb.dtypes.FP64, 5, 6).reduce_rowwise().new().nvals == 0
