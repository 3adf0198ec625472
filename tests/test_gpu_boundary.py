"""GPU parity tests of the boundary rows added in round 2: select, GrB_Scalar results of
reduce, the typed bind-1st / bind-2nd apply names, whole-object assign, Matrix_setElement.  Every call goes through the
C-ABI under the name the reference formats; expected values come from the reference's own known answers
(graphblas/tests/test_matrix.py:1238-1271) and from the oracle (oracle/semantics.py) on seeded random inputs."""
import ctypes

import numpy as np
import pytest

import golden_cases as G
import helpers as H
from oracle import semantics as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb():
    import graphblas_b200 as gb

    gb.init()
    return gb


def _fixture_A(gb):
    d = G.load(G.A_M)
    return gb.Matrix.from_coo(d["rows"], d["cols"], d["vals"], nrows=d["nrows"], ncols=d["ncols"]), d


def _coo_equal(got, rows, cols, vals):
    I, J, X = got.to_coo()
    order = np.lexsort((cols, rows))
    return (np.array_equal(I.astype(np.int64), np.asarray(rows)[order]) and np.array_equal(J.astype(np.int64), np.asarray(cols)[order])
            and np.array_equal(X, np.asarray(vals)[order].astype(X.dtype)))


def test_select_reference_known_answers(gb):
    """reference graphblas/tests/test_matrix.py:1238-1271 (fixture A: :34-49)"""
    A, _ = _fixture_A(gb)
    for expr in (A.select(gb.select.valueeq, 3), A.select("==", 3), gb.select.valueeq(A, 3)):
        assert _coo_equal(expr.new(), [0, 3, 3, 6], [3, 0, 2, 4], [3, 3, 3, 3])
    for expr in (gb.select.colle(A, 2), A.select("col<=", 2)):
        assert _coo_equal(expr.new(), [3, 0, 3, 5, 6], [0, 1, 2, 2, 2], [3, 2, 3, 1, 5])
    assert _coo_equal(A.select("TRIU").new(), [0, 0, 1, 2, 4, 1], [1, 3, 4, 5, 5, 6], [2, 3, 8, 1, 7, 4])
    for expr in (gb.select.rowle(A, 2), A.select("row<=", 2)):
        assert _coo_equal(expr.new(), [0, 0, 1, 1, 2], [1, 3, 4, 6, 5], [2, 3, 8, 4, 1])
    with pytest.raises(TypeError):
        A.select(gb.binary.plus, 3)


SELECT_CASES = [("tril", 0), ("tril", -2), ("triu", 1), ("diag", 0), ("diag", 3), ("offdiag", 0), ("colle", 7), ("colgt", 20),
                ("rowle", 11), ("rowgt", 30), ("valueeq", 2), ("valuene", 0), ("valuegt", 1), ("valuege", -1), ("valuelt", 0.5),
                ("valuele", 3)]


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64, np.int8, np.bool_])
@pytest.mark.parametrize("op,thunk", SELECT_CASES)
def test_matrix_select_vs_oracle(gb, op, thunk, dtype):
    if dtype == np.bool_ and op.startswith("value") and not isinstance(thunk, int):
        pytest.skip("fractional thunk for BOOL entries")
    rng = np.random.default_rng(hash((op, str(dtype))) % 2**31)
    m, n = 37, 45
    r, c = H.random_coo(rng, m, n, 400)
    v = H.random_values(rng, r.size, dtype)
    A, Ao = gb.Matrix.from_coo(r, c, v, nrows=m, ncols=n), S.SpMat.from_coo(r, c, v, m, n)
    mr, mc = H.random_coo(rng, m, n, 500)
    mv = rng.integers(0, 2, mr.size).astype(bool)
    M, Mo = gb.Matrix.from_coo(mr, mc, mv, nrows=m, ncols=n), S.SpMat.from_coo(mr, mc, mv, m, n)
    cr, cc = H.random_coo(rng, m, n, 300)
    cv = H.random_values(rng, cr.size, dtype)
    sel = getattr(gb.select, op)
    # plain
    got = A.select(sel, thunk).new()
    want = S.select(S.SpMat(m, n, np.dtype(dtype)), None, None, op, Ao, thunk)
    assert _coo_equal(got, *want.to_coo()), (op, thunk)
    # transposed input, value mask, accumulator, replace
    At, Ato = gb.Matrix.from_coo(c, r, v, nrows=n, ncols=m), S.SpMat.from_coo(c, r, v, n, m)
    C, Co = gb.Matrix.from_coo(cr, cc, cv, nrows=m, ncols=n), S.SpMat.from_coo(cr, cc, cv, m, n)
    accum = "lor" if dtype == np.bool_ else "plus"
    C(mask=~M.V, accum=getattr(gb.binary, accum), replace=True) << At.T.select(sel, thunk)
    want = S.select(Co, Mo, accum, op, Ato, thunk, t0=True, complement=True, structure=False, replace=True)
    assert _coo_equal(C, *want.to_coo()), (op, thunk, "masked")


@pytest.mark.parametrize("dtype", [np.int64, np.float32, np.uint8])
@pytest.mark.parametrize("op,thunk", [("rowle", 40), ("rowgt", 10), ("valuegt", 0), ("valuene", 1), ("tril", -5), ("triu", 0)])
def test_vector_select_vs_oracle(gb, op, thunk, dtype):
    rng = np.random.default_rng(11)
    n = 100
    idx = np.unique(rng.integers(0, n, 60))
    v = H.random_values(rng, idx.size, dtype)
    u = gb.Vector.from_coo(idx, v, size=n)
    uo = S.SpMat.from_coo(idx, np.zeros_like(idx), v, n, 1)
    got = u.select(getattr(gb.select, op), thunk).new()
    want = S.select(S.SpMat(n, 1, np.dtype(dtype)), None, None, op, uo, thunk)
    wi, _, wx = want.to_coo()
    gi, gx = got.to_coo()
    assert np.array_equal(gi.astype(np.int64), wi) and np.array_equal(gx, wx.astype(gx.dtype)), (op, thunk)
    # under a structural mask with an accumulator
    w = gb.Vector.from_coo(idx[::2], v[::2], size=n)
    m = gb.Vector.from_coo(idx[::3], np.ones(idx[::3].size, bool), size=n)
    w(mask=m.S, accum=gb.binary.plus) << u.select(getattr(gb.select, op), thunk)
    wo = S.SpMat.from_coo(idx[::2], np.zeros_like(idx[::2]), v[::2], n, 1)
    mo = S.SpMat.from_coo(idx[::3], np.zeros_like(idx[::3]), np.ones(idx[::3].size, bool), n, 1)
    want = S.select(wo, mo, "plus", op, uo, thunk, structure=True)
    wi, _, wx = want.to_coo()
    gi, gx = w.to_coo()
    assert np.array_equal(gi.astype(np.int64), wi) and np.array_equal(gx, wx.astype(gx.dtype)), (op, thunk, "masked")


@pytest.mark.parametrize("monoid,dtype", [("plus", np.int64), ("plus", np.float64), ("min", np.int32), ("max", np.float32), ("times", np.int64),
                                          ("any", np.int64), ("any", np.float32), ("lor", np.bool_), ("land", np.bool_)])
def test_reduce_into_grb_scalar(gb, monoid, dtype):
    """GrB_{Vector,Matrix}_reduce_Monoid_Scalar (reference core/vector.py:1670, core/matrix.py:2750) incl. the ANY monoid on sparse
    inputs (round-1 advisor finding: empty lanes must not overwrite a real value) and the empty case"""
    rng = np.random.default_rng(3)
    op = getattr(gb.monoid, monoid)
    for n, k in ((10, 1), (1000, 7), (100000, 300)):
        idx = np.unique(rng.integers(0, n, k))
        v = H.random_values(rng, idx.size, dtype)
        if monoid == "times":
            v = np.where(v == 0, 1, v).astype(dtype) % 3 + 1
        u = gb.Vector.from_coo(idx, v, size=n)
        s = u.reduce(op).new()
        assert s.is_grbscalar and not s.is_empty
        if monoid == "any":
            assert s.value in set(v.tolist()), (n, s.value)
        else:
            want = S.reduce_scalar(monoid, S.SpMat.from_coo(idx, np.zeros_like(idx), v, n, 1))
            assert s.value == want.item(), (n, s.value, want)
        A = gb.Matrix.from_coo(idx // 10, idx % 10, v, nrows=n // 10 + 1, ncols=10)
        t = A.reduce_scalar(op).new()
        if monoid == "any":
            assert t.value in set(v.tolist())
        else:
            assert t.value == want.item()
    empty = gb.Vector(dtype, 50)
    assert empty.reduce(op).new().is_empty and empty.reduce(op).new().value is None
    assert gb.Matrix(dtype, 5, 5).reduce_scalar(op).new().is_empty
    if monoid == "plus":
        assert empty.reduce(op, allow_empty=False).new().value == 0
        s = gb.Scalar(dtype)
        s << u.reduce(op)
        first = s.value
        s(gb.binary.plus) << u.reduce(op)          # accumulate into the GrB_Scalar
        assert s.value == 2 * first
        s(gb.binary.plus) << empty.reduce(op)      # empty result + accumulator: unchanged
        assert s.value == 2 * first
        s << empty.reduce(op)                      # empty result, no accumulator: cleared
        assert s.is_empty
    if monoid == "min":
        assert empty.reduce(op, allow_empty=False).new().value == np.iinfo(np.int32).max


@pytest.mark.parametrize("dtype", [np.int64, np.float32])
@pytest.mark.parametrize("opname", ["plus", "minus", "times", "min", "first", "second"])
def test_apply_bind_typed_names_vs_oracle(gb, opname, dtype):
    """GrB_{Vector,Matrix}_apply_BinaryOp1st/2nd_<T> and ..._Scalar (reference core/vector.py:1477-1525, core/matrix.py:2472-2520)"""
    rng = np.random.default_rng(21)
    m, n = 30, 26
    r, c = H.random_coo(rng, m, n, 200)
    v = H.random_values(rng, r.size, dtype)
    A, Ao = gb.Matrix.from_coo(r, c, v, nrows=m, ncols=n), S.SpMat.from_coo(r, c, v, m, n)
    op = getattr(gb.binary, opname)
    sc = 3 if dtype == np.int64 else 2.5
    with gb.Recorder() as rec:
        left = A.apply(op, left=sc).new()
        right = A.apply(op, right=sc).new()
        tr = A.T.apply(op, left=sc).new()
        grb = A.apply(op, right=gb.Scalar.from_value(sc)).new()
    text = " ".join(rec.data)
    tname = "INT64" if dtype == np.int64 else "FP64"
    assert f"GrB_Matrix_apply_BinaryOp1st_{tname}(" in text and f"GrB_Matrix_apply_BinaryOp2nd_{tname}(" in text
    assert "GrB_Matrix_apply_BinaryOp2nd_Scalar(" in text
    D = np.result_type(dtype, np.int64 if dtype == np.int64 else np.float64)
    for got, kw in ((left, dict(scalar_first=True)), (right, dict(scalar_first=False)), (grb, dict(scalar_first=False))):
        want = S.apply(S.SpMat(m, n, D), None, None, opname, Ao, scalar=sc, **kw)
        assert _coo_equal(got, *want.to_coo()), (opname, kw)
    want = S.apply(S.SpMat(n, m, D), None, None, opname, Ao, scalar=sc, scalar_first=True, t0=True)
    assert _coo_equal(tr, *want.to_coo()), (opname, "transposed bind-1st")
    # vectors
    idx = np.unique(rng.integers(0, 80, 40))
    x = H.random_values(rng, idx.size, dtype)
    u = gb.Vector.from_coo(idx, x, size=80)
    uo = S.SpMat.from_coo(idx, np.zeros_like(idx), x, 80, 1)
    with gb.Recorder() as rec:
        l, r2 = u.apply(op, left=sc).new(), u.apply(op, right=sc).new()
    text = " ".join(rec.data)
    assert f"GrB_Vector_apply_BinaryOp1st_{tname}(" in text and f"GrB_Vector_apply_BinaryOp2nd_{tname}(" in text
    for got, first in ((l, True), (r2, False)):
        want = S.apply(S.SpMat(80, 1, D), None, None, opname, uo, scalar=sc, scalar_first=first)
        wi, _, wx = want.to_coo()
        gi, gx = got.to_coo()
        assert np.array_equal(gi.astype(np.int64), wi) and np.array_equal(gx, wx.astype(gx.dtype))


def test_whole_object_assign_and_setelement(gb):
    """GrB_Vector_assign / GrB_Matrix_assign with GrB_ALL (reference core/vector.py:1928, core/matrix.py:3300) and
    GrB_Matrix_setElement_<T>"""
    from graphblas_b200.base import call
    from graphblas_b200._lib import GrB_Index, lib

    rng = np.random.default_rng(8)
    n = 64
    ui, wi_, mi = (np.unique(rng.integers(0, n, k)) for k in (30, 25, 35))
    uv, wv = rng.integers(1, 9, ui.size), rng.integers(1, 9, wi_.size)
    u, w = gb.Vector.from_coo(ui, uv, size=n), gb.Vector.from_coo(wi_, wv, size=n)
    m = gb.Vector.from_coo(mi, np.ones(mi.size, bool), size=n)
    ALL = lib().GrB_ALL
    # w<m.S> += u : inside the mask accumulate / insert, outside untouched
    call("GrB_Vector_assign", [w, m, gb.binary.plus[gb.dtypes.INT64], u, ALL, GrB_Index(n), None])
    uo = S.SpMat.from_coo(ui, np.zeros_like(ui), uv, n, 1)
    wo = S.SpMat.from_coo(wi_, np.zeros_like(wi_), wv, n, 1)
    mo = S.SpMat.from_coo(mi, np.zeros_like(mi), np.ones(mi.size, bool), n, 1)
    want = S._write_back(wo, dict(uo.e), uo.dtype, mo, "plus", False, False, False)
    oi, _, ox = want.to_coo()
    gi, gx = w.to_coo()
    assert np.array_equal(gi.astype(np.int64), oi) and np.array_equal(gx, ox)
    # an empty GrB_Scalar assigned under a mask with replace deletes
    e = gb.Scalar(gb.dtypes.INT64)
    call("GrB_Vector_assign_Scalar", [w, m, None, e, ALL, GrB_Index(n), gb.base.descriptor_lookup(output_replace=True)])
    assert w.nvals == 0
    # matrix
    r, c = H.random_coo(rng, 12, 9, 40)
    A = gb.Matrix.from_coo(r, c, rng.integers(1, 9, r.size), nrows=12, ncols=9)
    C = gb.Matrix(gb.dtypes.INT64, 9, 12)
    call("GrB_Matrix_assign", [C, None, None, A, ALL, GrB_Index(9), ALL, GrB_Index(12), gb.base.descriptor_lookup(transpose_first=True)])
    assert C.isequal(A.T.new())
    # setElement: new entry, overwrite, bounds
    k0 = C.nvals
    free = [(i, j) for i in range(9) for j in range(12) if C[i, j].new().value is None][0]
    call("GrB_Matrix_setElement_INT64", [C, ctypes.c_int64(77), GrB_Index(free[0]), GrB_Index(free[1])])
    assert C.nvals == k0 + 1 and C[free].new() == 77
    call("GrB_Matrix_setElement_FP64", [C, ctypes.c_double(5.9), GrB_Index(free[0]), GrB_Index(free[1])])
    assert C.nvals == k0 + 1 and C[free].new() == 5
    with pytest.raises(gb.exceptions.InvalidIndex):
        call("GrB_Matrix_setElement_INT64", [C, ctypes.c_int64(1), GrB_Index(9), GrB_Index(0)])


def test_scipy_and_matrix_market_converters(gb, tmp_path):
    """graphblas_b200.io (reference graphblas/io/_scipy.py:8-119, io/_matrixmarket.py:8-140): scipy.sparse in every layout,
    duplicate coordinates with dup_op, Matrix Market round trip, a Vector as an n x 1 array"""
    import scipy.sparse as ss

    rng = np.random.default_rng(12)
    S_ = ss.random(60, 45, density=0.1, format="csr", random_state=3, dtype=np.float64)
    S_.data = np.round(S_.data * 10) + 1
    for fmt in ("csr", "csc", "coo", "lil"):
        A = gb.io.from_scipy_sparse(S_.asformat(fmt))
        back = gb.io.to_scipy_sparse(A, "csr")
        assert (back != S_).nnz == 0 and back.shape == S_.shape and A.dtype == gb.dtypes.FP64
    assert (gb.io.to_scipy_sparse(A, "csc") != S_.tocsc()).nnz == 0
    dup = ss.coo_array((np.array([1, 2, 5]), (np.array([0, 0, 1]), np.array([1, 1, 0]))), shape=(2, 2))
    with pytest.raises(ValueError, match="Duplicate indices found"):
        gb.io.from_scipy_sparse(dup)
    assert gb.io.from_scipy_sparse(dup, dup_op=gb.binary.plus).to_coo()[2].tolist() == [3, 5]
    assert gb.io.from_scipy_sparse(ss.csr_array((4, 7), dtype=np.int32)).nvals == 0
    path = str(tmp_path / "m.mtx")
    gb.io.mmwrite(path, A)
    B = gb.io.mmread(path)
    assert B.isequal(A)
    v = gb.Vector.from_coo([1, 4], [2.5, -1.0], size=6)
    col = gb.io.to_scipy_sparse(v)
    assert col.shape == (6, 1) and col[4, 0] == -1.0 and col.nnz == 2


@pytest.mark.parametrize("kind", ["vector", "matrix"])
def test_mask_algebra_and_matrix_scalar_assign(gb, kind):
    """`m1 & m2`, `m1 | m2` for every pair of mask kinds (reference core/mask.py:205-513 lists 16 + 16 recipes; here one
    set model, graphblas_b200/base.py Mask._combine), checked against numpy booleans by using the combined mask on a FULL source;
    and `C(mask, accum, replace) << scalar` for a Matrix (GrB_Matrix_assign_<T> over GrB_ALL x GrB_ALL) against a dense model."""
    rng = np.random.default_rng(21 if kind == "vector" else 22)
    shape = (60,) if kind == "vector" else (9, 8)
    size = int(np.prod(shape))

    def random_obj(seed_density):
        keep = rng.random(size) < seed_density
        vals = rng.integers(0, 3, size)            # zeros among the stored values: .S and .V differ
        flat = np.flatnonzero(keep)
        if kind == "vector":
            return gb.Vector.from_coo(flat, vals[flat], size=size), keep, vals != 0
        return gb.Matrix.from_coo(flat // shape[1], flat % shape[1], vals[flat], nrows=shape[0], ncols=shape[1]), keep, vals != 0

    def full_source():
        v = np.arange(1, size + 1)
        if kind == "vector":
            return gb.Vector.from_coo(np.arange(size), v, size=size)
        return gb.Matrix.from_coo(np.arange(size) // shape[1], np.arange(size) % shape[1], v, nrows=shape[0], ncols=shape[1])

    def positions(obj):
        if kind == "vector":
            return set(obj.to_coo()[0].tolist())
        I, J, _ = obj.to_coo()
        return set((I.astype(np.int64) * shape[1] + J.astype(np.int64)).tolist())

    a, a_keep, a_nz = random_obj(0.5)
    b, b_keep, b_nz = random_obj(0.4)
    src = full_source()
    kinds = {"S": lambda o, keep, nz: (o.S, keep), "V": lambda o, keep, nz: (o.V, keep & nz),
             "CS": lambda o, keep, nz: (~o.S, ~keep), "CV": lambda o, keep, nz: (~o.V, ~(keep & nz))}
    for k1, f1 in kinds.items():
        for k2, f2 in kinds.items():
            m1, t1 = f1(a, a_keep, a_nz)
            m2, t2 = f2(b, b_keep, b_nz)
            for name, comb, want in (("and", m1 & m2, t1 & t2), ("or", m1 | m2, t1 | t2)):
                out = type(src)(src.dtype, *shape) if kind == "matrix" else gb.Vector(src.dtype, size)
                out(comb) << src
                assert positions(out) == set(np.flatnonzero(want).tolist()), (kind, k1, k2, name)
    if kind == "matrix":
        # ---- scalar into a Matrix under a mask
        C0, c_keep, _ = random_obj(0.3)
        cI, cJ, cX = C0.to_coo()
        dense = np.zeros(size, dtype=np.int64); have = np.zeros(size, dtype=bool)
        dense[cI.astype(np.int64) * shape[1] + cJ.astype(np.int64)] = cX; have[cI.astype(np.int64) * shape[1] + cJ.astype(np.int64)] = True
        for label, mask, mtrue, accum, replace in (("S", a.S, a_keep, None, False), ("V", a.V, a_keep & a_nz, None, False),
                                                   ("S+accum", a.S, a_keep, "plus", False), ("V+replace", a.V, a_keep & a_nz, None, True)):
            C = C0.dup()
            kw = {"replace": True} if replace else {}
            if accum:
                C(mask, gb.binary.plus, **kw) << 7
            else:
                C(mask, **kw)[:, :] = 7
            want_v = dense.copy(); want_h = have.copy()
            if accum:
                want_v[mtrue] = np.where(have[mtrue], dense[mtrue] + 7, 7)
            else:
                want_v[mtrue] = 7
            want_h[mtrue] = True
            if replace:
                want_h[~mtrue] = False
            I, J, X = C.to_coo()
            flat = I.astype(np.int64) * shape[1] + J.astype(np.int64)
            assert set(flat.tolist()) == set(np.flatnonzero(want_h).tolist()), label
            assert np.array_equal(X, want_v[flat]), label
        with pytest.raises(NotImplementedError):
            C0.dup()(~a.S) << 1
        with pytest.raises(NotImplementedError):
            C0.dup() << 1


def test_vector_diag_and_matrix_vector_broadcast(gb):
    """Vector.diag (GrB_Matrix_diag, reference core/vector.py:605-628) and the broadcast recipes built on it: A.ewise_mult(v) =
    A (any).(op) diag(v) -- a caller of GrB_mxm -- and A.ewise_add(v) (reference core/matrix.py:62-75); against dense numpy models."""
    rng = np.random.default_rng(31)
    m, n = 13, 11
    keep = rng.random((m, n)) < 0.4
    Ad = rng.integers(1, 9, (m, n)).astype(np.int64)
    I, J = np.nonzero(keep)
    A = gb.Matrix.from_coo(I, J, Ad[I, J], nrows=m, ncols=n)
    vkeep = rng.random(n) < 0.6
    vd = rng.integers(1, 9, n).astype(np.int64)
    v = gb.Vector.from_coo(np.flatnonzero(vkeep), vd[vkeep], size=n)
    # ---- diag, every offset
    for k in (0, 2, -3):
        D = v.diag(k)
        assert D.shape == (n + abs(k), n + abs(k)) and D.nvals == int(vkeep.sum())
        di, dj, dx = D.to_coo()
        idx = np.flatnonzero(vkeep)
        assert np.array_equal(di, idx + max(-k, 0)) and np.array_equal(dj, idx + max(k, 0)) and np.array_equal(dx, vd[vkeep])
    # ---- ewise_mult broadcast: intersection of A's pattern with the columns v holds
    for opname, f in (("times", lambda a, b: a * b), ("plus", lambda a, b: a + b), ("first", lambda a, b: a), ("minus", lambda a, b: a - b)):
        C = A.ewise_mult(v, getattr(gb.binary, opname)).new()
        ci, cj, cx = C.to_coo()
        want = keep & vkeep[None, :]
        wi, wj = np.nonzero(want)
        assert np.array_equal(ci, wi) and np.array_equal(cj, wj), opname
        assert np.array_equal(cx, f(Ad, vd[None, :])[wi, wj]), opname
    # the transposed operand
    Ct = A.T.ewise_mult(gb.Vector.from_coo(np.arange(m), np.arange(1, m + 1), size=m), gb.binary.times).new()
    ti, tj, tx = Ct.to_coo()
    wi, wj = np.nonzero(keep.T)
    assert np.array_equal(ti, wi) and np.array_equal(tj, wj) and np.array_equal(tx, (Ad.T * np.arange(1, m + 1)[None, :])[wi, wj])
    # ---- ewise_add broadcast: union of A's pattern with v repeated in every row
    C = A.ewise_add(v, gb.binary.plus).new()
    ci, cj, cx = C.to_coo()
    want = keep | vkeep[None, :]
    wi, wj = np.nonzero(want)
    assert np.array_equal(ci, wi) and np.array_equal(cj, wj)
    dense = np.where(keep, Ad, 0) + np.where(vkeep[None, :], vd[None, :], 0)
    assert np.array_equal(cx, dense[wi, wj])
    with pytest.raises(gb.exceptions.DimensionMismatch):
        A.ewise_mult(gb.Vector(gb.dtypes.INT64, n + 1))


@pytest.mark.parametrize("dtype", [np.int32, np.float64, np.bool_, np.uint8], ids=lambda d: np.dtype(d).name)
def test_comparison_multiply_semirings(gb, dtype):
    """GxB_{LOR,LAND,ANY}_{EQ,NE,GT,LT,GE,LE}_<T>: the multiply compares (T x T -> BOOL), the monoid is logical.  The kernels run them
    in T (comparison as 1 / 0, LOR = max, LAND = min; csrc/gen_builtins.py) and the write-back casts to BOOL; checked for mxm, mxv
    and vxm against a dense numpy model over the shared indices k."""
    rng = np.random.default_rng(abs(hash(("cmp", np.dtype(dtype).name))) % 2**32)
    m, k, n = 17, 23, 14
    def rnd(r, c, dens):
        keep = rng.random((r, c)) < dens
        vals = (rng.integers(0, 2, (r, c)) if dtype == np.bool_ else rng.integers(0, 4, (r, c))).astype(dtype)
        return keep, vals
    ak, av = rnd(m, k, 0.4)
    bk, bv = rnd(k, n, 0.4)
    ai, aj = np.nonzero(ak); bi, bj = np.nonzero(bk)
    A = gb.Matrix.from_coo(ai, aj, av[ai, aj], nrows=m, ncols=k)
    B = gb.Matrix.from_coo(bi, bj, bv[bi, bj], nrows=k, ncols=n)
    cmpf = {"eq": np.equal, "ne": np.not_equal, "gt": np.greater, "lt": np.less, "ge": np.greater_equal, "le": np.less_equal}
    for add in ("lor", "land", "any"):
        for mul, f in cmpf.items():
            sr = getattr(gb.semiring, f"{add}_{mul}")
            C = A.mxm(B, sr).new()
            assert C.dtype == gb.dtypes.BOOL, (add, mul, C.dtype)
            ci, cj, cx = C.to_coo()
            both = ak[:, :, None] & bk[None, :, :]                      # (i, k, j): A(i,k) and B(k,j) both present
            prod = f(av[:, :, None], bv[None, :, :])
            pattern = both.any(axis=1)
            wi, wj = np.nonzero(pattern)
            assert np.array_equal(ci, wi) and np.array_equal(cj, wj), (add, mul)
            if add == "lor":
                want = (both & prod).any(axis=1)
                assert np.array_equal(cx, want[wi, wj]), (add, mul)
            elif add == "land":
                want = (~both | prod).all(axis=1)
                assert np.array_equal(cx, want[wi, wj]), (add, mul)
            else:   # any: some product's value
                lo, hi = (both & prod).any(axis=1), (~both | prod).all(axis=1)   # True is possible / False is impossible
                assert np.all((cx <= lo[wi, wj]) & (cx >= hi[wi, wj])), (add, mul)
    # vectors: mxv and vxm with lor_gt
    xk = rng.random(k) < 0.6
    xv = rng.integers(0, 4, k).astype(dtype) if dtype != np.bool_ else rng.integers(0, 2, k).astype(dtype)
    x = gb.Vector.from_coo(np.flatnonzero(xk), xv[xk], size=k)
    w = A.mxv(x, gb.semiring.lor_gt).new()
    wi_, wx = w.to_coo()
    both = ak & xk[None, :]
    assert np.array_equal(wi_, np.flatnonzero(both.any(axis=1)))
    assert np.array_equal(wx, (both & (av > xv[None, :])).any(axis=1)[wi_])
    w2 = x.vxm(B, gb.semiring.land_le).new()
    w2i, w2x = w2.to_coo()
    both = xk[:, None] & bk
    assert np.array_equal(w2i, np.flatnonzero(both.any(axis=0)))
    assert np.array_equal(w2x, (~both | (xv[:, None] <= bv)).all(axis=0)[w2i])


def test_dlpack_round_trip(gb):
    """DLPack interop -- the library's device arrays out as capsules (zero-copy) and back in (copied)"""
    import torch

    rng = np.random.default_rng(41)
    n = 500
    idx = np.unique(rng.integers(0, n, 200))
    v = gb.Vector.from_coo(idx, rng.random(idx.size), size=n)
    cv, cp = gb.cuda.vector_to_dlpack(v)
    tv, tp = torch.from_dlpack(cv), torch.from_dlpack(cp)
    assert tv.is_cuda and tp.dtype == torch.uint8 and int(tp.sum()) == idx.size
    w = gb.cuda.vector_from_dlpack(tv, tp)          # torch tensors speak __dlpack__
    assert w.isequal(v)
    r, c = H.random_coo(rng, 40, 30, 300)
    A = gb.Matrix.from_coo(r, c, rng.integers(1, 9, r.size).astype(np.int64), nrows=40, ncols=30)
    caps = gb.cuda.matrix_to_dlpack(A)
    B = gb.cuda.matrix_from_dlpack(*[torch.from_dlpack(x) for x in caps], 40, 30)
    assert B.isequal(A)
