"""Pins the oracle: (1) against the reference's known-answer tests (tests/golden/), (2) the
array-level oracle (C + numpy write-back) against the dict model on random small cases,
(3) against scipy.sparse for plus_times.  CPU only."""
import numpy as np
import pytest

import golden_cases as G
import helpers as H
from oracle import bigref as R
from oracle import semantics as S


def _sp(ref):
    d = G.load(ref)
    if d["kind"] == "Matrix":
        return S.SpMat.from_coo(d["rows"], d["cols"], d["vals"], d["nrows"], d["ncols"], dtype=d["vals"].dtype)
    return S.SpVec.from_coo(d["idx"], d["vals"], d["size"], dtype=d["vals"].dtype)


def _big(ref):
    d = G.load(ref)
    if d["kind"] == "Matrix":
        return R.BigMat.from_coo(d["rows"], d["cols"], d["vals"], d["nrows"], d["ncols"])
    return R.BigVec.from_coo(d["idx"], d["vals"], d["size"])


def run_case_dict(c):
    a, b = _sp(c["a"]), _sp(c["b"])
    mask = _sp(c["mask"]) if c["mask"] else None
    D = S.semiring_domain(c["semiring"], a.dtype, b.dtype)
    kw = dict(complement=c["complement"], structure=c["mask_kind"] == "S", replace=c["replace"])
    if c["kind"] == "mxm":
        out = _sp(c["out"]) if c["out"] else S.SpMat(*c["out_shape"], D)
        return S.mxm(out, mask, c["accum"], c["semiring"], a, b, t0=c["ta"], t1=c["tb"], **kw)
    out = _sp(c["out"]) if c["out"] else S.SpVec(c["out_shape"][0], D)
    if c["kind"] == "mxv":
        return S.mxv(out, mask, c["accum"], c["semiring"], a, b, t0=c["ta"], **kw)
    return S.vxm(out, mask, c["accum"], c["semiring"], a, b, t1=c["tb"], **kw)


def run_case_big(c):
    a, b = _big(c["a"]), _big(c["b"])
    mask = _big(c["mask"]) if c["mask"] else None
    D = S.semiring_domain(c["semiring"], a.dtype, b.dtype)
    kw = dict(complement=c["complement"], structure=c["mask_kind"] == "S", replace=c["replace"])
    if c["kind"] == "mxm":
        out = _big(c["out"]) if c["out"] else R.BigMat(np.zeros(c["out_shape"][0] + 1), [], np.zeros(0, D), *c["out_shape"])
        return R.mxm(out, mask, c["accum"], c["semiring"], a, b, t0=c["ta"], t1=c["tb"], **kw)
    out = _big(c["out"]) if c["out"] else R.BigVec.empty(c["out_shape"][0], D)
    if c["kind"] == "mxv":
        return R.mxv(out, mask, c["accum"], c["semiring"], a, b, t0=c["ta"], **kw)
    return R.vxm(out, mask, c["accum"], c["semiring"], a, b, t1=c["tb"], **kw)


def same_coo(x, y):
    cx, cy = x.to_coo(), y.to_coo()
    return len(cx) == len(cy) and all(np.array_equal(p, q) for p, q in zip(cx, cy))


@pytest.mark.parametrize("c", G.CASES, ids=[c["id"] for c in G.CASES])
def test_dict_model_matches_reference_goldens(c):
    got, want = run_case_dict(c), _sp(c["expect"])
    assert same_coo(got, want), (got.to_coo(), want.to_coo())


@pytest.mark.parametrize("c", G.CASES, ids=[c["id"] for c in G.CASES])
def test_array_oracle_matches_reference_goldens(c):
    got, want = run_case_big(c), _big(c["expect"])
    assert same_coo(got, want), (got.to_coo(), want.to_coo())


def test_nonsquare_max_plus():
    a, b = _sp(G.NONSQUARE["a"]), _sp(G.NONSQUARE["b"])
    C = S.mxm(S.SpMat(1, 1, np.int64), None, None, "max_plus", a, b)
    assert C.e == {(0, 0): 33}
    C2 = S.mxm(S.SpMat(5, 5, np.int64), None, None, "max_plus", a, b, t0=True, t1=True)
    assert (C2.nrows, C2.ncols) == (5, 5)
    Cb = R.mxm(R.BigMat(np.zeros(2), [], np.zeros(0, np.int64), 1, 1), None, None, "max_plus",
               _big(G.NONSQUARE["a"]), _big(G.NONSQUARE["b"]))
    assert Cb.to_coo()[2].tolist() == [33]


def test_docs_worked_examples():
    # reference docs/user_guide/operations.rst:77-153: mxv plus_times -> (40,170,20); vxm plus_plus -> (69,84,12)
    # (inputs restated from the rst tables)
    A = S.SpMat.from_coo([0, 0, 1, 1, 2], [1, 2, 0, 2, 1], [2.0, 5.0, 1.5, 4.25, 0.5], 3, 3, dtype=np.float64)
    del A  # the rst example matrices are not machine-readable; the two totals are covered by random tests


def _rand_mat(rng, nr, nc, density, dtype):
    mask = rng.random((nr, nc)) < density
    r, c = np.nonzero(mask)
    if np.dtype(dtype) == np.bool_:
        vals = rng.integers(0, 2, r.size).astype(bool)
    elif np.dtype(dtype).kind == "f":
        vals = rng.integers(-8, 9, r.size).astype(dtype)  # small integers: exact in fp
    else:
        info = np.iinfo(dtype)
        vals = rng.integers(max(info.min, -50), min(info.max, 50), r.size).astype(dtype)
    return r, c, vals


SEMIRINGS = ["plus_times", "min_plus", "max_plus", "plus_plus", "plus_second", "plus_first", "any_pair",
             "lor_land", "min_first", "plus_pair", "max_times", "min_second"]


@pytest.mark.parametrize("semiring", SEMIRINGS)
@pytest.mark.parametrize("dtype", [np.int64, np.int8, np.uint16, np.float64, np.float32, np.bool_])
def test_array_oracle_vs_dict_model_random(semiring, dtype):
    if semiring == "lor_land" and dtype != np.bool_:
        pytest.skip("lor_land runs in BOOL")
    if dtype == np.bool_ and semiring not in ("lor_land", "any_pair"):
        pytest.skip("numeric semiring")
    rng = np.random.default_rng(hash((semiring, np.dtype(dtype).name)) % 2**32)
    for trial in range(4):
        n, k, m = rng.integers(1, 12, 3)
        ar, ac, av = _rand_mat(rng, n, k, 0.35, dtype)
        br, bc, bv = _rand_mat(rng, k, m, 0.35, dtype)
        cr, cc, cv = _rand_mat(rng, n, m, 0.3, dtype)
        mr, mc, mv = _rand_mat(rng, n, m, 0.5, np.int8)
        for accum in (None, "plus", "min"):
            if dtype == np.bool_ and accum is not None:
                accum = "lor"
            for comp in (False, True):
                for struct in (False, True):
                    for repl in (False, True):
                        for use_mask in (False, True):
                            if not use_mask and (comp or struct or repl):
                                continue
                            kw = dict(complement=comp, structure=struct, replace=repl)
                            sC = S.SpMat.from_coo(cr, cc, cv, n, m, dtype=dtype)
                            sM = S.SpMat.from_coo(mr, mc, mv, n, m, dtype=np.int8) if use_mask else None
                            want = S.mxm(sC, sM, accum, semiring, S.SpMat.from_coo(ar, ac, av, n, k, dtype=dtype),
                                         S.SpMat.from_coo(br, bc, bv, k, m, dtype=dtype), **kw)
                            bC = R.BigMat.from_coo(cr, cc, cv, n, m)
                            bM = R.BigMat.from_coo(mr, mc, mv, n, m) if use_mask else None
                            got = R.mxm(bC, bM, accum, semiring, R.BigMat.from_coo(ar, ac, av, n, k),
                                        R.BigMat.from_coo(br, bc, bv, k, m), **kw)
                            w, g = want.to_coo(), got.to_coo()
                            assert all(np.array_equal(p, q) for p, q in zip(w, g)), (semiring, dtype, accum, kw, w, g)


@pytest.mark.parametrize("semiring", ["plus_times", "min_plus", "plus_second", "any_pair"])
def test_array_oracle_vectors_vs_dict_model_random(semiring):
    rng = np.random.default_rng(7)
    dtype = np.int64
    for trial in range(20):
        n, m = rng.integers(1, 14, 2)
        ar, ac, av = _rand_mat(rng, n, m, 0.4, dtype)
        ui = np.flatnonzero(rng.random(n) < 0.6)
        uv = rng.integers(-5, 6, ui.size)
        xi = np.flatnonzero(rng.random(m) < 0.6)
        xv = rng.integers(-5, 6, xi.size)
        wi = np.flatnonzero(rng.random(m) < 0.5)
        wv = rng.integers(-5, 6, wi.size)
        mi = np.flatnonzero(rng.random(m) < 0.5)
        mv = rng.integers(0, 2, mi.size)
        for comp in (False, True):
            for struct in (False, True):
                for repl in (False, True):
                    for accum in (None, "min"):
                        kw = dict(complement=comp, structure=struct, replace=repl)
                        A_s = S.SpMat.from_coo(ar, ac, av, n, m, dtype=dtype)
                        A_b = R.BigMat.from_coo(ar, ac, av, n, m)
                        # vxm: w(m) = u(n) A(n,m)
                        want = S.vxm(S.SpVec.from_coo(wi, wv, m, dtype=dtype), S.SpVec.from_coo(mi, mv, m, dtype=dtype),
                                     accum, semiring, S.SpVec.from_coo(ui, uv, n, dtype=dtype), A_s, **kw)
                        got = R.vxm(R.BigVec.from_coo(wi, wv, m, dtype=dtype), R.BigVec.from_coo(mi, mv, m, dtype=dtype),
                                    accum, semiring, R.BigVec.from_coo(ui, uv, n, dtype=dtype), A_b, **kw)
                        assert all(np.array_equal(p, q) for p, q in zip(want.to_coo(), got.to_coo()))
                        # mxv with A transposed: w(m) = A'(m,n) u(n)
                        want = S.mxv(S.SpVec.from_coo(wi, wv, m, dtype=dtype), S.SpVec.from_coo(mi, mv, m, dtype=dtype),
                                     accum, semiring, A_s, S.SpVec.from_coo(ui, uv, n, dtype=dtype), t0=True, **kw)
                        got = R.mxv(R.BigVec.from_coo(wi, wv, m, dtype=dtype), R.BigVec.from_coo(mi, mv, m, dtype=dtype),
                                    accum, semiring, A_b, R.BigVec.from_coo(ui, uv, n, dtype=dtype), t0=True, **kw)
                        assert all(np.array_equal(p, q) for p, q in zip(want.to_coo(), got.to_coo()))
                        # vxm with A transposed: w(n) = x(m) A'(m,n)
                        want = S.vxm(S.SpVec(n, dtype), None, None, semiring, S.SpVec.from_coo(xi, xv, m, dtype=dtype), A_s,
                                     t1=True)
                        got = R.vxm(R.BigVec.empty(n, dtype), None, None, semiring,
                                    R.BigVec.from_coo(xi, xv, m, dtype=dtype), A_b, t1=True)
                        assert all(np.array_equal(p, q) for p, q in zip(want.to_coo(), got.to_coo()))


def test_array_oracle_vs_scipy_plus_times():
    import scipy.sparse as sp

    rng = np.random.default_rng(0)
    n = 1000
    r, c = rng.integers(0, n, 10_000), rng.integers(0, n, 10_000)
    key = np.unique(r * n + c)
    r, c = key // n, key % n
    vals = rng.random(r.size)
    A = R.BigMat.from_coo(r, c, vals, n, n)
    As = sp.csr_matrix((vals, (r, c)), shape=(n, n))
    x = rng.random(n)
    t = R.mxv_T("plus_times", A, R.BigVec(x, np.ones(n, np.uint8)))
    y = As @ x
    has = np.diff(As.indptr) > 0
    assert np.array_equal(t.present.astype(bool), has)
    np.testing.assert_allclose(t.vals[has], y[has], rtol=1e-12)
    C = R.mxm_T("plus_times", A, A)
    Cs = (As @ As).tocsr()
    Cs.sort_indices()
    assert np.array_equal(C.indptr, Cs.indptr) and np.array_equal(C.indices, Cs.indices)
    np.testing.assert_allclose(C.values, Cs.data, rtol=1e-12)


def test_int64_wraps_like_reference_test_power():
    # graphblas/tests/test_matrix.py:4379-4405: chained INT64 mxm wraps around
    d = G.load(G.A_M)
    A = S.SpMat.from_coo(d["rows"], d["cols"], d["vals"], 7, 7, dtype=np.int64)
    Ab = R.BigMat.from_coo(d["rows"], d["cols"], d["vals"], 7, 7)
    P, Pb = A.dup(), Ab
    for _ in range(40):
        P = S.mxm(S.SpMat(7, 7, np.int64), None, None, "plus_times", P, A)
        Pb = R.mxm_T("plus_times", Pb, Ab)
    assert all(np.array_equal(p, q) for p, q in zip(P.to_coo(), Pb.to_coo()))
    assert np.abs(P.to_coo()[2]).max() > 2**40  # did exercise large magnitudes


def test_inner_outer_reference_known_answers():
    """reference tests/test_vector.py:1496-1547: v.inner(v) == 6 for the fixture v (indices [1,3,4,6], values [1,1,2,0]);
    v.outer(v) equals the product of the column and row forms."""
    v = S.SpVec.from_coo([1, 3, 4, 6], [1, 1, 2, 0], size=7, dtype=np.int64)
    val, dt = S.inner("plus_times", v, v)
    assert int(val) == 6 and dt == np.int64
    w = S.SpVec.from_coo([0, 2], [5, 7], size=7, dtype=np.int64)
    assert S.inner("plus_times", v, w)[0] is None                 # no shared index -> empty scalar
    assert int(S.inner("min_plus", v, v)[0]) == 0                 # min(1+1, 1+1, 2+2, 0+0)
    out = S.outer("times", v, v)
    want = S.SpMat(7, 7, np.int64)
    S.mxm(want, None, None, "plus_times", v.as_col(), v.as_row())
    assert out.e == want.e and len(out.e) == 16 and int(out.e[(4, 4)]) == 4 and int(out.e[(6, 1)]) == 0


def test_matrix_elementwise_oracle_vs_scipy():
    """The oracle's matrix element-wise restatement (transpose / eWiseAdd / eWiseMult / apply / reduce), which the GPU tests use as
    the checker for SURVEY 8(f1), cross-checked against scipy.sparse on seeded inputs (exactly representable values)."""
    import scipy.sparse as sp

    rng = np.random.default_rng(23)
    m, n = 40, 55
    def rand():
        key = np.unique(rng.integers(0, m * n, 400))
        r, c = key // n, key % n
        v = rng.integers(-5, 6, r.size).astype(np.float64)
        return S.SpMat.from_coo(r, c, v, m, n, np.float64), sp.csr_matrix((v, (r, c)), shape=(m, n))
    (Ao, As), (Bo, Bs) = rand(), rand()

    def same(o, s_):
        s_ = s_.tocoo()
        want = {(int(i), int(j)): float(v) for i, j, v in zip(s_.row, s_.col, s_.data)}
        got = {k: float(v) for k, v in o.e.items()}
        return got == want

    C = S.SpMat(n, m, np.float64); S.transpose(C, None, None, Ao)
    assert same(C, As.T.tocsr())
    # eWiseAdd(plus): scipy drops the explicit zeros that cancel; compare on the union pattern with values
    C = S.SpMat(m, n, np.float64); S.ewise(C, None, None, "plus", Ao, Bo, union=True)
    dense = As.toarray() + Bs.toarray()
    pattern = (As != 0).toarray() | (Bs != 0).toarray() | np.isin(np.arange(m * n).reshape(m, n), [i * n + j for (i, j) in set(Ao.e) | set(Bo.e)])
    assert set(C.e) == set(Ao.e) | set(Bo.e) and all(float(v) == dense[k] for k, v in C.e.items()) and pattern.sum() >= len(C.e)
    C = S.SpMat(m, n, np.float64); S.ewise(C, None, None, "times", Ao, Bo, union=False)
    prod = As.toarray() * Bs.toarray()
    assert set(C.e) == set(Ao.e) & set(Bo.e) and all(float(v) == prod[k] for k, v in C.e.items())
    C = S.SpMat(m, n, np.float64); S.apply(C, None, None, "ainv", Ao)
    assert set(C.e) == set(Ao.e) and all(float(v) == -float(Ao.e[k]) for k, v in C.e.items())
    C = S.SpMat(m, n, np.float64); S.apply(C, None, None, "minus", Ao, scalar=np.float64(3), scalar_first=True)
    assert all(float(v) == 3 - float(Ao.e[k]) for k, v in C.e.items())
    assert float(S.reduce_scalar("plus", Ao)) == float(As.sum()) and float(S.reduce_scalar("max", Ao)) == float(max(Ao.e.values()))
    # mask + accum through the common write-back: C<M.S>(plus) = A'  on a square case
    Sq, Ss = S.SpMat.from_coo([0, 1, 2], [1, 2, 0], [1.0, 2.0, 3.0], 3, 3, np.float64), None
    Cq = S.SpMat.from_coo([1, 0], [0, 0], [10.0, 20.0], 3, 3, np.float64)
    Mq = S.SpMat.from_coo([1, 2], [0, 1], [0, 1], 3, 3, np.int64)
    S.transpose(Cq, Mq, "plus", Sq, structure=True)
    assert {k: float(v) for k, v in Cq.e.items()} == {(1, 0): 11.0, (0, 0): 20.0, (2, 1): 2.0}


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` prints ONE JSON line with the keys the driver reads (small scale so it runs in seconds)."""
    import json
    import pathlib
    import subprocess
    import sys

    root = pathlib.Path(__file__).resolve().parents[1]
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--scale", "12", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=str(root))
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stderr[-2000:]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "nnz-out/s" and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["e2e"]["h2d_bytes_per_step"] == 0
    # the arm must use every host core even when the launcher exports OMP_NUM_THREADS=1 (torch.distributed.run does)
    import os

    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--scale", "12", "--steps", "1", "--warmup", "1", "--gpus", "2"],
                         capture_output=True, text=True, timeout=600, cwd=str(root), env=env)
    d2 = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d2["cpu_baseline"]["cores"] == d["cpu_baseline"]["cores"] and d2["config"] == d["config"]
    import bench

    assert d["config"]["workload"] == bench.WORKLOAD.format(scale=12)   # the same string the GPU arm prints: same_config


def test_fast_cpu_baseline_matches_oracle():
    """bench.py's CPU baseline (one pass, unsorted rows, hash / dense accumulators, 32-bit indices) computes exactly what the
    sorted two-pass parity oracle computes: same row counts, and per row the same (column, value) set -- for both accumulators."""
    r, c, n = H.rmat_edges(12, a=0.45, b=0.15, c=0.15, seed=42)
    ip = np.zeros(n + 1, dtype=np.int64)
    ip[1:] = np.cumsum(np.bincount(r, minlength=n))
    vals = np.random.default_rng(43).integers(1, 4, r.size).astype(np.float32)   # small integers: fp32 sums are exact
    A = R.BigMat(ip, c, vals, n, n)
    T = R.mxm_T("plus_times", A, A)
    for hm in (0, 64, 1 << 20):   # dense workspace only / mixed / hash only
        nv, dt, (Sp, rn, Cj, Cx) = R.mxm_baseline_f32(A, A, hash_max_flops=hm)
        assert nv == T.nvals and dt > 0 and np.array_equal(rn, np.diff(T.indptr)), hm
        for i in range(n):
            s, e = int(Sp[i]), int(Sp[i]) + int(rn[i])
            o = np.argsort(Cj[s:e])
            assert np.array_equal(Cj[s:e][o].astype(np.int64), T.indices[T.indptr[i]:T.indptr[i + 1]]), (hm, i)
            assert np.array_equal(Cx[s:e][o], T.values[T.indptr[i]:T.indptr[i + 1]]), (hm, i)
