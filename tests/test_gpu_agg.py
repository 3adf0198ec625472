"""Aggregators (reference graphblas/core/operator/agg.py): recipes over GrB_mxv / GrB_vxm / eWise / apply, compared with
numpy over the stored entries of every row / column / the whole object.  Integer results exact; floating point rtol 1e-12 (fp64
sums of < 100 terms).  Rows without entries must have no entry in the result."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gb():
    import graphblas_b200 as gb

    gb.init()
    return gb


def _np_agg(name, x):
    """numpy restatement over the stored values x (1-D, non-empty) of one row / the whole object"""
    xf = x.astype(np.float64)
    n = x.size
    with np.errstate(all="ignore"):
        return {
            "sum": lambda: x.sum(), "prod": lambda: x.prod(), "min": lambda: x.min(), "max": lambda: x.max(),
            "all": lambda: bool(np.all(x != 0)), "any": lambda: bool(np.any(x != 0)),
            "count": lambda: n, "count_nonzero": lambda: int(np.count_nonzero(x)), "count_zero": lambda: int(n - np.count_nonzero(x)),
            "sum_of_squares": lambda: (x * x).sum(), "sum_of_inverses": lambda: (1.0 / xf).sum(), "exists": lambda: 1,
            "hypot": lambda: np.sqrt((xf * xf).sum()), "logaddexp": lambda: np.log(np.exp(xf).sum()), "logaddexp2": lambda: np.log2(np.exp2(xf).sum()),
            "L1norm": lambda: np.abs(x).sum(), "Linfnorm": lambda: np.abs(x).max(), "mean": lambda: xf.sum() / n,
            "peak_to_peak": lambda: x.max() - x.min(), "varp": lambda: (xf * xf).sum() / n - (xf.sum() / n) ** 2,
            "vars": lambda: (xf * xf).sum() / (n - 1) - xf.sum() ** 2 / (n * (n - 1)) if n > 1 else np.nan,
            "stdp": lambda: np.sqrt(max((xf * xf).sum() / n - (xf.sum() / n) ** 2, 0.0)),
            "stds": lambda: np.sqrt((xf * xf).sum() / (n - 1) - xf.sum() ** 2 / (n * (n - 1))) if n > 1 else np.nan,
            "geometric_mean": lambda: xf.prod() ** (1.0 / n), "harmonic_mean": lambda: n / (1.0 / xf).sum(),
            "root_mean_square": lambda: np.sqrt((xf * xf).sum() / n),
        }[name]()


NAMES = ["sum", "prod", "min", "max", "all", "any", "count", "count_nonzero", "count_zero", "sum_of_squares", "sum_of_inverses", "exists",
         "hypot", "logaddexp", "logaddexp2", "L1norm", "Linfnorm", "mean", "peak_to_peak", "varp", "vars", "stdp", "stds", "geometric_mean",
         "harmonic_mean", "root_mean_square"]
INT_OK = {"sum", "prod", "min", "max", "all", "any", "count", "count_nonzero", "count_zero", "sum_of_squares", "exists", "L1norm", "Linfnorm",
          "mean", "peak_to_peak", "varp", "root_mean_square", "hypot"}


def _close(got, want, exact):
    if want is None or (isinstance(want, float) and np.isnan(want)):
        return got is None or (isinstance(got, float) and np.isnan(got))
    if got is None:
        return False
    if exact:
        return got == want
    return np.isclose(float(got), float(want), rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("dtype", [np.float64, np.int64], ids=["fp64", "int64"])
@pytest.mark.parametrize("name", NAMES)
def test_aggregators_vs_numpy(gb, name, dtype):
    if dtype == np.int64 and name not in INT_OK:
        pytest.skip("floating-point aggregator")
    rng = np.random.default_rng(abs(hash(("agg", name, np.dtype(dtype).name))) % 2**32)
    m, n = 37, 29
    r, c = H.random_coo(rng, m, n, 260)
    if dtype == np.int64:
        v = rng.integers(-3, 4, r.size).astype(np.int64)
        if name in ("prod",):
            v = rng.integers(1, 3, r.size).astype(np.int64)
    else:
        v = np.round(rng.uniform(0.5, 2.0, r.size), 3)
        if name in ("count_zero", "count_nonzero", "all", "any"):
            v[rng.random(r.size) < 0.3] = 0.0
    if name in ("count_zero", "count_nonzero") and dtype == np.int64:
        pass   # zeros already present among the integers
    A = gb.Matrix.from_coo(r, c, v, nrows=m, ncols=n)
    op = getattr(gb.agg, name)
    exact = dtype == np.int64 and name in {"sum", "prod", "min", "max", "all", "any", "count", "count_nonzero", "count_zero", "sum_of_squares",
                                           "exists", "L1norm", "Linfnorm", "peak_to_peak"} or name in ("count", "exists")
    # ---- row-wise and column-wise
    for axis, method, size in ((0, "reduce_rowwise", m), (1, "reduce_columnwise", n)):
        w = getattr(A, method)(op).new()
        idx, vals = w.to_coo()
        key = r if axis == 0 else c
        want_idx = np.unique(key)
        if name == "vars" or name == "stds":
            pass
        assert np.array_equal(np.sort(idx.astype(np.int64)), want_idx), (name, method)
        got = dict(zip(idx.tolist(), vals.tolist()))
        for i in want_idx.tolist():
            want = _np_agg(name, v[key == i])
            assert _close(got[i], want, exact), (name, method, i, got[i], want)
    # ---- the transposed operand goes through the same recipe
    wt = A.T.reduce_rowwise(op).new()
    wc = A.reduce_columnwise(op).new()
    it, vt = wt.to_coo()
    ic, vc = wc.to_coo()
    assert np.array_equal(it, ic) and np.allclose(vt.astype(np.float64), vc.astype(np.float64), rtol=1e-12, equal_nan=True)
    # ---- whole matrix -> scalar, vector -> scalar
    got = A.reduce_scalar(op).new().value
    assert _close(got, _np_agg(name, v), exact), (name, "matrix scalar", got, _np_agg(name, v))
    vi = np.unique(rng.integers(0, 500, 120))
    vv = v[: vi.size]
    u = gb.Vector.from_coo(vi, vv, size=500)
    got = u.reduce(op).new().value
    assert _close(got, _np_agg(name, vv), exact), (name, "vector scalar", got, _np_agg(name, vv))
    # ---- empty objects give an empty scalar / no entries
    assert gb.Vector(dtype, 10).reduce(op).new().value is None
    assert gb.Matrix(dtype, 4, 5).reduce_rowwise(op).new().nvals == 0
    assert gb.Matrix(dtype, 4, 5).reduce_scalar(op).new().value is None


def test_aggregator_under_mask_and_accum(gb):
    """`w(mask, accum) << A.reduce_rowwise(agg.x)`: the aggregated vector enters the ordinary write-back"""
    rng = np.random.default_rng(5)
    m, n = 20, 15
    r, c = H.random_coo(rng, m, n, 120)
    v = rng.integers(1, 9, r.size).astype(np.int64)
    A = gb.Matrix.from_coo(r, c, v, nrows=m, ncols=n)
    w = gb.Vector.from_coo(np.arange(m), np.full(m, 100, np.int64), size=m)
    mask = gb.Vector.from_coo(np.arange(0, m, 2), np.ones((m + 1) // 2, np.bool_), size=m)
    w(mask.S, gb.binary.plus) << A.reduce_rowwise(gb.agg.count)
    idx, vals = w.to_coo()
    cnt = np.bincount(r, minlength=m)
    want = np.where((np.arange(m) % 2 == 0) & (cnt > 0), 100 + cnt, 100)
    assert np.array_equal(idx, np.arange(m)) and np.array_equal(vals, want)


def test_float_unary_ops_and_pow(gb):
    """the GxB floating-point unary ops and GxB_POW the aggregator finalizers use, against numpy (bit-exact for sqrt / floor / ceil /
    trunc / rint which are correctly rounded; 2 ulp for exp / log family)"""
    rng = np.random.default_rng(9)
    x = rng.uniform(0.1, 9.0, 300)
    u = gb.Vector.from_coo(np.arange(300), x, size=300)
    for name, f, ulp in (("sqrt", np.sqrt, 0), ("floor", np.floor, 0), ("ceil", np.ceil, 0), ("trunc", np.trunc, 0), ("round", np.rint, 0),
                         ("exp", np.exp, 2), ("log", np.log, 2), ("exp2", np.exp2, 2), ("log2", np.log2, 2), ("log10", np.log10, 2), ("signum", np.sign, 0)):
        _, got = u.apply(getattr(gb.unary, name)).new().to_coo()
        want = f(x)
        if ulp == 0:
            assert np.array_equal(got, want), name
        else:
            assert np.all(np.abs(got - want) <= ulp * np.spacing(np.abs(want))), name
    _, got = u.apply(gb.binary.pow, right=2.0).new().to_coo()
    assert np.all(np.abs(got - x ** 2) <= 2 * np.spacing(x ** 2))
    xi = rng.integers(-5, 6, 300).astype(np.int64)
    ui = gb.Vector.from_coo(np.arange(300), xi, size=300)
    _, got = ui.apply(gb.binary.pow, right=3).new().to_coo()
    assert np.array_equal(got, xi ** 3)
