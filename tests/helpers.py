"""Shared helpers for the parity tests: seeded generators and GPU <-> oracle conversions."""
import numpy as np

from oracle import bigref as R
from oracle import semantics as S


def rmat_edges(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=42, rng=None):
    """Graph500-style R-MAT: per edge, `scale` quadrant picks; no permutation; dedup."""
    rng = np.random.default_rng(seed) if rng is None else rng
    n = 1 << scale
    m = edge_factor * n
    rows = np.zeros(m, dtype=np.int64)
    cols = np.zeros(m, dtype=np.int64)
    ab, abc = a + b, a + b + c
    for bit in range(scale):
        r = rng.random(m)
        rows += ((r >= ab).astype(np.int64)) << (scale - 1 - bit)
        cols += (((r >= a) & (r < ab)) | (r >= abc)).astype(np.int64) << (scale - 1 - bit)
    key = np.unique(rows * n + cols)
    return key // n, key % n, n


def random_coo(rng, nrows, ncols, nnz):
    key = np.unique(rng.integers(0, nrows * ncols, size=nnz))
    return key // ncols, key % ncols


def random_values(rng, n, dtype, small=True):
    dtype = np.dtype(dtype)
    if dtype == np.bool_:
        return rng.integers(0, 2, n).astype(bool)
    if dtype.kind == "f":
        return rng.integers(-8, 9, n).astype(dtype) if small else rng.random(n).astype(dtype)
    info = np.iinfo(dtype)
    return rng.integers(max(info.min, -20), min(info.max, 20) + 1, n).astype(dtype)


def gb_matrix(gb, r, c, v, nrows, ncols):
    return gb.Matrix.from_coo(r, c, v, nrows=nrows, ncols=ncols)


def gb_vector(gb, idx, vals, size):
    return gb.Vector.from_coo(idx, vals, size=size) if len(idx) else gb.Vector(vals.dtype, size)


def vec_equal(gbv, big: R.BigVec, rtol=0.0):
    idx, vals = gbv.to_coo()
    oi, ov = big.to_coo()
    if not np.array_equal(idx.astype(np.int64), oi):
        return False, f"pattern differs: {idx[:10]} vs {oi[:10]} (n {idx.size} vs {oi.size})"
    if rtol == 0.0:
        ok = np.array_equal(vals, ov.astype(vals.dtype))
    else:
        ok = np.allclose(vals.astype(np.float64), ov.astype(np.float64), rtol=rtol, atol=0)
    if not ok:
        bad = np.flatnonzero(vals != ov.astype(vals.dtype))[:5]
        return False, f"values differ at {idx[bad]}: {vals[bad]} vs {ov[bad]}"
    return True, ""


def mat_equal(gbm, big: R.BigMat, rtol=0.0):
    I, J, X = gbm.to_coo()
    oi, oj, ox = big.to_coo()
    if not (np.array_equal(I.astype(np.int64), oi) and np.array_equal(J.astype(np.int64), oj)):
        return False, f"pattern differs: nvals {I.size} vs {oi.size}"
    if rtol == 0.0:
        ok = np.array_equal(X, ox.astype(X.dtype))
    else:
        ok = np.allclose(X.astype(np.float64), ox.astype(np.float64), rtol=rtol, atol=0)
    if not ok:
        bad = np.flatnonzero(X != ox.astype(X.dtype))[:5]
        return False, f"values differ at {list(zip(I[bad], J[bad]))}: {X[bad]} vs {ox[bad]}"
    return True, ""
