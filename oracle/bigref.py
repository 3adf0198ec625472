"""ORACLE (test infrastructure only -- never imported by the product path).

Array-level restatement of GrB_mxm / GrB_mxv / GrB_vxm for inputs too large for the
dict model in oracle/semantics.py.  The multiply T = A (+).(x) B runs in the C
restatement oracle/grb_oracle.c (OpenMP); the mask / accum / replace write-back
is vectorised numpy and follows exactly the block in SURVEY.md section 8(c):

    Z = accum is None ? T : accum(C, T) on C u T
    m = M is None ? true : (structure ? present : present and bool(value)); complement flips
    m(i,j)  -> C(i,j) := Z(i,j) if present else delete
    !m(i,j) -> replace ? delete : keep

Reference call sites whose behaviour this pins: graphblas/core/matrix.py:2252-2259 (GrB_mxv),
:2319-2328 (GrB_mxm), graphblas/core/vector.py:1368-1375 (GrB_vxm), core/base.py:458-503
(mask flags / descriptor), core/base.py:254-260 (accum typed by the output dtype).

Validated in tests/test_oracle.py against (a) the reference's golden vectors and
(b) the dict model on random small cases, (c) scipy.sparse for plus_times.
"""
from __future__ import annotations

import ctypes
import os
import pathlib
import subprocess

import numpy as np

from . import semantics as S

_HERE = pathlib.Path(__file__).resolve().parent
_LIB = None

TYPE_CODE = {
    np.dtype(np.bool_): 0, np.dtype(np.int8): 1, np.dtype(np.int16): 2, np.dtype(np.int32): 3,
    np.dtype(np.int64): 4, np.dtype(np.uint8): 5, np.dtype(np.uint16): 6, np.dtype(np.uint32): 7,
    np.dtype(np.uint64): 8, np.dtype(np.float32): 9, np.dtype(np.float64): 10,
}
OP_CODE = {"first": 1, "second": 2, "pair": 3, "oneb": 3, "plus": 4, "minus": 5, "times": 6, "div": 7,
           "min": 8, "max": 9, "lor": 10, "land": 11, "lxor": 12, "any": 13}


def build(force=False):
    so = _HERE / "_build" / "libgrb_oracle.so"
    src = _HERE / "grb_oracle.c"
    if force or not so.exists() or (src.exists() and so.stat().st_mtime < src.stat().st_mtime):
        subprocess.check_call(["make", "-C", str(_HERE)], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(str(build()))
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


# --------------------------------------------------------------------------- containers
class BigMat:
    """CSR with int64 indptr/indices, sorted columns, no duplicates."""

    def __init__(self, indptr, indices, values, nrows, ncols):
        self.indptr, self.indices = _i64(indptr), _i64(indices)
        self.values = np.ascontiguousarray(values)
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.dtype = self.values.dtype

    @classmethod
    def from_coo(cls, rows, cols, vals, nrows, ncols, dtype=None):
        import scipy.sparse as sp

        vals = np.asarray(vals, dtype=dtype)
        # scipy would SUM duplicates; callers pass deduplicated coordinates
        order = np.lexsort((np.asarray(cols), np.asarray(rows)))
        rows = np.asarray(rows, dtype=np.int64)[order]
        cols = np.asarray(cols, dtype=np.int64)[order]
        vals = vals[order]
        indptr = np.zeros(nrows + 1, dtype=np.int64)
        np.add.at(indptr, rows + 1, 1)
        np.cumsum(indptr, out=indptr)
        return cls(indptr, cols, vals, nrows, ncols)

    def T(self):
        rows = np.repeat(np.arange(self.nrows, dtype=np.int64), np.diff(self.indptr))
        return BigMat.from_coo(self.indices, rows, self.values, self.ncols, self.nrows)

    def to_coo(self):
        rows = np.repeat(np.arange(self.nrows, dtype=np.int64), np.diff(self.indptr))
        return rows, self.indices.copy(), self.values.copy()

    @property
    def nvals(self):
        return int(self.indptr[-1])


class BigVec:
    """Dense value array + byte presence array (the same layout the device uses)."""

    def __init__(self, vals, present):
        self.vals = np.ascontiguousarray(vals)
        self.present = np.ascontiguousarray(present, dtype=np.uint8)
        self.size = self.vals.shape[0]
        self.dtype = self.vals.dtype

    @classmethod
    def from_coo(cls, idx, vals, size, dtype=None):
        vals = np.asarray(vals, dtype=dtype)
        v = np.zeros(size, dtype=vals.dtype)
        p = np.zeros(size, dtype=np.uint8)
        v[np.asarray(idx, dtype=np.int64)] = vals
        p[np.asarray(idx, dtype=np.int64)] = 1
        return cls(v, p)

    @classmethod
    def empty(cls, size, dtype):
        return cls(np.zeros(size, dtype=dtype), np.zeros(size, dtype=np.uint8))

    def to_coo(self):
        idx = np.flatnonzero(self.present).astype(np.int64)
        return idx, self.vals[idx]

    @property
    def nvals(self):
        return int(self.present.sum())


# --------------------------------------------------------------------------- multiply
def _vals_in_domain(vals, D, mul, side):
    """Cast to the semiring domain; None when the multiply never reads this side."""
    return np.ascontiguousarray(vals.astype(D, copy=False))


def mxv_T(semiring, A: BigMat, x: BigVec, flip=False):
    """t = A (+).(x) x  (row-wise pull).  flip=True computes mul(x_k, a_ik) (vxm with A')."""
    add, mul = S.split_semiring(semiring)
    D = S.semiring_domain(semiring, x.dtype, A.dtype) if flip else S.semiring_domain(semiring, A.dtype, x.dtype)
    Ax = _vals_in_domain(A.values, D, mul, 0)
    xv = _vals_in_domain(x.vals, D, mul, 1)
    tv = np.zeros(A.nrows, dtype=D)
    tp = np.zeros(A.nrows, dtype=np.uint8)
    xp = None if x.present.all() else x.present
    rc = lib().oracle_mxv(OP_CODE[add], OP_CODE[mul], TYPE_CODE[np.dtype(D)], int(flip),
                          ctypes.c_int64(A.nrows), _p(A.indptr), _p(A.indices), _p(Ax), _p(xv), _p(xp),
                          _p(tv), _p(tp))
    assert rc == 0, rc
    return BigVec(tv, tp)


def vxm_push_T(semiring, u: BigVec, A: BigMat, flip=False):
    """t = u (+).(x) A (scatter over the rows of A selected by u).  flip=True: mul(a_ij, u_i)."""
    add, mul = S.split_semiring(semiring)
    D = S.semiring_domain(semiring, A.dtype, u.dtype) if flip else S.semiring_domain(semiring, u.dtype, A.dtype)
    Ax = _vals_in_domain(A.values, D, mul, 1)
    uv = _vals_in_domain(u.vals, D, mul, 0)
    tv = np.zeros(A.ncols, dtype=D)
    tp = np.zeros(A.ncols, dtype=np.uint8)
    up = None if u.present.all() else u.present
    rc = lib().oracle_vxm_push(OP_CODE[add], OP_CODE[mul], TYPE_CODE[np.dtype(D)], int(flip),
                               ctypes.c_int64(A.nrows), ctypes.c_int64(A.ncols), _p(A.indptr), _p(A.indices),
                               _p(Ax), _p(uv), _p(up), _p(tv), _p(tp))
    assert rc == 0, rc
    return BigVec(tv, tp)


def mxm_T(semiring, A: BigMat, B: BigMat, rows=None):
    """T = A (+).(x) B for rows [r0, r1) of A (default all): sorted CSR."""
    add, mul = S.split_semiring(semiring)
    D = S.semiring_domain(semiring, A.dtype, B.dtype)
    r0, r1 = (0, A.nrows) if rows is None else rows
    Ax = _vals_in_domain(A.values, D, mul, 0)
    Bx = _vals_in_domain(B.values, D, mul, 1)
    row_nnz = np.zeros(r1 - r0, dtype=np.int64)
    L = lib()
    rc = L.oracle_mxm_count(ctypes.c_int64(r0), ctypes.c_int64(r1), ctypes.c_int64(B.ncols), _p(A.indptr),
                            _p(A.indices), _p(B.indptr), _p(B.indices), _p(row_nnz))
    assert rc == 0, rc
    Cp = np.zeros(r1 - r0 + 1, dtype=np.int64)
    np.cumsum(row_nnz, out=Cp[1:])
    Cj = np.empty(int(Cp[-1]), dtype=np.int64)
    Cx = np.empty(int(Cp[-1]), dtype=D)
    rc = L.oracle_mxm_fill(OP_CODE[add], OP_CODE[mul], TYPE_CODE[np.dtype(D)], ctypes.c_int64(r0),
                           ctypes.c_int64(r1), ctypes.c_int64(B.ncols), _p(A.indptr), _p(A.indices), _p(Ax),
                           _p(B.indptr), _p(B.indices), _p(Bx), _p(Cp), _p(Cj), _p(Cx))
    assert rc == 0, rc
    return BigMat(Cp, Cj, Cx, r1 - r0, B.ncols)


def mxm_baseline_f32(A: BigMat, B: BigMat, *, hash_max_flops=4096, indices32=None):
    """The CPU BASELINE of bench.py (not the parity oracle): plus_times fp32 A*B in one numeric pass, unsorted rows, 32-bit
    indices, hash accumulator for rows with <= hash_max_flops products and the dense Gustavson workspace above (see the C
    side).  Returns (nvals, seconds of the multiply itself, staging arrays) -- the timed region covers the per-row flop
    bound, its prefix and the numeric pass; converting the operands to 32-bit indices is input preparation, as importing a
    matrix is for the library under test.  `indices32` = (Aj32, Bj32) reuses converted index arrays."""
    import time

    L = lib()
    Aj = A.indices.astype(np.int32) if indices32 is None else indices32[0]
    Bj = B.indices.astype(np.int32) if indices32 is None else indices32[1]
    Ax, Bx = np.ascontiguousarray(A.values, dtype=np.float32), np.ascontiguousarray(B.values, dtype=np.float32)
    m = A.nrows
    t0 = time.perf_counter()
    flops = np.empty(m + 1, dtype=np.int64)
    L.oracle_row_flops32(ctypes.c_int64(0), ctypes.c_int64(m), _p(A.indptr), _p(Aj), _p(B.indptr), _p(flops))
    flops[m] = 0
    Sp = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(flops[:m], out=Sp[1:])
    total = int(Sp[-1])
    Cj = np.empty(max(total, 1), dtype=np.int32)
    Cx = np.empty(max(total, 1), dtype=np.float32)
    row_nnz = np.empty(m, dtype=np.int64)
    rc = L.oracle_mxm_baseline_f32(ctypes.c_int64(0), ctypes.c_int64(m), ctypes.c_int64(B.ncols), _p(A.indptr), _p(Aj), _p(Ax),
                                   _p(B.indptr), _p(Bj), _p(Bx), _p(Sp), _p(Cj), _p(Cx), _p(row_nnz), ctypes.c_int64(hash_max_flops))
    dt = time.perf_counter() - t0
    assert rc == 0, rc
    return int(row_nnz.sum()), dt, (Sp, row_nnz, Cj, Cx)


# --------------------------------------------------------------------------- write-back
def _accum_np(name, c, t):
    with np.errstate(all="ignore"):
        if name == "plus":
            return c + t
        if name == "minus":
            return c - t
        if name == "times":
            return c * t
        if name == "min":
            return np.minimum(c, t)
        if name == "max":
            return np.maximum(c, t)
        if name == "first":
            return c
        if name == "second":
            return t
        if name in ("pair", "oneb"):
            return np.ones_like(c)
        if name == "lor":
            return ((c != 0) | (t != 0)).astype(c.dtype)
        if name == "land":
            return ((c != 0) & (t != 0)).astype(c.dtype)
        if name == "lxor":
            return ((c != 0) ^ (t != 0)).astype(c.dtype)
        if name == "any":
            return c
    raise KeyError(name)


def _cast_np(a, dtype):
    with np.errstate(all="ignore"):
        dtype = np.dtype(dtype)
        if dtype == np.bool_:
            return a != 0
        return a.astype(dtype, copy=False)


def vec_write_back(c: BigVec, t: BigVec, mask: BigVec | None, accum, complement, structure, replace):
    ct = c.dtype
    tv = _cast_np(t.vals, ct)
    cp, tp = c.present.astype(bool), t.present.astype(bool)
    if accum is None:
        zv, zp = tv, tp
    else:
        both = cp & tp
        zv = np.where(cp, c.vals, tv)
        if both.any():
            zv = zv.copy()
            zv[both] = _cast_np(_accum_np(accum, c.vals[both], tv[both]), ct)
        zp = cp | tp
    if mask is None:
        m = np.ones(c.size, dtype=bool)
    elif structure:
        m = mask.present.astype(bool)
    else:
        m = mask.present.astype(bool) & (mask.vals != 0)
    if complement:
        m = ~m
    wp = np.where(m, zp, (cp if not replace else False))
    wv = np.where(m, zv, c.vals)
    wv = np.where(wp, wv, np.zeros((), dtype=ct)).astype(ct)
    return BigVec(wv, wp.astype(np.uint8))


def _keys(M: BigMat):
    rows = np.repeat(np.arange(M.nrows, dtype=np.int64), np.diff(M.indptr))
    return rows * np.int64(M.ncols) + M.indices


def mat_write_back(C: BigMat, T: BigMat, M: BigMat | None, accum, complement, structure, replace):
    ct = C.dtype
    kc, kt = _keys(C), _keys(T)
    tv = _cast_np(T.values, ct)
    if accum is None:
        kz, zv = kt, tv
    else:
        kz = np.union1d(kc, kt)
        zv = np.zeros(kz.shape[0], dtype=ct)
        ic = np.searchsorted(kz, kc)
        it = np.searchsorted(kz, kt)
        zv[it] = tv
        zv[ic] = C.values
        both = np.intersect1d(kc, kt)
        if both.size:
            ib = np.searchsorted(kz, both)
            cvals = C.values[np.searchsorted(kc, both)]
            tvals = tv[np.searchsorted(kt, both)]
            zv[ib] = _cast_np(_accum_np(accum, cvals, tvals), ct)

    def allowed(keys):
        if M is None:
            m = np.ones(keys.shape[0], dtype=bool)
        else:
            km = _keys(M)
            if not structure:
                km = km[M.values != 0]
            m = np.isin(keys, km)
        return ~m if complement else m

    mz = allowed(kz)
    out_k, out_v = kz[mz], zv[mz]
    if not replace:
        mc = ~allowed(kc)
        out_k = np.concatenate([out_k, kc[mc]])
        out_v = np.concatenate([out_v, C.values[mc]])
    order = np.argsort(out_k, kind="stable")
    out_k, out_v = out_k[order], out_v[order]
    ncols = max(C.ncols, 1)
    rows, cols = out_k // ncols, out_k % ncols
    indptr = np.zeros(C.nrows + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    np.cumsum(indptr, out=indptr)
    return BigMat(indptr, cols, out_v.astype(ct), C.nrows, C.ncols)


# --------------------------------------------------------------------------- the three entry points
def mxv(w: BigVec, mask, accum, semiring, A: BigMat, u: BigVec, *, t0=False, complement=False,
        structure=False, replace=False):
    A1 = A.T() if t0 else A
    t = mxv_T(semiring, A1, u)
    return vec_write_back(w, t, mask, accum, complement, structure, replace)


def vxm(w: BigVec, mask, accum, semiring, u: BigVec, A: BigMat, *, t1=False, complement=False,
        structure=False, replace=False):
    if t1:
        t = mxv_T(semiring, A, u, flip=True)  # u'A' = (A u)' with the multiply's operands swapped
    else:
        t = vxm_push_T(semiring, u, A)
    return vec_write_back(w, t, mask, accum, complement, structure, replace)


def mxm(C: BigMat, M, accum, semiring, A: BigMat, B: BigMat, *, t0=False, t1=False, complement=False,
        structure=False, replace=False):
    A1 = A.T() if t0 else A
    B1 = B.T() if t1 else B
    T = mxm_T(semiring, A1, B1)
    return mat_write_back(C, T, M, accum, complement, structure, replace)
