"""ORACLE (test infrastructure only -- never imported by the product path).

Dict-based restatement of the GraphBLAS semantics of GrB_mxm / GrB_mxv / GrB_vxm
with mask, accumulator, replace and the T0/T1 descriptor bits, as consumed by
python-graphblas.

The arithmetic of this path is NOT in /root/reference: it lives in the third-party
C library SuiteSparse:GraphBLAS (PyPI ``suitesparse-graphblas >=7.4.0.0``,
reference pyproject.toml:64), which is not installed here.  This file restates
the published algorithm (GraphBLAS C API 2.0, cited by reference README.md:52)
and is pinned by the reference's own known-answer tests (tests/golden/*.json,
transcribed from graphblas/tests/test_matrix.py:307-392 and
graphblas/tests/test_vector.py:299-368).

What it follows, by reference file:line
  * operation definition  : docs/user_guide/operations.rst:4-153
  * mask/accum/replace    : docs/user_guide/fundamentals.rst:40-75,
                            graphblas/core/base.py:458-503 (which flags exist)
  * S / V / ~ mask flavours: graphblas/core/mask.py:133-203
  * T0/T1/R/S/C descriptor : graphblas/core/descriptor.py:51-84
  * domain = unify(A,B)    : graphblas/core/dtypes.py:552-568
  * pair forces INT64      : graphblas/core/operator/binary.py:387-388
  * integer wrap-around    : graphblas/tests/test_matrix.py:4379-4405 (test_power)

Pure-Python loops: use only for small cases.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------- dtypes
DTYPES = {
    "BOOL": np.bool_,
    "INT8": np.int8,
    "INT16": np.int16,
    "INT32": np.int32,
    "INT64": np.int64,
    "UINT8": np.uint8,
    "UINT16": np.uint16,
    "UINT32": np.uint32,
    "UINT64": np.uint64,
    "FP32": np.float32,
    "FP64": np.float64,
}
NAME_OF = {np.dtype(v): k for k, v in DTYPES.items()}


def np_dtype(d):
    if isinstance(d, str) and d.upper() in DTYPES:
        return np.dtype(DTYPES[d.upper()])
    return np.dtype(d)


def unify(d1, d2):
    """reference graphblas/core/dtypes.py:552-568 -- numpy promote_types."""
    d1, d2 = np_dtype(d1), np_dtype(d2)
    if d1 == d2:
        return d1
    return np.promote_types(d1, d2)


def cast(value, dtype):
    """C-style typecast between builtin types (GraphBLAS spec 3.4: as in C)."""
    dtype = np_dtype(dtype)
    with np.errstate(all="ignore"):
        if dtype == np.bool_:
            return np.bool_(bool(value))
        v = np.asarray(value)
        if v.dtype.kind == "f" and dtype.kind in "iu":
            # C cast of float->int truncates toward zero (in-range values only are pinned)
            return dtype.type(int(np.trunc(float(value)))) if np.isfinite(value) else dtype.type(0)
        if v.dtype.kind in "iub" and dtype.kind in "iu":
            return v.astype(dtype)[()]  # wraps modulo 2^n
        return v.astype(dtype)[()]


# --------------------------------------------------------------------------- binary ops
def _cdiv(x, y):
    """binary.cdiv: C division, truncating toward zero for integers (tests/test_op.py:959-974)."""
    if np.asarray(x).dtype.kind in "iu":
        if y == 0:
            return type(x)(0)
        q = abs(int(x)) // abs(int(y))
        if (int(x) < 0) != (int(y) < 0):
            q = -q
        return np.asarray(q).astype(np.asarray(x).dtype)[()]
    return x / y


BINARY = {
    "first": lambda x, y: x,
    "second": lambda x, y: y,
    "pair": lambda x, y: type(x)(1),
    "oneb": lambda x, y: type(x)(1),
    "plus": lambda x, y: x + y,
    "minus": lambda x, y: x - y,
    "rminus": lambda x, y: y - x,
    "times": lambda x, y: x * y,
    "div": _cdiv,
    "cdiv": _cdiv,
    "min": lambda x, y: x if x <= y else y,
    "max": lambda x, y: x if x >= y else y,
    "lor": lambda x, y: type(x)(bool(x) or bool(y)),
    "land": lambda x, y: type(x)(bool(x) and bool(y)),
    "lxor": lambda x, y: type(x)(bool(x) != bool(y)),
    "lxnor": lambda x, y: type(x)(bool(x) == bool(y)),
    "any": lambda x, y: x,  # nondeterministic by definition; oracle keeps the first seen
}
# for floats, min/max must ignore NaN the way fmin/fmax do; inputs in tests are NaN-free.


def binary(name):
    return BINARY[name]


def split_semiring(name):
    add, mul = name.split("_", 1)
    return add, mul


def semiring_domain(name, dA, dB):
    """Type the multiply runs in (reference semantics summarised in SURVEY.md section 8c).

    pair   -> INT64 regardless of inputs (operator/binary.py:387-388)
    lor/land/lxor multiplies or adds on non-bool input -> BOOL (operator/semiring.py:538-547)
    else   -> unify(A.dtype, B.dtype)  (first/second included: operator/utils.py:60-82 falls through
              to unify because binary.first/second._custom_dtype only handles UDTs, binary.py:988-989)
    """
    add, mul = split_semiring(name)
    if mul in ("pair", "oneb"):
        return np.dtype(np.int64) if add != "lor" and add != "land" else np.dtype(np.bool_)
    if add in ("lor", "land", "lxor", "lxnor") or mul in ("lor", "land", "lxor", "lxnor"):
        return np.dtype(np.bool_)
    return unify(dA, dB)


# --------------------------------------------------------------------------- containers
class SpMat:
    """Sparse matrix: dict {(i, j): numpy scalar}. Absent != 0; stored zeros are entries."""

    def __init__(self, nrows, ncols, dtype, entries=None):
        self.nrows, self.ncols, self.dtype = int(nrows), int(ncols), np_dtype(dtype)
        self.e = {}
        if entries:
            for (i, j), v in entries.items():
                self.e[(int(i), int(j))] = cast(v, self.dtype)

    @classmethod
    def from_coo(cls, rows, cols, vals, nrows=None, ncols=None, dtype=None):
        vals = np.asarray(vals)
        if dtype is None:
            dtype = vals.dtype if vals.dtype != np.dtype(int) else np.int64
        rows = list(map(int, rows))
        cols = list(map(int, cols))
        if nrows is None:
            nrows = max(rows) + 1
        if ncols is None:
            ncols = max(cols) + 1
        return cls(nrows, ncols, dtype, {(i, j): v for i, j, v in zip(rows, cols, vals.tolist())})

    def T(self):
        return SpMat(self.ncols, self.nrows, self.dtype, {(j, i): v for (i, j), v in self.e.items()})

    def dup(self):
        return SpMat(self.nrows, self.ncols, self.dtype, dict(self.e))

    def to_coo(self):
        keys = sorted(self.e)
        return (
            np.array([k[0] for k in keys], dtype=np.int64),
            np.array([k[1] for k in keys], dtype=np.int64),
            np.array([self.e[k] for k in keys], dtype=self.dtype),
        )


class SpVec:
    def __init__(self, size, dtype, entries=None):
        self.size, self.dtype = int(size), np_dtype(dtype)
        self.e = {}
        if entries:
            for i, v in entries.items():
                self.e[int(i)] = cast(v, self.dtype)

    @classmethod
    def from_coo(cls, idx, vals, size=None, dtype=None):
        vals = np.asarray(vals)
        if dtype is None:
            dtype = vals.dtype if vals.dtype != np.dtype(int) else np.int64
        idx = list(map(int, idx))
        if size is None:
            size = max(idx) + 1
        return cls(size, dtype, dict(zip(idx, vals.tolist())))

    def dup(self):
        return SpVec(self.size, self.dtype, dict(self.e))

    def as_col(self):
        return SpMat(self.size, 1, self.dtype, {(i, 0): v for i, v in self.e.items()})

    def as_row(self):
        return SpMat(1, self.size, self.dtype, {(0, i): v for i, v in self.e.items()})

    def to_coo(self):
        keys = sorted(self.e)
        return np.array(keys, dtype=np.int64), np.array([self.e[k] for k in keys], dtype=self.dtype)


# --------------------------------------------------------------------------- the operation
def _multiply(A, B, semiring):
    """T = A (+).(x) B over the structural intersection; integers wrap; no entry is dropped."""
    add_name, mul_name = split_semiring(semiring)
    D = semiring_domain(semiring, A.dtype, B.dtype)
    add, mul = binary(add_name), binary(mul_name)
    Brows = {}
    for (k, j), b in B.e.items():
        Brows.setdefault(k, []).append((j, b))
    for k in Brows:
        Brows[k].sort()
    T = {}
    with np.errstate(all="ignore"):
        for (i, k) in sorted(A.e):
            a = cast(A.e[(i, k)], D)
            for j, b in Brows.get(k, ()):
                p = cast(mul(a, cast(b, D)), D)
                key = (i, j)
                if key in T:
                    T[key] = cast(add(T[key], p), D)
                else:
                    T[key] = p
    return T, D


def _mask_fn(M, complement, structure):
    if M is None:
        return lambda key: not complement
    if structure:
        return lambda key: (key in M.e) != complement
    return lambda key: ((key in M.e) and bool(M.e[key])) != complement


def _write_back(C, T, Tdtype, M, accum, complement, structure, replace):
    """Z = accum(C,T) on the union; C<M> = Z with replace semantics. Mutates C.e.

    accum is typed by the OUTPUT dtype (reference graphblas/core/base.py:254-260).
    """
    with np.errstate(all="ignore"):
        if accum is None:
            Z = {k: cast(v, C.dtype) for k, v in T.items()}
        else:
            f = binary(accum)
            Z = dict(C.e)
            for k, t in T.items():
                if k in Z:
                    Z[k] = cast(f(Z[k], cast(t, C.dtype)), C.dtype)
                else:
                    Z[k] = cast(t, C.dtype)
    m = _mask_fn(M, complement, structure)
    new = {}
    for k in set(C.e) | set(Z):
        if m(k):
            if k in Z:
                new[k] = Z[k]
        elif not replace and k in C.e:
            new[k] = C.e[k]
    C.e = new
    return C


def mxm(C, M, accum, semiring, A, B, *, t0=False, t1=False, complement=False, structure=False, replace=False):
    """GrB_mxm(C, M, accum, semiring, A, B, desc) -- reference call site core/matrix.py:2319-2328."""
    A1 = A.T() if t0 else A
    B1 = B.T() if t1 else B
    if A1.ncols != B1.nrows or C.nrows != A1.nrows or C.ncols != B1.ncols:
        raise ValueError("GrB_DIMENSION_MISMATCH")
    T, D = _multiply(A1, B1, semiring)  # before touching C: C may alias A, B, M
    Mcopy = None if M is None else M.dup()
    return _write_back(C, T, D, Mcopy, accum, complement, structure, replace)


def _vec_from_mat(C, mat, axis):
    C.e = {(k[0] if axis == 0 else k[1]): v for k, v in mat.e.items()}
    return C


def mxv(w, m, accum, semiring, A, u, *, t0=False, complement=False, structure=False, replace=False):
    """GrB_mxv(w, mask, accum, semiring, A, u, desc) -- core/matrix.py:2252-2259; only INP0 can be transposed."""
    Cm = w.as_col()
    mxm(Cm, None if m is None else m.as_col(), accum, semiring, A, u.as_col(), t0=t0,
        complement=complement, structure=structure, replace=replace)
    return _vec_from_mat(w, Cm, 0)


def vxm(w, m, accum, semiring, u, A, *, t1=False, complement=False, structure=False, replace=False):
    """GrB_vxm(w, mask, accum, semiring, u, A, desc) -- core/vector.py:1368-1375; only INP1 can be transposed."""
    Cm = w.as_row()
    mxm(Cm, None if m is None else m.as_row(), accum, semiring, u.as_row(), A, t1=t1,
        complement=complement, structure=structure, replace=replace)
    return _vec_from_mat(w, Cm, 1)


def inner(semiring, u, v):
    """Vector.inner -- reference core/vector.py:1715-1744: GrB_vxm of u against v cast to an n x 1 matrix; the single entry of
    the 1-vector result, or None when u and v share no index.  Returns (value, dtype)."""
    if u.size != v.size:
        raise ValueError("GrB_DIMENSION_MISMATCH")
    T, D = _multiply(u.as_row(), v.as_col(), semiring)
    return (T.get((0, 0)), D)


def outer(binop, u, v):
    """Vector.outer -- reference core/vector.py:1746-1787: GrB_mxm(any_<binop>) of u as n x 1 and v as 1 x m (GrB_DESC_T1 on the
    column form).  Returns an SpMat."""
    T, D = _multiply(u.as_col(), v.as_row(), f"any_{binop}")
    return SpMat(u.size, v.size, D, T)


# --------------------------------------------------------------------------- matrix element-wise operations
COMPARE = {"eq": lambda x, y: x == y, "ne": lambda x, y: x != y, "lt": lambda x, y: x < y, "gt": lambda x, y: x > y,
           "le": lambda x, y: x <= y, "ge": lambda x, y: x >= y}
UNARY = {"identity": lambda x: x, "ainv": lambda x: -x, "abs": lambda x: abs(x), "one": lambda x: type(x)(1),
         "lnot": lambda x: type(x)(not bool(x))}


def transpose(C, M, accum, A, *, t0=False, complement=False, structure=False, replace=False):
    """GrB_transpose(C, M, accum, A, desc) -- reference core/base.py:401-411; with INP0 transposed it is a plain copy."""
    A1 = A if t0 else A.T()
    if (C.nrows, C.ncols) != (A1.nrows, A1.ncols):
        raise ValueError("GrB_DIMENSION_MISMATCH")
    return _write_back(C, dict(A1.e), A.dtype, None if M is None else M.dup(), accum, complement, structure, replace)


def ewise(C, M, accum, op, A, B, *, union, t0=False, t1=False, complement=False, structure=False, replace=False):
    """GrB_Matrix_eWiseAdd / eWiseMult_BinaryOp -- reference core/matrix.py:1972-2108: the op on the intersection; eWiseAdd
    also copies the entries present on one side only.  Comparison ops return BOOL."""
    A1, B1 = (A.T() if t0 else A), (B.T() if t1 else B)
    if (A1.nrows, A1.ncols) != (B1.nrows, B1.ncols) or (C.nrows, C.ncols) != (A1.nrows, A1.ncols):
        raise ValueError("GrB_DIMENSION_MISMATCH")
    D = unify(A.dtype, B.dtype)
    cmp = op in COMPARE
    f = COMPARE[op] if cmp else binary(op)
    Z = np.dtype(np.bool_) if cmp else D
    T = {}
    with np.errstate(all="ignore"):
        for k in set(A1.e) | set(B1.e):
            if k in A1.e and k in B1.e:
                T[k] = cast(f(cast(A1.e[k], D), cast(B1.e[k], D)), Z)
            elif union:
                T[k] = cast(cast(A1.e[k] if k in A1.e else B1.e[k], D), Z)
    return _write_back(C, T, Z, None if M is None else M.dup(), accum, complement, structure, replace)


def apply(C, M, accum, op, A, *, scalar=None, scalar_first=False, t0=False, complement=False, structure=False, replace=False):
    """GrB_Matrix_apply (unary) / apply_BinaryOp1st / 2nd -- reference core/matrix.py:2440-2533; the pattern is A's."""
    A1 = A.T() if t0 else A
    D = A.dtype if scalar is None else unify(A.dtype, np.asarray(scalar).dtype if not isinstance(scalar, (bool, int, float)) else
                                            (np.bool_ if isinstance(scalar, bool) else np.int64 if isinstance(scalar, int) else np.float64))
    T = {}
    with np.errstate(all="ignore"):
        for k, v in A1.e.items():
            x = cast(v, D)
            if scalar is None:
                T[k] = cast(UNARY[op](x), D)
            else:
                sc = cast(scalar, D)
                T[k] = cast(binary(op)(sc, x) if scalar_first else binary(op)(x, sc), D)
    return _write_back(C, T, D, None if M is None else M.dup(), accum, complement, structure, replace)


def reduce_scalar(monoid, A):
    """all entries folded in row-major order with the monoid; None for an empty matrix"""
    keys = sorted(A.e)
    if not keys:
        return None
    f = binary(monoid)
    acc = A.e[keys[0]]
    with np.errstate(all="ignore"):
        for k in keys[1:]:
            acc = cast(f(acc, A.e[k]), A.dtype)
    return acc


SELECT = {   # GrB_IndexUnaryOp predicates f(x, i, j, thunk): GraphBLAS C API 2.0 table 3.9; reference core/operator/select.py
    "tril": lambda x, i, j, y: j <= i + y, "triu": lambda x, i, j, y: j >= i + y,
    "diag": lambda x, i, j, y: j == i + y, "offdiag": lambda x, i, j, y: j != i + y,
    "colle": lambda x, i, j, y: j <= y, "colgt": lambda x, i, j, y: j > y,
    "rowle": lambda x, i, j, y: i <= y, "rowgt": lambda x, i, j, y: i > y,
    "valueeq": lambda x, i, j, y: x == y, "valuene": lambda x, i, j, y: x != y,
    "valuegt": lambda x, i, j, y: x > y, "valuege": lambda x, i, j, y: x >= y,
    "valuelt": lambda x, i, j, y: x < y, "valuele": lambda x, i, j, y: x <= y,
}
_POSITIONAL = {"tril", "triu", "diag", "offdiag", "colle", "colgt", "rowle", "rowgt"}


def select(C, M, accum, op, A, thunk, *, t0=False, complement=False, structure=False, replace=False):
    """GrB_{Matrix,Vector}_select -- reference core/matrix.py:2560-2630, core/vector.py:1560-1631: keep the entries of A for which
    op(value, row, col, thunk) holds; kept values are unchanged.  Value comparisons run in unify(A.dtype, thunk dtype); a vector is
    an n x 1 SpMat here (col = 0)."""
    A1 = A.T() if t0 else A
    if (C.nrows, C.ncols) != (A1.nrows, A1.ncols):
        raise ValueError("GrB_DIMENSION_MISMATCH")
    f = SELECT[op]
    if op in _POSITIONAL:
        y = int(thunk)
        T = {k: v for k, v in A1.e.items() if f(v, k[0], k[1], y)}
    else:
        tdt = np.bool_ if isinstance(thunk, bool) else np.int64 if isinstance(thunk, int) else np.float64 if isinstance(thunk, float) \
            else np.asarray(thunk).dtype
        D = unify(A.dtype, tdt)
        y = cast(thunk, D)
        T = {k: v for k, v in A1.e.items() if f(cast(v, D), k[0], k[1], y)}
    return _write_back(C, T, A.dtype, None if M is None else M.dup(), accum, complement, structure, replace)
