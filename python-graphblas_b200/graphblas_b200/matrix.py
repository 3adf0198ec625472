"""Matrix, TransposedMatrix, MatrixExpression -- host mirror of reference graphblas/core/matrix.py for the hot path:
construction (:190-203), from_coo/build (:627-681), from_csr/_from_csx (:992-1068), to_csr/_to_csx (:1601-1645),
to_coo (:525-594), mxv (:2233-2262), mxm (:2294-2331), TransposedMatrix (:3825-3920), isequal (:373-415)."""
import ctypes

import numpy as np

from . import operator
from ._lib import GrB_Index, lib
from .base import BaseExpression, BaseType, StructuralMask, ValueMask, call
from .dtypes import BOOL, FP64, INT64, lookup_dtype, unify
from .exceptions import InvalidValue, NoValue
from .scalar import Scalar, ScalarExpression
from .vector import (Vector, VectorExpression, _CScalar, _monoid_identity, _ptr, _Ref, _scalar_dtype, ints_to_numpy_buffer,
                     values_to_numpy_buffer)

_name_counter = [0]
_CSR, _CSC, _COO = 0, 1, 2


class Matrix(BaseType):
    ndim = 2
    _is_transposed = False

    def __init__(self, dtype=FP64, nrows=0, ncols=0, *, name=None):
        self.dtype = lookup_dtype(dtype)
        self._nrows, self._ncols = int(nrows), int(ncols)
        self.gb_obj = ctypes.c_void_p()
        if name is None:
            _name_counter[0] += 1
            name = f"M_{_name_counter[0]}"
        self.name = name
        call("GrB_Matrix_new", [_Ref(self), self.dtype, GrB_Index(self._nrows), GrB_Index(self._ncols)])

    @classmethod
    def _from_handle(cls, handle, dtype, nrows, ncols, name=None):
        self = object.__new__(cls)
        self.dtype, self._nrows, self._ncols, self.gb_obj = lookup_dtype(dtype), int(nrows), int(ncols), handle
        _name_counter[0] += 1
        self.name = name or f"M_{_name_counter[0]}"
        return self

    def __del__(self):
        gb_obj = getattr(self, "gb_obj", None)
        if gb_obj is not None and gb_obj.value and lib is not None:
            try:
                lib().GrB_Matrix_free(ctypes.byref(gb_obj))
            except Exception:
                pass

    @property
    def _carg(self):
        return self.gb_obj

    # ---- metadata
    @property
    def nrows(self):
        return self._nrows

    @property
    def ncols(self):
        return self._ncols

    @property
    def shape(self):
        return (self._nrows, self._ncols)

    @property
    def nvals(self):
        n = GrB_Index()
        call("GrB_Matrix_nvals", [ctypes.byref(n), self])
        return n.value

    @property
    def T(self):
        return TransposedMatrix(self)

    @property
    def S(self):
        return StructuralMask(self)

    @property
    def V(self):
        return ValueMask(self)

    def __repr__(self):
        return f"Matrix({self.name!r}, nvals={self.nvals}, nrows={self._nrows}, ncols={self._ncols}, dtype={self.dtype})"

    # ---- data in / out
    @classmethod
    def from_coo(cls, rows, columns, values=1.0, dtype=None, *, nrows=None, ncols=None, dup_op=None, name=None):
        rows, columns = ints_to_numpy_buffer(rows, "rows"), ints_to_numpy_buffer(columns, "columns")
        if np.ndim(values) == 0:
            values = np.full(rows.shape[0], values)
        values, dtype = values_to_numpy_buffer(values, dtype)
        if nrows is None:
            if rows.size == 0:
                raise ValueError("No row indices provided. Unable to infer nrows.")
            nrows = int(rows.max()) + 1
        if ncols is None:
            if columns.size == 0:
                raise ValueError("No column indices provided. Unable to infer ncols.")
            ncols = int(columns.max()) + 1
        C = cls(dtype, nrows, ncols, name=name)
        C.build(rows, columns, values, dup_op=dup_op)
        return C

    def build(self, rows, columns, values, *, dup_op=None, clear=False):
        rows, columns = ints_to_numpy_buffer(rows, "rows"), ints_to_numpy_buffer(columns, "columns")
        values, _ = values_to_numpy_buffer(values, self.dtype)
        n = values.shape[0]
        if rows.shape[0] != n or columns.shape[0] != n:
            raise ValueError(f"`rows` and `columns` and `values` lengths must match: {rows.size}, {columns.size}, {n}")
        if clear:
            self.clear()
        if dup_op is not None:
            dup_op = operator.get_typed_op(dup_op, self.dtype, kind="binary")
            if dup_op.opclass == "Monoid":
                dup_op = dup_op.binaryop
        try:
            call(f"GrB_Matrix_build_{self.dtype.name}", [self, _ptr(rows), _ptr(columns), _ptr(values), GrB_Index(n), dup_op])
        except InvalidValue:
            if dup_op is None:   # reference core/matrix.py:680-681 raises ValueError here
                raise ValueError("Duplicate indices found, must provide `dup_op` BinaryOp") from None
            raise

    @classmethod
    def _from_csx(cls, fmt, indptr, indices, values, dtype, num, check_num, name):
        """reference core/matrix.py:992-1068 (vanilla path: GrB_Matrix_import_<T>)"""
        indptr = ints_to_numpy_buffer(indptr, "indptr")
        indices = ints_to_numpy_buffer(indices, "indices")
        if np.ndim(values) == 0:
            values = np.full(indices.shape[0], values)
        values, dtype = values_to_numpy_buffer(values, dtype)
        if num is None:
            if indices.size > 0:
                num = int(indices.max()) + 1
            else:
                raise ValueError(f"No indices provided. Unable to infer {check_num}.")
        if fmt == _CSR:
            nrows, ncols = indptr.size - 1, num
        else:
            ncols, nrows = indptr.size - 1, num
        h = ctypes.c_void_p()
        out = cls._from_handle(h, dtype, nrows, ncols, name)
        call(f"GrB_Matrix_import_{dtype.name}",
             [ctypes.byref(h), dtype, GrB_Index(nrows), GrB_Index(ncols), _ptr(indptr), _ptr(indices), _ptr(values),
              GrB_Index(indptr.size), GrB_Index(indices.size), GrB_Index(values.shape[0]), fmt])
        return out

    @classmethod
    def from_csr(cls, indptr, col_indices, values=1.0, dtype=None, *, ncols=None, name=None):
        return cls._from_csx(_CSR, indptr, col_indices, values, dtype, ncols, "ncols", name)

    @classmethod
    def from_csc(cls, indptr, row_indices, values=1.0, dtype=None, *, nrows=None, name=None):
        return cls._from_csx(_CSC, indptr, row_indices, values, dtype, nrows, "nrows", name)

    def _to_csx(self, fmt, dtype=None):
        """reference core/matrix.py:1601-1645: exportSize, then export into caller-owned numpy buffers"""
        a, b, c = GrB_Index(), GrB_Index(), GrB_Index()
        call("GrB_Matrix_exportSize", [ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), fmt, self])
        Ap = np.empty(a.value, dtype=np.uint64)
        Ai = np.empty(b.value, dtype=np.uint64)
        Ax = np.empty(c.value, dtype=self.dtype.np_type)
        call(f"GrB_Matrix_export_{self.dtype.name}",
             [_ptr(Ap), _ptr(Ai), _ptr(Ax), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), fmt, self])
        if dtype is not None and lookup_dtype(dtype) is not self.dtype:
            Ax = Ax.astype(lookup_dtype(dtype).np_type)
        return Ap, Ai, Ax   # rows come back sorted by this backend (the reference sorts in numpy when asked)

    def to_csr(self, dtype=None, *, sort=True):
        return self._to_csx(_CSR, dtype)

    def to_csc(self, dtype=None, *, sort=True):
        return self._to_csx(_CSC, dtype)

    def to_coo(self, dtype=None, *, rows=True, columns=True, values=True, sort=True):
        n = self.nvals
        I = np.empty(n, dtype=np.uint64)
        J = np.empty(n, dtype=np.uint64)
        X = np.empty(n, dtype=self.dtype.np_type)
        nn = GrB_Index(n)
        call(f"GrB_Matrix_extractTuples_{self.dtype.name}", [_ptr(I), _ptr(J), _ptr(X), ctypes.byref(nn), self])
        if dtype is not None and lookup_dtype(dtype) is not self.dtype:
            X = X.astype(lookup_dtype(dtype).np_type)
        return (I if rows else None, J if columns else None, X if values else None)

    def dup(self, dtype=None, *, name=None):
        h = ctypes.c_void_p()
        call("GrB_Matrix_dup", [ctypes.byref(h), self])
        out = Matrix._from_handle(h, self.dtype, self._nrows, self._ncols, name)
        if dtype is not None and lookup_dtype(dtype) is not self.dtype:
            r, c, v = out.to_coo()
            return Matrix.from_coo(r, c, v.astype(lookup_dtype(dtype).np_type), nrows=self._nrows, ncols=self._ncols, name=name)
        return out

    def clear(self):
        call("GrB_Matrix_clear", [self])

    def wait(self, how="materialize"):
        call("GrB_Matrix_wait", [self, 1 if how == "materialize" else 0])
        return self

    def __getitem__(self, ij):
        if isinstance(ij, tuple) and len(ij) == 2 and all(isinstance(k, (int, np.integer)) for k in ij):
            def thunk():
                x = self.dtype.ctype()
                rv = call(f"GrB_Matrix_extractElement_{self.dtype.name}",
                          [ctypes.byref(x), self, GrB_Index(int(ij[0])), GrB_Index(int(ij[1]))])
                return None if rv is NoValue else x.value
            return ScalarExpression(self.dtype, thunk)
        raise NotImplementedError("only scalar extraction A[i, j] is on this backend's path")

    # ---- expressions (each is ONE C call; mask / accum / replace are applied by the library's write-back)
    def _dup_expr(self):
        return MatrixExpression("apply", "GrB_Matrix_apply", [self], op=operator.unary.identity[self.dtype],
                                nrows=self._nrows, ncols=self._ncols, at=self._is_transposed)

    def _scalar_assign_expr(self, value):
        """C(mask, accum, replace) << scalar -- GrB_Matrix_assign_<T> over GrB_ALL x GrB_ALL (reference core/matrix.py:3120-3180,
        used by the mask-combination recipes core/mask.py:232-287).  Without a mask, or under a complemented one, the result is
        a DENSE matrix: outside this backend's path.  Under a mask that names positions the call is a recipe over kernels that
        exist: T = the scalar on the positions the mask allows (second(M, scalar) on the mask's pattern, for a value mask first
        select(valuene 0)), then the ordinary masked write-back C<M, replace> accum= T."""
        if isinstance(value, Scalar):
            value = value.value
        if value is None:
            raise NotImplementedError("assigning an empty scalar (deleting entries) is outside this backend's path")
        vt = _scalar_dtype(value)
        nrows, ncols = self._nrows, self._ncols

        def run(out, mask, accum, desc):
            if mask is None or mask.complement:
                raise NotImplementedError("a scalar assigned to every position (no mask, or a complemented one) makes the matrix "
                                          "dense: outside this backend's path")
            src = mask._true_set()
            T = src.apply(operator.binary.second, right=value).new()
            call("GrB_Matrix_apply", [out, mask, accum, operator.unary.identity[T.dtype], T, desc])

        return MatrixExpression("assign", None, [], dtype=vt, nrows=nrows, ncols=ncols, custom=run)

    def ewise_add(self, other, op=None):
        """reference core/matrix.py:1972-2056 (union of the patterns)"""
        return _ewise(self, other, op, "ewise_add", "GrB_Matrix_eWiseAdd_BinaryOp", operator.monoid.plus)

    def ewise_mult(self, other, op=None):
        """reference core/matrix.py:2058-2108 (intersection of the patterns)"""
        return _ewise(self, other, op, "ewise_mult", "GrB_Matrix_eWiseMult_BinaryOp", operator.binary.times)

    def apply(self, op, right=None, *, left=None):
        """reference core/matrix.py:2440-2533: unary op, or a binary op with the scalar bound first (left=) / second (right=)"""
        return _apply(self, op, right, left)

    def power(self, n, op=None):
        """reference core/matrix.py:2840-2905 (+ `_power`, :101-164): A^n over a semiring by repeated squaring, every product one
        GrB_mxm on the device; n = 0 gives the diagonal matrix of the multiply's monoid identity."""
        from .exceptions import DimensionMismatch

        if self._nrows != self._ncols:
            raise DimensionMismatch(f"power only works for square Matrix; shape is {self.shape}")
        if isinstance(n, (bool, np.bool_)) or not isinstance(n, (int, np.integer)):
            raise TypeError(f"n must be a nonnegative integer; got bad type: {type(n)}")
        N = int(n)
        if N < 0:
            raise ValueError(f"n must be a nonnegative integer; got: {N}")
        op = operator.semiring.plus_times if op is None else op
        op = operator.get_typed_op(op, self.dtype, kind="semiring")
        if op.opclass != "Semiring":
            raise TypeError(f"power expects a Semiring, got {op.opclass}")
        ident = _MULT_IDENTITY.get(op.parent.binaryop.name)
        if N == 0 and ident is None:
            raise ValueError(f"Binary operator of {op} semiring does not have a monoid with an identity. When n=0, the result is a "
                             "diagonal matrix with values equal to the identity of the binaryop, so the binaryop must be associated "
                             "with a monoid.")
        me = self

        def run(out, mask, accum, desc):
            if N == 0:
                idx = np.arange(me._nrows, dtype=np.uint64)
                P = Matrix.from_coo(idx, idx, np.full(me._nrows, ident(me.dtype.np_type), dtype=me.dtype.np_type),
                                    nrows=me._nrows, ncols=me._ncols, dtype=me.dtype)
            else:
                P, square, k = None, me, N
                while True:
                    if k & 1:
                        P = square if P is None else P.mxm(square, op).new()
                    k >>= 1
                    if k == 0:
                        break
                    square = square.mxm(square, op).new()
            call("GrB_Matrix_apply", [out, mask, accum, operator.unary.identity[P.dtype], P, desc])

        return MatrixExpression("power", None, [], dtype=op.return_type, nrows=self._nrows, ncols=self._ncols, custom=run)

    def reduce_rowwise(self, op=None):
        """reference core/matrix.py:2600-2650 (GrB_Matrix_reduce_Monoid): w(i) = (+)_j A(i, j) over the entries of row i; rows
        without entries give no entry.  Run as ONE pull SpMV with the semiring <monoid>_first against an all-present vector:
        first(a, x) = a, so the multiply passes A's values through and the vector's values are never read."""
        if getattr(op, "opclass", None) == "Aggregator":   # a recipe over the multiply (graphblas_b200/agg.py)
            return op._rowwise_expr(self)
        return _reduce_to_vector(self, op, "reduce_rowwise")

    def reduce_columnwise(self, op=None):
        """reference core/matrix.py:2652-2701: the row-wise reduction of the transpose (GrB_DESC_T0)"""
        if getattr(op, "opclass", None) == "Aggregator":
            return op._rowwise_expr(self.T)
        return _reduce_to_vector(self.T, op, "reduce_columnwise")

    def reduce_scalar(self, op=None, *, allow_empty=True):
        """reference core/matrix.py:2703-2760: GrB_Matrix_reduce_Monoid_Scalar into a GrB_Scalar, or (allow_empty=False)
        GrB_Matrix_reduce_<T> into a C scalar"""
        if getattr(op, "opclass", None) == "Aggregator":
            return op._scalar_expr(self, True)
        op = operator.monoid.plus if op is None else op
        op = operator.get_typed_op(op, self.dtype, kind="monoid")
        if op.opclass != "Monoid":
            raise TypeError("reduce_scalar expects a Monoid")
        me = self
        if allow_empty:
            def run(out, accum):
                call("GrB_Matrix_reduce_Monoid_Scalar", [out, accum, op, me, None])

            return ScalarExpression(op.return_type, run=run)

        def thunk():
            x = op.return_type.ctype(op.return_type.np_type.type(_monoid_identity(op)).item())
            call(f"GrB_Matrix_reduce_{op.return_type.name}", [ctypes.byref(x), None, op, me, None])
            return x.value

        return ScalarExpression(op.return_type, thunk)

    def select(self, op, thunk=None):
        """reference core/matrix.py:2560-2630"""
        return _select(self, op, thunk)

    def mxv(self, other, op=None):
        """reference core/matrix.py:2233-2262"""
        return _mxv(self, other, op)

    def mxm(self, other, op=None):
        """reference core/matrix.py:2294-2331"""
        return _mxm(self, other, op)

    # ---- comparison: exactly the reference's recipe (core/matrix.py:373-415) -- eWiseMult(EQ) then reduce(LAND), on the device
    def isequal(self, other, *, check_dtype=False):
        if type(other) is not Matrix:
            raise TypeError(f"isequal expects a Matrix, got {type(other).__name__}")
        if check_dtype and self.dtype != other.dtype:
            return False
        if self.shape != other.shape or self.nvals != other.nvals:
            return False
        common = unify(self.dtype, other.dtype)
        matches = self.ewise_mult(other, operator.binary.eq[common]).new(BOOL)
        if matches.nvals != self.nvals:
            return False
        return bool(matches.reduce_scalar(operator.monoid.land, allow_empty=False).value)

    def isclose(self, other, *, rel_tol=1e-7, abs_tol=0.0, check_dtype=False):
        if check_dtype and self.dtype != other.dtype:
            return False
        if self.shape != other.shape or self.nvals != other.nvals:
            return False
        a, b = self.to_coo(), other.to_coo()
        return bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and
                    np.all(np.isclose(a[2].astype(np.float64), b[2].astype(np.float64), rtol=rel_tol, atol=abs_tol)))


class TransposedMatrix:
    """A.T: the transpose lives only in the descriptor (reference core/matrix.py:3825-3920, _carg :3886-3888)."""

    ndim = 2
    _is_transposed = True

    def __init__(self, matrix):
        self._matrix = matrix

    @property
    def dtype(self):
        return self._matrix.dtype

    @property
    def _carg(self):
        return self._matrix.gb_obj

    @property
    def name(self):
        return f"{self._matrix.name}.T"

    @property
    def gb_name(self):
        return self._matrix.name

    @property
    def _nrows(self):
        return self._matrix._ncols

    @property
    def _ncols(self):
        return self._matrix._nrows

    nrows, ncols = _nrows, _ncols

    @property
    def shape(self):
        return (self._nrows, self._ncols)

    @property
    def T(self):
        return self._matrix

    def mxv(self, other, op=None):
        return _mxv(self, other, op)

    def mxm(self, other, op=None):
        return _mxm(self, other, op)

    def __matmul__(self, other):
        from .infix import matmul

        return matmul(self, other)

    def _transpose_expr(self):
        """C(mask, accum) << A.T  ->  GrB_transpose(C, mask, accum, A, desc), reference core/base.py:401-411"""
        me = self._matrix
        return MatrixExpression("transpose", "GrB_transpose", [me], dtype=me.dtype, nrows=me._ncols, ncols=me._nrows)

    def ewise_add(self, other, op=None):
        return _ewise(self, other, op, "ewise_add", "GrB_Matrix_eWiseAdd_BinaryOp", operator.monoid.plus)

    def ewise_mult(self, other, op=None):
        return _ewise(self, other, op, "ewise_mult", "GrB_Matrix_eWiseMult_BinaryOp", operator.binary.times)

    def apply(self, op, right=None, *, left=None):
        return _apply(self, op, right, left)

    def select(self, op, thunk=None):
        return _select(self, op, thunk)

    def reduce_rowwise(self, op=None):
        if getattr(op, "opclass", None) == "Aggregator":
            return op._rowwise_expr(self)
        return _reduce_to_vector(self, op, "reduce_rowwise")

    def reduce_columnwise(self, op=None):
        if getattr(op, "opclass", None) == "Aggregator":
            return op._rowwise_expr(self._matrix)
        return _reduce_to_vector(self._matrix, op, "reduce_columnwise")

    def new(self, dtype=None, *, name=None):
        out = Matrix(dtype or self.dtype, self._nrows, self._ncols, name=name)
        out << self
        return out


def _np_limit(kind):
    def f(np_type):
        dt = np.dtype(np_type)
        if dt == np.bool_:
            return kind == "max"          # identity of min on BOOL (= land) is True, of max (= lor) is False
        if dt.kind == "f":
            return np.inf if kind == "max" else -np.inf
        info = np.iinfo(dt)
        return info.max if kind == "max" else info.min
    return f


# identity of the monoid the semiring's multiply belongs to (what A.power(0) puts on the diagonal)
_MULT_IDENTITY = {"times": lambda t: 1, "plus": lambda t: 0, "min": _np_limit("max"), "max": _np_limit("min"),
                  "land": lambda t: True, "lor": lambda t: False, "lxor": lambda t: False, "lxnor": lambda t: True,
                  "any": lambda t: 0}


def _reduce_to_vector(self, op, method_name):
    op = operator.monoid.plus if op is None else op
    op = operator.get_typed_op(op, self.dtype, kind="monoid")
    if op.opclass != "Monoid":
        raise TypeError(f"{method_name} expects a Monoid, got {op.opclass}")
    sr = getattr(operator.semiring, f"{op.parent.name}_first", None)
    if sr is None or op.type not in sr:
        raise NotImplementedError(f"no builtin semiring {op.parent.name}_first[{op.type.name}] behind {method_name}")
    ones = Vector(self.dtype, self._ncols)
    ones[:] = 1
    return VectorExpression(method_name, "GrB_mxv", [self, ones], op=sr[op.type], size=self._nrows, at=self._is_transposed)


def _ewise_broadcast(self, other, op, method_name, default_op):
    """Matrix (op) Vector: the vector is broadcast along the rows (reference core/matrix.py:62-75, 1923-1940, 2014-2031).
    ewise_mult: C = A (any).(op) diag(v) -- one GrB_mxm against the diagonal matrix of v, so C(i, j) = op(A(i, j), v(j)) wherever both
    exist.  ewise_add: the union with the vector repeated in every row -- a dense nrows x ncols operand (0 (any).(second) v as an outer
    product, built under the update's mask like the reference does), so only sensible for small or masked results."""
    from .exceptions import DimensionMismatch

    op = default_op if op is None else op
    op = operator.get_typed_op(op, self.dtype, other.dtype, kind="binary")
    if op.opclass == "Monoid":
        op = op.binaryop
    if op.opclass != "BinaryOp":
        raise TypeError(f"{method_name} expects a BinaryOp or Monoid, got {op.opclass}")
    if self._ncols != other._size:
        raise DimensionMismatch(f"Dimensions not compatible for broadcasting Vector from the right to rows of Matrix in {method_name}.  "
                                f"Matrix.ncols (={self._ncols}) must equal Vector.size (={other._size}).")
    me = self

    def run(out, mask, accum, desc):
        if method_name == "ewise_mult":
            sr = getattr(operator.semiring, f"any_{op.parent.name}", None)
            if sr is None or op.type not in sr:
                raise NotImplementedError(f"no builtin semiring any_{op.parent.name}[{op.type.name}] behind the broadcast {method_name}")
            call("GrB_mxm", [out, mask, accum, sr[op.type], me._matrix if me._is_transposed else me, other.diag(),
                             _with_transpose(desc, me._is_transposed)])
        else:
            full = Vector(other.dtype, me._nrows)
            full[:] = 0
            temp = full.outer(other, operator.binary.second).new(mask=mask) if (mask is not None and not mask.complement) else \
                full.outer(other, operator.binary.second).new()
            call("GrB_Matrix_eWiseAdd_BinaryOp", [out, mask, accum, op, me._matrix if me._is_transposed else me, temp,
                                                  _with_transpose(desc, me._is_transposed)])

    return MatrixExpression(method_name, None, [], dtype=op.return_type, nrows=self._nrows, ncols=self._ncols, custom=run)


def _with_transpose(desc, at):
    """the update's descriptor plus GrB_DESC_T0 when the matrix operand of a broadcast recipe is A.T"""
    if not at:
        return desc
    from .base import descriptor_lookup

    name = desc.name[len("GrB_DESC_"):] if desc is not None else ""
    return descriptor_lookup(output_replace="R" in name, mask_structure="S" in name, mask_complement="C" in name, transpose_first=True)


def _ewise(self, other, op, method_name, cfunc, default_op):
    if isinstance(other, Vector):
        return _ewise_broadcast(self, other, op, method_name, default_op)
    if not isinstance(other, (Matrix, TransposedMatrix)):
        raise TypeError(f"{method_name} expects a Matrix, got {type(other).__name__}")
    op = default_op if op is None else op
    op = operator.get_typed_op(op, self.dtype, other.dtype, kind="binary")
    if op.opclass == "Monoid":
        op = op.binaryop
    if op.opclass != "BinaryOp":
        raise TypeError(f"{method_name} expects a BinaryOp or Monoid, got {op.opclass}")
    expr = MatrixExpression(method_name, cfunc, [self, other], op=op, nrows=self._nrows, ncols=self._ncols,
                            at=self._is_transposed, bt=other._is_transposed)
    if self.shape != other.shape:
        expr.new(name="")  # incompatible shape; raise now
    return expr


def _apply(self, op, right, left):
    if right is None and left is None:
        op = operator.get_typed_op(op, self.dtype, kind="unary")
        return MatrixExpression("apply", "GrB_Matrix_apply", [self], op=op, nrows=self._nrows, ncols=self._ncols,
                                at=self._is_transposed)
    scalar = right if right is not None else left
    if isinstance(scalar, Scalar) and not scalar._is_cscalar:
        sdt, carg, sfx = scalar.dtype, scalar, "Scalar"
    else:
        if isinstance(scalar, Scalar):
            scalar = scalar.value
        sdt = _scalar_dtype(scalar)
        carg, sfx = _CScalar(scalar, sdt), sdt.name
    op = operator.get_typed_op(op, self.dtype, sdt, kind="binary")
    if op.opclass == "Monoid":
        op = op.binaryop
    # reference core/matrix.py:2472 / 2518: f"GrB_Matrix_apply_BinaryOp1st_{T}" (C, Mask, accum, op, x, A, desc), 2nd: (..., A, y, desc)
    base = self._matrix if self._is_transposed else self
    # at = bt like the reference (core/matrix.py:2531-2532): the matrix is input 1 of bind-1st (GrB_INP1) and input 0 of bind-2nd
    if left is not None:
        return MatrixExpression("apply", f"GrB_Matrix_apply_BinaryOp1st_{sfx}", [carg, base], op=op, nrows=self._nrows,
                                ncols=self._ncols, at=self._is_transposed, bt=self._is_transposed)
    return MatrixExpression("apply", f"GrB_Matrix_apply_BinaryOp2nd_{sfx}", [base, carg], op=op, nrows=self._nrows, ncols=self._ncols,
                            at=self._is_transposed, bt=self._is_transposed)


def _select(self, op, thunk):
    """reference core/matrix.py:2560-2630: GrB_Matrix_select_<T>(C, Mask, accum, op, A, thunk, desc)"""
    if thunk is None:
        thunk = 0
    if isinstance(thunk, Scalar) and not thunk._is_cscalar:
        tdt, carg, sfx = thunk.dtype, thunk, "Scalar"
    else:
        if isinstance(thunk, Scalar):
            thunk = thunk.value
        tdt = _scalar_dtype(thunk)
        carg, sfx = _CScalar(thunk, tdt), tdt.name
    op = operator.get_typed_op(op, self.dtype, tdt, kind="select")
    if op.opclass != "SelectOp":
        raise TypeError(f"select expects a SelectOp, got {op.opclass}")
    base = self._matrix if self._is_transposed else self
    return MatrixExpression("select", f"GrB_Matrix_select_{sfx}", [base, carg], op=op, dtype=self.dtype, nrows=self._nrows,
                            ncols=self._ncols, at=self._is_transposed)


def _mxv(self, other, op):
    if not isinstance(other, Vector):
        raise TypeError(f"mxv expects a Vector, got {type(other).__name__}")
    op = operator.semiring.plus_times if op is None else op
    op = operator.get_typed_op(op, self.dtype, other.dtype, kind="semiring")
    if op.opclass != "Semiring":
        raise TypeError(f"mxv expects a Semiring, got {op.opclass}")
    expr = VectorExpression("mxv", "GrB_mxv", [self, other], op=op, size=self._nrows, at=self._is_transposed)
    if self._ncols != other._size:
        expr.new(name="")  # incompatible shape; raise now
    return expr


def _mxm(self, other, op):
    if not isinstance(other, (Matrix, TransposedMatrix)):
        raise TypeError(f"mxm expects a Matrix, got {type(other).__name__}")
    op = operator.semiring.plus_times if op is None else op
    op = operator.get_typed_op(op, self.dtype, other.dtype, kind="semiring")
    if op.opclass != "Semiring":
        raise TypeError(f"mxm expects a Semiring, got {op.opclass}")
    expr = MatrixExpression("mxm", "GrB_mxm", [self, other], op=op, nrows=self._nrows, ncols=other._ncols,
                            at=self._is_transposed, bt=other._is_transposed)
    if self._ncols != other._nrows:
        expr.new(name="")  # incompatible shape; raise now
    return expr


class MatrixExpression(BaseExpression):
    output_type = Matrix

    def __init__(self, *args, nrows, ncols, **kw):
        super().__init__(*args, **kw)
        self._nrows, self._ncols = nrows, ncols

    def construct_output(self, dtype=None, *, name=None):
        return Matrix(dtype or self.dtype, self._nrows, self._ncols, name=name or None)
