"""Vector, VectorExpression -- host mirror of reference graphblas/core/vector.py for the hot path:
construction (:159-170), build / to_coo (:465-568), vxm (:1341-1378), ewise / apply / reduce (:1050-1681),
isequal (:340-379).  Every operation is a lazy expression executed by ONE C call through base.call()."""
import ctypes

import numpy as np

from . import operator
from ._lib import GrB_Index, lib
from .base import (BaseExpression, BaseType, ComplementedStructuralMask, StructuralMask, ValueMask, call)
from .dtypes import BOOL, FP64, INT64, lookup_dtype, unify
from .exceptions import DimensionMismatch, InvalidValue, NoValue
from .scalar import Scalar, ScalarExpression

_name_counter = [0]


def _ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


def ints_to_numpy_buffer(array, name="indices"):
    """reference core/utils.py:58-69: index arrays cross the boundary as uint64."""
    a = np.asarray(array)
    if a.dtype.kind not in "iu" and a.size:
        raise ValueError(f"{name} must be integers, not {a.dtype}")
    return np.ascontiguousarray(a, dtype=np.uint64)


def values_to_numpy_buffer(array, dtype=None):
    """reference core/utils.py:78-114"""
    if dtype is not None:
        dtype = lookup_dtype(dtype)
        a = np.ascontiguousarray(np.asarray(array), dtype=dtype.np_type)
    else:
        a = np.ascontiguousarray(np.asarray(array))
        if a.dtype == np.dtype(object):
            raise ValueError("values must be numeric")
        dtype = lookup_dtype(a.dtype)
    return a, dtype


class Vector(BaseType):
    ndim = 1

    def __init__(self, dtype=FP64, size=0, *, name=None):
        self.dtype = lookup_dtype(dtype)
        self._size = int(size)
        self.gb_obj = ctypes.c_void_p()
        if name is None:
            _name_counter[0] += 1
            name = f"v_{_name_counter[0]}"
        self.name = name
        call("GrB_Vector_new", [_Ref(self), self.dtype, GrB_Index(self._size)])

    @classmethod
    def _from_handle(cls, handle, dtype, size, name=None):
        self = object.__new__(cls)
        self.dtype, self._size, self.gb_obj = lookup_dtype(dtype), int(size), handle
        _name_counter[0] += 1
        self.name = name or f"v_{_name_counter[0]}"
        return self

    def __del__(self):
        gb_obj = getattr(self, "gb_obj", None)
        if gb_obj is not None and gb_obj.value and lib is not None:
            try:
                lib().GrB_Vector_free(ctypes.byref(gb_obj))
            except Exception:
                pass

    @property
    def _carg(self):
        return self.gb_obj

    # ---- metadata
    @property
    def size(self):
        return self._size

    @property
    def shape(self):
        return (self._size,)

    @property
    def nvals(self):
        n = GrB_Index()
        call("GrB_Vector_nvals", [ctypes.byref(n), self])
        return n.value

    @property
    def S(self):
        return StructuralMask(self)

    @property
    def V(self):
        return ValueMask(self)

    def __repr__(self):
        return f"Vector({self.name!r}, nvals={self.nvals}, size={self._size}, dtype={self.dtype})"

    # ---- data in / out
    @classmethod
    def from_coo(cls, indices, values=1.0, dtype=None, *, size=None, dup_op=None, name=None):
        indices = ints_to_numpy_buffer(indices)
        if np.ndim(values) == 0:
            values = np.full(indices.shape[0], values)
        values, dtype = values_to_numpy_buffer(values, dtype)
        if size is None:
            if indices.size == 0:
                raise ValueError("No indices provided. Unable to infer size.")
            size = int(indices.max()) + 1
        w = cls(dtype, size, name=name)
        w.build(indices, values, dup_op=dup_op)
        return w

    def build(self, indices, values, *, dup_op=None, clear=False):
        indices = ints_to_numpy_buffer(indices)
        values, vdtype = values_to_numpy_buffer(values, self.dtype)
        if indices.shape[0] != values.shape[0]:
            raise ValueError(f"`indices` and `values` lengths must match: {indices.size}, {values.size}")
        if clear:
            self.clear()
        if dup_op is not None:
            dup_op = operator.get_typed_op(dup_op, self.dtype, kind="binary")
            if dup_op.opclass == "Monoid":
                dup_op = dup_op.binaryop
        try:
            call(f"GrB_Vector_build_{self.dtype.name}", [self, _ptr(indices), _ptr(values), GrB_Index(indices.shape[0]), dup_op])
        except InvalidValue:
            if dup_op is None:   # reference core/vector.py build: same message as Matrix.build (core/matrix.py:680-681)
                raise ValueError("Duplicate indices found, must provide `dup_op` BinaryOp") from None
            raise

    def to_coo(self, dtype=None, *, indices=True, values=True, sort=True):
        n = self.nvals
        idx = np.empty(n, dtype=np.uint64)
        out_dtype = self.dtype if dtype is None else lookup_dtype(dtype)
        vals = np.empty(n, dtype=self.dtype.np_type)
        nn = GrB_Index(n)
        call(f"GrB_Vector_extractTuples_{self.dtype.name}", [_ptr(idx), _ptr(vals), ctypes.byref(nn), self])
        if out_dtype is not self.dtype:
            vals = vals.astype(out_dtype.np_type)
        return (idx if indices else None, vals if values else None)

    @classmethod
    def from_dense(cls, values, missing_value=None, *, dtype=None, name=None):
        values, dtype = values_to_numpy_buffer(values, dtype)
        if missing_value is None:
            idx = np.arange(values.shape[0], dtype=np.uint64)
            return cls.from_coo(idx, values, dtype, size=values.shape[0], name=name)
        keep = values != missing_value
        return cls.from_coo(np.flatnonzero(keep), values[keep], dtype, size=values.shape[0], name=name)

    def to_dense(self, fill_value=None, dtype=None):
        idx, vals = self.to_coo(dtype)
        if fill_value is None and idx.size < self._size:
            raise TypeError("fill_value must be given if there are missing values")
        out = np.full(self._size, 0 if fill_value is None else fill_value, dtype=vals.dtype)
        out[idx.astype(np.int64)] = vals
        return out

    def dup(self, dtype=None, *, name=None):
        if dtype is not None and lookup_dtype(dtype) is not self.dtype:
            w = Vector(dtype, self._size, name=name)
            w << self
            return w
        h = ctypes.c_void_p()
        call("GrB_Vector_dup", [ctypes.byref(h), self])
        return Vector._from_handle(h, self.dtype, self._size, name)

    def clear(self):
        call("GrB_Vector_clear", [self])

    def wait(self, how="materialize"):
        call("GrB_Vector_wait", [self, 1 if how == "materialize" else 0])
        return self

    def __getitem__(self, i):
        if isinstance(i, (int, np.integer)):
            def thunk():
                x = self.dtype.ctype()
                rv = call(f"GrB_Vector_extractElement_{self.dtype.name}", [ctypes.byref(x), self, GrB_Index(int(i))])
                return None if rv is NoValue else x.value
            return ScalarExpression(self.dtype, thunk)
        raise NotImplementedError("only scalar extraction v[i] is on this backend's path")

    def __setitem__(self, i, value):
        if isinstance(i, (int, np.integer)):
            x = self.dtype.ctype(self.dtype.np_type.type(value).item())
            call(f"GrB_Vector_setElement_{self.dtype.name}", [self, x, GrB_Index(int(i))])
            return
        if i in (Ellipsis, slice(None)):
            self()[...] = value
            return
        raise NotImplementedError("only v[i] = x and v[:] = scalar are supported")

    def __delitem__(self, i):
        call("GrB_Vector_removeElement", [self, GrB_Index(int(i))])

    # ---- expressions
    def _dup_expr(self):
        return VectorExpression("apply", "GrB_Vector_apply", [self], op=operator.unary.identity[self.dtype], size=self._size)

    def _scalar_assign_expr(self, value):
        """w(mask, accum)[:] = scalar -> GrB_Vector_assign_<T>(w, mask, accum, x, GrB_ALL, n, desc) (reference core/vector.py:2020-2035)"""
        if isinstance(value, Scalar):
            if not value._is_cscalar:
                return VectorExpression("assign", "GrB_Vector_assign_Scalar", [value, _All(), GrB_Index(self._size)],
                                        dtype=value.dtype, size=self._size)
            value = value.value
        vt = INT64 if isinstance(value, (int, np.integer)) and not isinstance(value, (bool, np.bool_)) else \
            BOOL if isinstance(value, (bool, np.bool_)) else FP64
        return VectorExpression("assign", f"GrB_Vector_assign_{vt.name}", [_CScalar(value, vt), _All(), GrB_Index(self._size)],
                                dtype=vt, size=self._size)

    def vxm(self, other, op=None):
        """reference core/vector.py:1341-1378: w' = v' (+).(x) A ; `other` may be A.T (-> GrB_DESC_T1)."""
        from .matrix import Matrix, TransposedMatrix

        if not isinstance(other, (Matrix, TransposedMatrix)):
            raise TypeError(f"vxm expects a Matrix, got {type(other).__name__}")
        op = operator.semiring.plus_times if op is None else op
        op = operator.get_typed_op(op, self.dtype, other.dtype, kind="semiring")
        if op.opclass != "Semiring":
            raise TypeError(f"vxm expects a Semiring, got {op.opclass}")
        expr = VectorExpression("vxm", "GrB_vxm", [self, other], op=op, size=other._ncols, bt=other._is_transposed)
        if self._size != other._nrows:
            expr.new(name="")  # incompatible shape; raise now
        return expr

    def _as_matrix(self, *, name=None):
        """This vector as an n x 1 Matrix (reference core/vector.py:193-209).  The reference's vanilla backend fills a fresh
        Matrix with a column assign; here the column is built on the device by one scan + one compaction kernel."""
        from .matrix import Matrix

        h = ctypes.c_void_p()
        out = Matrix._from_handle(h, self.dtype, self._size, 1, name or f"(GrB_Matrix){self.name}")
        call("GrB_cuda_Matrix_from_Vector", [ctypes.byref(h), self])
        return out

    def diag(self, k=0, *, name=None):
        """reference core/vector.py:605-628: the square Matrix of order size + |k| with this vector on its k-th diagonal
        (one C call: GrB_Matrix_diag)"""
        from .matrix import Matrix

        k = int(k)
        n = self._size + abs(k)
        h = ctypes.c_void_p()
        out = Matrix._from_handle(h, self.dtype, n, n, name)
        call("GrB_Matrix_diag", [ctypes.byref(h), self, ctypes.c_int64(k)])
        return out

    def inner(self, other, op=None):
        """reference core/vector.py:1715-1744: s = v' (+).(x) w, run as GrB_vxm against w cast to an n x 1 matrix; the result is
        a Scalar (empty when no index is shared)."""
        if type(other) is not Vector:
            raise TypeError(f"inner expects a Vector, got {type(other).__name__}")
        op = operator.semiring.plus_times if op is None else op
        op = operator.get_typed_op(op, self.dtype, other.dtype, kind="semiring")
        if op.opclass != "Semiring":
            raise TypeError(f"inner expects a Semiring, got {op.opclass}")
        if self._size != other._size:
            raise DimensionMismatch(f"inner: sizes differ ({self._size} vs {other._size})")
        me = self

        def thunk():
            w = Vector(op.return_type, 1)
            w << VectorExpression("inner", "GrB_vxm", [me, other._as_matrix()], op=op, size=1)
            return w[0].new().value

        return ScalarExpression(op.return_type, thunk)

    def outer(self, other, op=None):
        """reference core/vector.py:1746-1787: C = v (any).(op) w', run as GrB_mxm of the two column matrices with GrB_DESC_T1."""
        from .matrix import MatrixExpression

        if type(other) is not Vector:
            raise TypeError(f"outer expects a Vector, got {type(other).__name__}")
        op = operator.binary.times if op is None else op
        op = operator.get_typed_op(op, self.dtype, other.dtype, kind="binary")
        if op.opclass == "Monoid":
            op = op.binaryop
        if op.opclass != "BinaryOp":
            raise TypeError(f"outer expects a BinaryOp or Monoid, got {op.opclass}")
        sr = getattr(operator.semiring, f"any_{op.parent.name}", None)
        if sr is None or op.type not in sr:
            raise NotImplementedError(f"no builtin semiring any_{op.parent.name}[{op.type.name}] behind Vector.outer")
        return MatrixExpression("outer", "GrB_mxm", [self._as_matrix(), other._as_matrix()], op=sr[op.type],
                                nrows=self._size, ncols=other._size, bt=True)

    def ewise_add(self, other, op=None):
        op = operator.monoid.plus if op is None else op
        op = operator.get_typed_op(op, self.dtype, other.dtype, kind="binary")
        if op.opclass == "Monoid":
            op = op.binaryop
        return VectorExpression("ewise_add", "GrB_Vector_eWiseAdd_BinaryOp", [self, other], op=op, size=self._size)

    def ewise_mult(self, other, op=None):
        op = operator.binary.times if op is None else op
        op = operator.get_typed_op(op, self.dtype, other.dtype, kind="binary")
        if op.opclass == "Monoid":
            op = op.binaryop
        return VectorExpression("ewise_mult", "GrB_Vector_eWiseMult_BinaryOp", [self, other], op=op, size=self._size)

    def apply(self, op, right=None, *, left=None):
        if right is None and left is None:
            op = operator.get_typed_op(op, self.dtype, kind="unary")
            return VectorExpression("apply", "GrB_Vector_apply", [self], op=op, size=self._size)
        scalar = right if right is not None else left
        if isinstance(scalar, Scalar) and not scalar._is_cscalar:   # a GrB_Scalar: ..._BinaryOp1st_Scalar / 2nd_Scalar
            sdt, carg, sfx = scalar.dtype, scalar, "Scalar"
        else:
            if isinstance(scalar, Scalar):
                scalar = scalar.value
            sdt = _scalar_dtype(scalar)
            carg, sfx = _CScalar(scalar, sdt), sdt.name
        op = operator.get_typed_op(op, self.dtype, sdt, kind="binary")
        if op.opclass == "Monoid":
            op = op.binaryop
        # reference core/vector.py:1477 / 1523: f"GrB_Vector_apply_BinaryOp1st_{T}" (w, mask, accum, op, x, u, desc), 2nd: (..., u, y, desc)
        if left is not None:
            return VectorExpression("apply", f"GrB_Vector_apply_BinaryOp1st_{sfx}", [carg, self], op=op, size=self._size)
        return VectorExpression("apply", f"GrB_Vector_apply_BinaryOp2nd_{sfx}", [self, carg], op=op, size=self._size)

    def select(self, op, thunk=None):
        """reference core/vector.py:1560-1631: keep the entries for which op(value, index, 0, thunk) holds.
        One C call: GrB_Vector_select_<T>(w, mask, accum, op, u, thunk, desc)."""
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
        return VectorExpression("select", f"GrB_Vector_select_{sfx}", [self, carg], op=op, dtype=self.dtype, size=self._size)

    def reduce(self, op=None, *, allow_empty=True):
        """reference core/vector.py:1633-1690: GrB_Vector_reduce_Monoid_Scalar into a GrB_Scalar (an empty vector gives an empty
        scalar); allow_empty=False reduces into a C scalar with GrB_Vector_reduce_<T> (an empty vector gives the identity)."""
        if getattr(op, "opclass", None) == "Aggregator":   # a recipe over the multiply (graphblas_b200/agg.py)
            return op._scalar_expr(self, False)
        op = operator.monoid.plus if op is None else op
        op = operator.get_typed_op(op, self.dtype, kind="monoid")
        if op.opclass != "Monoid":
            raise TypeError("reduce expects a Monoid")
        me = self
        if allow_empty:
            def run(out, accum):
                call("GrB_Vector_reduce_Monoid_Scalar", [out, accum, op, me, None])

            return ScalarExpression(op.return_type, run=run)

        def thunk():
            x = op.return_type.ctype(op.return_type.np_type.type(_monoid_identity(op)).item())
            call(f"GrB_Vector_reduce_{op.return_type.name}", [ctypes.byref(x), None, op, me, None])
            return x.value

        return ScalarExpression(op.return_type, thunk)

    # ---- comparison (reference core/vector.py:340-379: eWiseMult(EQ) + reduce(LAND))
    def isequal(self, other, *, check_dtype=False):
        if type(other) is not Vector:
            raise TypeError(f"isequal expects a Vector, got {type(other).__name__}")
        if check_dtype and self.dtype != other.dtype:
            return False
        if self._size != other._size or self.nvals != other.nvals:
            return False
        common = unify(self.dtype, other.dtype)
        matches = self.ewise_mult(other, operator.binary.eq[common]).new(BOOL)
        if matches.nvals != self.nvals:
            return False
        return bool(matches.reduce(operator.monoid.land, allow_empty=False).value)

    def isclose(self, other, *, rel_tol=1e-7, abs_tol=0.0, check_dtype=False):
        if check_dtype and self.dtype != other.dtype:
            return False
        if self._size != other._size or self.nvals != other.nvals:
            return False
        i1, v1 = self.to_coo()
        i2, v2 = other.to_coo()
        return bool(np.array_equal(i1, i2) and np.all(np.isclose(v1.astype(np.float64), v2.astype(np.float64), rtol=rel_tol, atol=abs_tol)))


def _scalar_dtype(x):
    if isinstance(x, (bool, np.bool_)):
        return BOOL
    if isinstance(x, (int, np.integer)) and not isinstance(x, np.generic):
        return INT64
    if isinstance(x, float):
        return FP64
    return lookup_dtype(np.asarray(x).dtype)


def _monoid_identity(op):
    """identity of a builtin monoid in its own type (what GrB_*_reduce_<T> leaves for an empty input)"""
    name, t = op.parent.name, op.return_type.np_type
    if name in ("plus", "lor", "lxor", "any"):
        return 0
    if name in ("times", "land", "lxnor", "eq"):
        return 1
    info = np.finfo(t) if t.kind == "f" else None
    if name == "min":
        return np.inf if info else (True if t.kind == "b" else np.iinfo(t).max)
    if name == "max":
        return -np.inf if info else (False if t.kind == "b" else np.iinfo(t).min)
    return 0


class _CScalar:
    """a C scalar argument passed by value with its C type (the reference passes Scalar(is_cscalar=True)._carg)"""

    def __init__(self, value, dtype):
        self.value, self.dtype = value, dtype
        self.name = repr(value)

    @property
    def _carg(self):
        return self.dtype.ctype(self.dtype.np_type.type(self.value).item())


class _All:
    """GrB_ALL (reference core/utils.py: `_CArray`/lib.GrB_ALL for slice(None))"""
    name = gb_name = "GrB_ALL"

    @property
    def _carg(self):
        return lib().GrB_ALL


class _Ref:
    """&obj for *_new(&obj, ...) (reference core/utils.py:346-370 _Pointer)"""

    def __init__(self, obj):
        self.obj = obj

    @property
    def _carg(self):
        return ctypes.byref(self.obj.gb_obj)

    @property
    def name(self):
        return f"&{self.obj.name}"


class VectorExpression(BaseExpression):
    output_type = Vector

    def __init__(self, *args, size, **kw):
        super().__init__(*args, **kw)
        self._size = size

    def construct_output(self, dtype=None, *, name=None):
        return Vector(dtype or self.dtype, self._size, name=name or None)
