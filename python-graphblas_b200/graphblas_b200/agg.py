"""graphblas_b200.agg -- aggregators: reductions that are RECIPES over the multiply, not new kernels.

The reference defines an aggregator as a short program over the operations this backend already runs on the device
(graphblas/core/operator/agg.py): a monoid reduction, or a semiring multiply against a dense "init" vector (count = plus_pair,
count_nonzero = plus_isne against 0, sum_of_squares = plus_pow against 2, exists = any_pair ...), optionally preceded by a unary
apply (L1norm: abs), followed by a unary finalizer (hypot: sqrt) or combined from other aggregators (mean = sum / count,
varp = <x2>/n - (<x>/n)^2 ...).  The recipes below follow that file (lines cited per aggregator) so results match the reference:

  Matrix row / column wise : step = semiring(A @ init)                         (agg.py:262-283; `switch` swaps the operands)
  Vector -> scalar         : step = semiring(v @ init-column)  == v.inner(init) (agg.py:284-303)
  Matrix -> scalar         : rows first (above), then the Vector recipe with `semiring2` (agg.py:304-333)

Every step is ONE GrB_mxv / GrB_vxm / eWise / apply call into libgrb_cuda.so; nothing is computed on the host except the
arithmetic on the final 0-d scalars.  Use through the usual methods: ``A.reduce_rowwise(agg.mean)``, ``v.reduce(agg.count)``,
``A.reduce_scalar(agg.varp)``; ``C(mask, accum) << A.reduce_rowwise(agg.x)`` works like any other expression.
"""
import math

import numpy as np

from . import dtypes, operator
from .dtypes import FP64, INT64

__all__ = ["Aggregator"]


def _float_type(dt):
    return dt if dt.name in ("FP32", "FP64") else FP64


def _as_float(v):
    """floating-point copy of a Vector (FP32 stays FP32): unary finalizers such as sqrt are typed FP32 / FP64"""
    if v.dtype.name in ("FP32", "FP64"):
        return v
    return v.apply(operator.unary.identity).new(FP64)


def _init_vector(initval, n):
    """the dense `init` operand of the recipe (agg.py:54-56, 279-281): BOOL False when the multiply ignores it, else INT64 / FP64
    by the Python type of the value -- the multiply then unifies it with the data's dtype exactly as the reference does"""
    from .vector import Vector

    if initval is None:
        init = Vector(dtypes.BOOL, n)
        init[:] = False
    else:
        init = Vector(FP64 if isinstance(initval, float) else INT64, n)
        init[:] = initval
    return init


class Aggregator:
    opclass = "Aggregator"

    def __init__(self, name, *, monoid=None, semiring=None, semiring2=None, initval=None, switch=False, applybegin=None,
                 finalize=None, composite=None, combine=None):
        self.name = name
        self._monoid, self._semiring, self._semiring2 = monoid, semiring, semiring2
        self._initval, self._switch, self._applybegin, self._finalize = initval, switch, applybegin, finalize
        self._composite, self._combine = composite, combine

    def __repr__(self):
        return f"agg.{self.name}"

    def _begin(self, x):
        if self._monoid in ("land", "lor") and x.dtype.name != "BOOL":   # logical monoids run in BOOL (x != 0), reference monoid.py:494-534
            x = x.apply(operator.unary.identity).new(dtypes.BOOL)
        return x if self._applybegin is None else x.apply(getattr(operator.unary, self._applybegin)).new()

    # ---- Matrix (or TransposedMatrix) -> Vector: rows without entries have no entry          (reference agg.py:262-283)
    def _rowwise(self, A, finalize=True):
        if self._composite is not None:
            return self._combine([p._rowwise(A) for p in self._composite], vector=True)
        A = self._begin(A)
        if self._monoid is not None:
            return A.reduce_rowwise(getattr(operator.monoid, self._monoid)).new()
        sr = getattr(operator.semiring, self._semiring)
        init = _init_vector(self._initval, A.shape[1])
        step = (init.vxm(A.T, sr) if self._switch else A.mxv(init, sr)).new()   # switch: the init value is the multiply's LEFT operand
        if finalize and self._finalize is not None:
            step = _as_float(step).apply(getattr(operator.unary, self._finalize)).new()
        return step

    # ---- Vector -> Python scalar, None when the vector has no entries                          (reference agg.py:284-303)
    def _vector(self, v, finalize=True, semiring=None):
        if self._composite is not None:
            parts = [p._vector(v) for p in self._composite]
            return None if builtins_any(p is None for p in parts) else self._combine(parts, vector=False)
        if semiring is None:
            v = self._begin(v)
        if self._monoid is not None:
            return v.reduce(getattr(operator.monoid, self._monoid)).new().value
        if v.nvals == 0:
            return None
        sr = getattr(operator.semiring, semiring or self._semiring)
        init = _init_vector(self._initval if semiring is None else None, v.size)
        val = (init.inner(v, sr) if (self._switch and semiring is None) else v.inner(init, sr)).new().value
        if finalize and self._finalize is not None and val is not None:
            val = _SCALAR_UNARY[self._finalize](float(val))
        return val

    # ---- Matrix -> Python scalar: rows first, then the rows' results with `semiring2`          (reference agg.py:304-333)
    def _matrix_scalar(self, A):
        if self._composite is not None:
            parts = [p._matrix_scalar(A) for p in self._composite]
            return None if builtins_any(p is None for p in parts) else self._combine(parts, vector=False)
        if self._monoid is not None:
            return self._begin(A).reduce_scalar(getattr(operator.monoid, self._monoid)).new().value
        step1 = self._rowwise(A, finalize=False)
        val = self._vector(step1, finalize=False, semiring=self._semiring2)
        if self._finalize is not None and val is not None:
            val = _SCALAR_UNARY[self._finalize](float(val))
        return val

    # ---- expressions handed back by Matrix.reduce_* / Vector.reduce
    def _rowwise_expr(self, A):
        return self._rowwise(A)._dup_expr()

    def _scalar_expr(self, obj, is_matrix):
        me = self

        def thunk():
            return me._matrix_scalar(obj) if is_matrix else me._vector(obj)

        return _LazyScalar(thunk)


class _LazyScalar:
    """`.new()` / `.value` of an aggregated scalar; the dtype is whatever the recipe's multiplies produced"""

    def __init__(self, thunk):
        self._thunk = thunk

    def new(self, dtype=None, *, name=None, **kw):
        from .scalar import Scalar

        val = self._thunk()
        if dtype is None:
            dtype = dtypes.BOOL if isinstance(val, (bool, np.bool_)) else INT64 if isinstance(val, (int, np.integer)) else FP64
        return Scalar(dtype, val, name, is_cscalar=True)

    @property
    def value(self):
        return self._thunk()


builtins_any = any
_SCALAR_UNARY = {"sqrt": lambda x: math.nan if x != x or x < 0 else math.sqrt(x),
                 "log": lambda x: -math.inf if x == 0 else (math.nan if x < 0 or x != x else math.log(x)),
                 "log2": lambda x: -math.inf if x == 0 else (math.nan if x < 0 or x != x else math.log2(x))}


# ------------------------------------------------------------------ combine steps of the composite aggregators (agg.py:425-476)
def _ew(a, b, opname):
    return a.ewise_mult(b, getattr(operator.binary, opname)).new()


class _Combine:
    def __init__(self, vec, sca):
        self._vec, self._sca = vec, sca

    def __call__(self, parts, *, vector):
        return self._vec(*parts) if vector else self._sca(*parts)


def _sqrt_vec(x):
    return _as_float(x).apply(operator.unary.sqrt).new()


def _varp_vec(c, x, x2):   # <x2> / n - (<x> / n) ** 2
    left = _ew(x2, c, "truediv")
    right = _ew(x, c, "truediv")
    right = right.apply(operator.binary.pow, right=2).new()
    return _ew(left, right, "minus")


def _vars_vec(c, x, x2):   # <x2> / (n - 1) - <x> ** 2 / (n (n - 1))
    xx = x.apply(operator.binary.pow, right=2).new()
    right = _ew(xx, c, "truediv")
    c1 = c.apply(operator.binary.minus, right=1).new()
    right = _ew(right, c1, "truediv")
    left = _ew(x2, c1, "truediv")
    return _ew(left, right, "minus")


def _div(a, b):
    a, b = float(a), float(b)
    if b == 0:
        return math.nan if a == 0 or a != a else math.copysign(math.inf, a)
    return a / b


def _varp_sca(c, x, x2):
    return _div(x2, c) - _div(x, c) ** 2


def _vars_sca(c, x, x2):
    return _div(x2, c - 1) - _div(_div(float(x) ** 2, c), c - 1)


def _safe_sqrt(x):
    return math.nan if x != x or x < 0 else math.sqrt(x)


def _geo_vec(c, x):
    inv = c.apply(operator.unary.identity).new(FP64).apply(operator.unary.minv).new()
    xf = x.apply(operator.unary.identity).new(FP64) if x.dtype.name not in ("FP32", "FP64") else x
    return _ew(xf, inv, "pow")


# ------------------------------------------------------------------ the aggregators (names and recipes: reference agg.py:347-534)
sum = Aggregator("sum", monoid="plus")
prod = Aggregator("prod", monoid="times")
all = Aggregator("all", monoid="land")
any = Aggregator("any", monoid="lor")
min = Aggregator("min", monoid="min")
max = Aggregator("max", monoid="max")
any_value = Aggregator("any_value", monoid="any")
count = Aggregator("count", semiring="plus_pair", semiring2="plus_first")
count_nonzero = Aggregator("count_nonzero", semiring="plus_isne", semiring2="plus_first", initval=0)
count_zero = Aggregator("count_zero", semiring="plus_iseq", semiring2="plus_first", initval=0)
sum_of_squares = Aggregator("sum_of_squares", semiring="plus_pow", semiring2="plus_first", initval=2)
sum_of_inverses = Aggregator("sum_of_inverses", semiring="plus_pow", semiring2="plus_first", initval=-1.0)
exists = Aggregator("exists", semiring="any_pair", semiring2="any_pair")
hypot = Aggregator("hypot", semiring="plus_pow", semiring2="plus_first", initval=2, finalize="sqrt")
logaddexp = Aggregator("logaddexp", semiring="plus_pow", semiring2="plus_first", initval=math.e, switch=True, finalize="log")
logaddexp2 = Aggregator("logaddexp2", semiring="plus_pow", semiring2="plus_first", initval=2, switch=True, finalize="log2")
L0norm = count_nonzero
L2norm = hypot
L1norm = Aggregator("L1norm", applybegin="abs", semiring="plus_first", semiring2="plus_first")
Linfnorm = Aggregator("Linfnorm", applybegin="abs", semiring="max_first", semiring2="max_first")
mean = Aggregator("mean", composite=[count, sum], combine=_Combine(lambda c, x: _ew(x, c, "truediv"), lambda c, x: _div(x, c)))
peak_to_peak = Aggregator("peak_to_peak", composite=[max, min],
                          combine=_Combine(lambda mx, mn: _ew(mx, mn, "minus"), lambda mx, mn: mx - mn))
varp = Aggregator("varp", composite=[count, sum, sum_of_squares], combine=_Combine(_varp_vec, _varp_sca))
vars = Aggregator("vars", composite=[count, sum, sum_of_squares], combine=_Combine(_vars_vec, _vars_sca))
stdp = Aggregator("stdp", composite=[count, sum, sum_of_squares],
                  combine=_Combine(lambda c, x, x2: _sqrt_vec(_varp_vec(c, x, x2)), lambda c, x, x2: _safe_sqrt(_varp_sca(c, x, x2))))
stds = Aggregator("stds", composite=[count, sum, sum_of_squares],
                  combine=_Combine(lambda c, x, x2: _sqrt_vec(_vars_vec(c, x, x2)), lambda c, x, x2: _safe_sqrt(_vars_sca(c, x, x2))))
geometric_mean = Aggregator("geometric_mean", composite=[count, prod],
                            combine=_Combine(_geo_vec, lambda c, x: float(x) ** (1.0 / c)))
harmonic_mean = Aggregator("harmonic_mean", composite=[count, sum_of_inverses],
                           combine=_Combine(lambda c, x: _ew(c.apply(operator.unary.identity).new(x.dtype), x, "truediv"), lambda c, x: _div(c, x)))
root_mean_square = Aggregator("root_mean_square", composite=[count, sum_of_squares],
                              combine=_Combine(lambda c, x2: _sqrt_vec(_ew(x2, c, "truediv")), lambda c, x2: _safe_sqrt(_div(x2, c))))

_ALL = [sum, prod, all, any, min, max, any_value, count, count_nonzero, count_zero, sum_of_squares, sum_of_inverses, exists, hypot,
        logaddexp, logaddexp2, L1norm, Linfnorm, mean, peak_to_peak, varp, vars, stdp, stds, geometric_mean, harmonic_mean,
        root_mean_square]
