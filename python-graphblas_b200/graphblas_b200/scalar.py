"""Scalar: a GrB_Scalar object (default, like the reference) or a plain C scalar held on the host.

Mirrors reference graphblas/core/scalar.py: ``Scalar(dtype, is_cscalar=False)`` creates a ``GrB_Scalar`` (:76-84), ``value``
reads it with ``GrB_Scalar_extractElement_<T>`` (:199-216) and writes it with ``GrB_Scalar_setElement_<T>`` / ``GrB_Scalar_clear``
(:262-280), ``nvals`` is ``GrB_Scalar_nvals`` (:235-246).  Reductions write into it through
``GrB_{Vector,Matrix}_reduce_Monoid_Scalar`` (core/vector.py:1670, core/matrix.py:2750), with an optional accumulator:
``s(binary.plus) << v.reduce(monoid.plus)``.
"""
import ctypes

from ._lib import GrB_Index
from .dtypes import FP64, lookup_dtype
from .exceptions import NoValue


def _call(name, args):
    from .base import call

    return call(name, args)


class Scalar:
    _is_scalar = True
    ndim = 0
    shape = ()

    def __init__(self, dtype=FP64, value=None, name=None, *, is_cscalar=False):
        self.dtype = lookup_dtype(dtype)
        self.name = name or "s"
        self._is_cscalar = bool(is_cscalar)
        self._value = None
        self.gb_obj = None
        if not self._is_cscalar:
            self.gb_obj = ctypes.c_void_p()
            _call("GrB_Scalar_new", [ctypes.byref(self.gb_obj), self.dtype])
        if value is not None:
            self.value = value

    @classmethod
    def from_value(cls, value, dtype=None, *, is_cscalar=False, name=None):
        """reference core/scalar.py:470-520"""
        if isinstance(value, Scalar):
            return value.dup(dtype, is_cscalar=is_cscalar, name=name)
        if dtype is None:
            import numpy as np

            dtype = lookup_dtype(type(value)) if isinstance(value, (bool, int, float)) else lookup_dtype(np.asarray(value).dtype)
        return cls(dtype, value, name, is_cscalar=is_cscalar)

    def __del__(self):
        h = getattr(self, "gb_obj", None)
        if h is not None and h.value:
            try:
                from ._lib import lib

                lib().GrB_Scalar_free(ctypes.byref(h))
            except Exception:   # interpreter shutdown
                pass

    @property
    def _carg(self):
        if self._is_cscalar:   # a C scalar crosses the boundary by value, typed (reference core/scalar.py `_carg`)
            return self.dtype.ctype(self._value.item() if self._value is not None else 0)
        return self.gb_obj

    @property
    def is_cscalar(self):
        return self._is_cscalar

    @property
    def is_grbscalar(self):
        return not self._is_cscalar

    @property
    def value(self):
        if self._is_cscalar:
            return None if self._value is None else self._value.item()
        x = self.dtype.ctype()
        rv = _call(f"GrB_Scalar_extractElement_{self.dtype.name}", [ctypes.byref(x), self])
        return None if rv is NoValue else x.value

    @value.setter
    def value(self, val):
        if isinstance(val, Scalar):
            val = val.value
        if val is None:
            self.clear()
        elif self._is_cscalar:
            self._value = self.dtype.np_type.type(val)
        else:
            _call(f"GrB_Scalar_setElement_{self.dtype.name}", [self, self.dtype.ctype(self.dtype.np_type.type(val).item())])

    def clear(self):
        if self._is_cscalar:
            self._value = None
        else:
            _call("GrB_Scalar_clear", [self])

    @property
    def is_empty(self):
        return self.nvals == 0

    @property
    def nvals(self):
        if self._is_cscalar:
            return 0 if self._value is None else 1
        n = GrB_Index()
        _call("GrB_Scalar_nvals", [ctypes.byref(n), self])
        return n.value

    def dup(self, dtype=None, *, is_cscalar=None, name=None):
        return Scalar(self.dtype if dtype is None else dtype, self.value, name,
                      is_cscalar=self._is_cscalar if is_cscalar is None else is_cscalar)

    def new(self, dtype=None, **kw):
        return self if dtype is None else self.dup(dtype)

    def wait(self, how="materialize"):
        return self

    # ---- s << expr, s(accum) << expr  (reference core/scalar.py:282-300 via Updater)
    def __call__(self, accum=None, **kw):
        if kw.get("mask") is not None or kw.get("replace"):
            raise TypeError("a Scalar output takes no mask")
        return _ScalarUpdater(self, accum)

    def __lshift__(self, expr):
        _ScalarUpdater(self, None) << expr

    def update(self, expr):
        _ScalarUpdater(self, None) << expr

    def isequal(self, other, *, check_dtype=False):
        other = other if isinstance(other, Scalar) else Scalar.from_value(other, is_cscalar=True) if other is not None else None
        if other is None:
            return self.is_empty
        if check_dtype and self.dtype != other.dtype:
            return False
        return self.value == other.value

    def __eq__(self, other):
        other = other.value if isinstance(other, Scalar) else other
        return self.value == other

    __hash__ = None

    def __bool__(self):
        return bool(self.value)

    def __repr__(self):
        return f"Scalar({self.value}, dtype={self.dtype})"


class _ScalarUpdater:
    def __init__(self, out, accum):
        self.out, self.accum = out, accum

    def __lshift__(self, expr):
        from . import operator

        accum = self.accum
        if accum is not None:
            accum = operator.get_typed_op(accum, self.out.dtype, kind="binary")
            if accum.opclass == "Monoid":
                accum = accum.binaryop
        if isinstance(expr, ScalarExpression):
            expr._run_into(self.out, accum)
        else:
            if accum is not None:
                raise TypeError("accumulating a plain value into a Scalar is not supported")
            self.out.value = expr


class ScalarExpression:
    """Lazy scalar result; ``.new()`` / ``.value`` run it (reference core/scalar.py ScalarExpression).  Two forms: ``thunk``
    returns a Python value (element extraction, inner); ``run(out, accum)`` makes the C call that writes a GrB_Scalar."""

    def __init__(self, dtype, thunk=None, *, run=None):
        self.dtype, self._thunk, self._run = lookup_dtype(dtype), thunk, run

    def _run_into(self, out, accum):
        if self._run is not None and not out._is_cscalar:
            self._run(out, accum)
            return
        val = self.new().value
        if accum is not None and val is not None and out.value is not None:
            raise TypeError("an accumulator needs a GrB_Scalar output and a C call that takes one")
        if val is not None or accum is None:
            out.value = val

    def new(self, dtype=None, *, is_cscalar=False, name=None, **kw):
        if self._run is not None:
            out = Scalar(dtype or self.dtype, name=name, is_cscalar=False)
            self._run(out, None)
            return out
        return Scalar(dtype or self.dtype, self._thunk(), name, is_cscalar=True)

    @property
    def value(self):
        return self.new().value

    def __eq__(self, other):
        return self.value == (other.value if isinstance(other, (Scalar, ScalarExpression)) else other)

    __hash__ = None

    def __bool__(self):
        return bool(self.value)
