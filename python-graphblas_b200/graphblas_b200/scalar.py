"""Minimal Scalar: results of reduce / element extraction (reference core/scalar.py is out of the hot path)."""
from .dtypes import lookup_dtype


class Scalar:
    _is_scalar = True
    ndim = 0

    def __init__(self, dtype, value=None, name=None):
        self.dtype = lookup_dtype(dtype)
        self._value = None if value is None else self.dtype.np_type.type(value)
        self.name = name or "s"

    @property
    def value(self):
        return None if self._value is None else self._value.item()

    @property
    def is_empty(self):
        return self._value is None

    @property
    def nvals(self):
        return 0 if self._value is None else 1

    def new(self, dtype=None, **kw):
        return self if dtype is None else Scalar(dtype, self.value)

    def __eq__(self, other):
        other = other.value if isinstance(other, Scalar) else other
        return self.value == other

    def __bool__(self):
        return bool(self.value)

    def __repr__(self):
        return f"Scalar({self.value}, dtype={self.dtype})"


class ScalarExpression:
    """Lazy scalar result; `.new()` / `.value` run it (reference core/scalar.py ScalarExpression)."""

    def __init__(self, dtype, thunk):
        self.dtype, self._thunk = lookup_dtype(dtype), thunk

    def new(self, dtype=None, **kw):
        return Scalar(dtype or self.dtype, self._thunk())

    @property
    def value(self):
        return self.new().value
