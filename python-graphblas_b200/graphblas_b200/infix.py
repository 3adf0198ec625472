"""`A @ B` sugar: reference graphblas/core/infix.py:546-596 (_matmul_infix_expr) -> *MatMulExpr.
Only the matmul infix is on the hot path; `semiring(A @ B)` and `(A @ B).new()` both resolve to the
method call with the default plus_times (reference core/expr.py:508-512)."""


class MatMulExpr:
    def __init__(self, left, right, method_name):
        self.left, self.right, self.method_name = left, right, method_name

    def _to_expr(self, op=None):
        from . import operator

        op = operator.semiring.plus_times if op is None else op
        return getattr(self.left, self.method_name)(self.right, op)

    def new(self, dtype=None, *, mask=None, name=None, **opts):
        return self._to_expr().new(dtype, mask=mask, name=name, **opts)


def matmul(left, right):
    from .matrix import Matrix, TransposedMatrix
    from .vector import Vector

    lm, rm = isinstance(left, (Matrix, TransposedMatrix)), isinstance(right, (Matrix, TransposedMatrix))
    lv, rv = isinstance(left, Vector), isinstance(right, Vector)
    if lm and rm:
        return MatMulExpr(left, right, "mxm")
    if lm and rv:
        return MatMulExpr(left, right, "mxv")
    if lv and rm:
        return MatMulExpr(left, right, "vxm")
    if lv and rv:
        return MatMulExpr(left, right, "inner")
    raise TypeError(f"unsupported operand types for @: {type(left).__name__} and {type(right).__name__}")
