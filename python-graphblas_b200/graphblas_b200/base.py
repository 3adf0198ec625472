"""Dispatch seam + the `C(mask, accum, replace) << expr` machinery.

Mirrors reference graphblas/core/base.py: ``call`` (:23-54), ``BaseType.__call__`` (:192-263, mask / accum /
replace parsing), ``BaseType._update`` (:338-514, which assembles ``[C, mask, accum, op, *args, desc]`` and makes
ONE C call), ``BaseExpression`` (:535-616, lazy ``.new()``); ``Updater`` from core/expr.py:404-473; mask flag
classes from core/mask.py:9-203; descriptor lookup from core/descriptor.py:92-156.
"""
import ctypes
from contextvars import ContextVar

from . import operator
from ._lib import NULL, lib
from .dtypes import BOOL, lookup_dtype
from .exceptions import check_status

_recorder = ContextVar("recorder", default=None)


class _ReplaceSingleton:
    def __repr__(self):
        return "replace"


replace = _ReplaceSingleton()


class Recorder:
    """Records every C call as C-like text (reference graphblas/core/recorder.py:34-182)."""

    def __init__(self):
        self.data = []
        self._token = None

    def __enter__(self):
        self._token = _recorder.set(self)
        return self

    def __exit__(self, *exc):
        _recorder.reset(self._token)

    def record(self, cfunc_name, args, exc=None):
        def s(x):
            if x is None:
                return "NULL"
            return getattr(x, "gb_name", None) or getattr(x, "name", None) or str(x)

        text = f"{cfunc_name}({', '.join(s(a) for a in args)});"
        if exc is not None:
            text += f" /* ERROR: {type(exc).__name__} */"
        self.data.append(text)

    def __iter__(self):
        return iter(self.data)


def call(cfunc_name, args):
    """The only place control leaves Python (reference core/base.py:23-54)."""
    call_args = [getattr(x, "_carg", x) if x is not None else NULL for x in args]
    cfunc = getattr(lib(), cfunc_name)
    err_code = cfunc(*call_args)
    rec = _recorder.get()
    try:
        rv = check_status(err_code, args)
    except Exception as exc:
        if rec is not None:
            rec.record(cfunc_name, args, exc=exc)
        raise
    if rec is not None:
        rec.record(cfunc_name, args)
    return rv


# ------------------------------------------------------------------ descriptors (reference core/descriptor.py:51-156)
class Descriptor:
    __slots__ = ("gb_obj", "name", "gb_name")

    def __init__(self, gb_obj, name):
        self.gb_obj, self.name, self.gb_name = gb_obj, name, name

    @property
    def _carg(self):
        return self.gb_obj


_desc_cache = {}


def descriptor_lookup(*, output_replace=False, mask_complement=False, mask_structure=False, transpose_first=False,
                      transpose_second=False, **opts):
    if opts:
        # the reference raises for unknown options on non-suitesparse backends (core/descriptor.py:122-125)
        raise ValueError(
            f"Extra descriptor options not possible with 'grb_cuda' backend; got {', '.join(str(x) for x in opts)}"
        )
    key = (output_replace, mask_structure, mask_complement, transpose_first, transpose_second)
    if not any(key):
        return None
    if key not in _desc_cache:
        name = "GrB_DESC_" + ("R" if key[0] else "") + ("S" if key[1] else "") + ("C" if key[2] else "") + \
            ("T0" if key[3] else "") + ("T1" if key[4] else "")
        _desc_cache[key] = Descriptor(getattr(lib(), name), name)
    return _desc_cache[key]


# ------------------------------------------------------------------ masks (reference core/mask.py:9-203)
class Mask:
    complement = False
    structure = False
    value = False
    __slots__ = ("parent",)

    def __init__(self, parent):
        self.parent = parent

    @property
    def _carg(self):
        return self.parent._carg

    @property
    def name(self):
        return self.parent.name

    @property
    def gb_name(self):
        return self.parent.name

    # ---- mask algebra (reference core/mask.py:205-513: `m1 & m2`, `m1 | m2` give a new mask object).  The reference spells out
    # 16 + 16 recipes; all of them follow from one model: a mask is (P, c) -- the set P of positions where its parent says
    # "true" (the pattern for .S, the entries with a true value for .V) and whether it is complemented:
    #   (P1,0) & (P2,0) = (P1 n P2, 0)   (P1,0) & (P2,1) = (P1 \ P2, 0)   (P1,1) & (P2,1) = (P1 u P2, 1)
    #   (P1,0) | (P2,0) = (P1 u P2, 0)   (P1,0) | (P2,1) = (P2 \ P1, 1)   (P1,1) | (P2,1) = (P1 n P2, 1)
    # and every set operation is ONE device call: n = eWiseMult(pair), u = eWiseAdd(pair), \ = one(P1) under the mask ~P2.S.
    def _true_set(self):
        """BOOL object whose PATTERN is the set of positions where the uncomplemented mask holds"""
        from . import operator

        par = self.parent
        if self.structure:
            return par
        zero = False if par.dtype == BOOL else 0
        return par.select(operator.select.valuene, zero).new()

    def _combine(self, other, is_and):
        from . import operator

        if not isinstance(other, Mask):
            raise TypeError(f"Invalid mask: {type(other)}")
        if type(self.parent) is not type(other.parent) or self.parent.shape != other.parent.shape:
            raise ValueError("masks can only be combined when their parents have the same kind and shape")
        p1, p2, c1, c2 = self._true_set(), other._true_set(), self.complement, other.complement
        pair, one = operator.binary.pair, operator.unary.one

        def inter(a, b):
            return a.ewise_mult(b, pair).new(BOOL)

        def union(a, b):
            return a.ewise_add(b, pair).new(BOOL)

        def minus(a, b):
            return a.apply(one).new(BOOL, mask=ComplementedStructuralMask(b))

        if not c1 and not c2:
            return StructuralMask(inter(p1, p2) if is_and else union(p1, p2))
        if c1 and c2:
            return ComplementedStructuralMask(union(p1, p2) if is_and else inter(p1, p2))
        pos, neg = (p1, p2) if c2 else (p2, p1)          # pos: the uncomplemented operand, neg: the complemented one
        if is_and:
            return StructuralMask(minus(pos, neg))
        return ComplementedStructuralMask(minus(neg, pos))

    def __and__(self, other):
        return self._combine(other, True)

    def __or__(self, other):
        return self._combine(other, False)


class StructuralMask(Mask):
    structure = True

    def __invert__(self):
        return ComplementedStructuralMask(self.parent)


class ValueMask(Mask):
    value = True

    def __invert__(self):
        return ComplementedValueMask(self.parent)


class ComplementedStructuralMask(Mask):
    complement = True
    structure = True

    def __invert__(self):
        return StructuralMask(self.parent)


class ComplementedValueMask(Mask):
    complement = True
    value = True

    def __invert__(self):
        return ValueMask(self.parent)


def _check_mask(mask, output=None):
    if not isinstance(mask, Mask):
        if type(mask).__name__ in {"Vector", "Matrix"}:
            if mask.dtype != BOOL:
                raise TypeError(
                    f"Mask must be boolean objects (got {mask.dtype}) or indicate values (M.V) or structure (M.S)"
                )
            mask = mask.V
        else:
            raise TypeError(f"Invalid mask: {type(mask)}")
    if output is not None and output.ndim == 1 and mask.parent.ndim != 1:
        raise TypeError(f"Mask object must be type Vector; got {type(mask.parent)}")
    return mask


# ------------------------------------------------------------------ collections base
class BaseType:
    ndim = None
    _is_scalar = False

    def __call__(self, *optional_mask_accum_replace, mask=None, accum=None, replace=False, **opts):
        mask_arg = accum_arg = None
        for arg in optional_mask_accum_replace:
            if arg is globals()["replace"]:
                replace = True
            elif isinstance(arg, (BaseType, Mask)) or type(arg).__name__ == "TransposedMatrix":
                if mask_arg is not None:
                    raise TypeError("Got multiple values for argument 'mask'")
                mask_arg = arg
            else:
                if accum_arg is not None:
                    raise TypeError("Got multiple values for argument 'accum'")
                if isinstance(arg, str):
                    arg = operator.from_string(arg, "binary")
                elif operator.find_opclass(arg) is None:
                    raise TypeError(f"Invalid item found in output params: {type(arg)}")
                accum_arg = arg
        if mask_arg is not None and mask is not None:
            raise TypeError("Got multiple values for argument 'mask'")
        if mask_arg is not None:
            mask = mask_arg
        if mask is None:
            if replace:
                raise TypeError("'replace' argument may only be True if a mask is provided")
        else:
            mask = _check_mask(mask)
        if accum_arg is not None:
            if accum is not None:
                raise TypeError("Got multiple values for argument 'accum'")
            accum = accum_arg
        if accum is not None:
            # accumulator is typed by the OUTPUT dtype (reference core/base.py:254-260)
            accum = operator.get_typed_op(accum, self.dtype, kind="binary")
            if accum.opclass == "Monoid":
                accum = accum.binaryop
            elif accum.opclass != "BinaryOp":
                raise TypeError(f"accum must be a BinaryOp or Monoid, not {accum.opclass}")
        return Updater(self, mask=mask, accum=accum, replace=replace, opts=opts)

    def __lshift__(self, expr):
        return self._update(expr, opts={})

    def update(self, expr, **opts):
        return self._update(expr, opts=opts)

    def __matmul__(self, other):
        from .infix import matmul

        return matmul(self, other)

    def _update(self, expr, mask=None, accum=None, replace=False, *, opts):
        from .infix import MatMulExpr

        if isinstance(expr, MatMulExpr):
            expr = expr._to_expr()
        if not isinstance(expr, BaseExpression):
            if type(expr) is type(self):
                expr = expr._dup_expr()            # w << v  (simple assignment)
            elif type(expr).__name__ == "TransposedMatrix" and type(self).__name__ == "Matrix":
                expr = expr._transpose_expr()
            else:
                expr = self._scalar_assign_expr(expr)
        if expr.output_type is not type(self):
            raise TypeError(f"Expected {type(self).__name__} expression, got {expr.output_type.__name__}")
        if mask is None:
            complement = structure = False
        else:
            mask = _check_mask(mask, self)
            complement, structure = mask.complement, mask.structure
        desc = descriptor_lookup(transpose_first=expr.at, transpose_second=expr.bt, mask_complement=complement,
                                 mask_structure=structure, output_replace=replace, **opts)
        if expr.custom is not None:
            return expr.custom(self, mask, accum, desc)
        args = [self, mask, accum]
        if expr.op is not None:
            args.append(expr.op)
        args.extend(expr.args)
        args.append(desc)
        call(expr.cfunc_name, args)
        self._changed()

    def _changed(self):
        pass


class Updater:
    """reference core/expr.py:404-473"""

    __slots__ = ("parent", "kwargs", "opts")

    def __init__(self, parent, *, opts, **kwargs):
        self.parent, self.kwargs, self.opts = parent, kwargs, opts

    def __lshift__(self, expr):
        self.parent._update(expr, **self.kwargs, opts=self.opts)

    def update(self, expr):
        self.parent._update(expr, **self.kwargs, opts=self.opts)

    def __setitem__(self, key, value):
        # only `w(mask...)[:] = scalar` / `[...]` / `C(mask...)[:, :] = scalar` is on the path (GrB_*_assign_<T> with GrB_ALL)
        if key not in (Ellipsis, slice(None), (slice(None), slice(None))):
            raise NotImplementedError("only [:] assignment is supported by this backend")
        self.parent._update(self.parent._scalar_assign_expr(value), **self.kwargs, opts=self.opts)

    def __getitem__(self, key):
        if key not in (Ellipsis, slice(None)):
            raise NotImplementedError("only [:] is supported by this backend")
        return _AllIndexer(self)


class _AllIndexer:
    def __init__(self, updater):
        self.updater = updater

    def __lshift__(self, value):
        self.updater[...] = value


class BaseExpression:
    """Lazy expression (reference core/base.py:535-616): nothing runs until `<<` or `.new()`."""

    output_type = None

    def __init__(self, method_name, cfunc_name, args, *, at=False, bt=False, op=None, dtype=None, custom=None, **shape):
        self.method_name, self.cfunc_name, self.args = method_name, cfunc_name, args
        self.at, self.bt, self.op, self.custom = at, bt, op, custom
        self.dtype = op.return_type if dtype is None else dtype
        self.shape = shape

    def new(self, dtype=None, *, mask=None, name=None, **opts):
        output = self.construct_output(dtype, name=name)
        if mask is None:
            output.update(self, **opts)
        else:
            mask = _check_mask(mask, output)
            output(mask=mask, **opts).update(self)
        return output
