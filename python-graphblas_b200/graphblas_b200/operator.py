"""Operator registry: typed builtin BinaryOp / Monoid / Semiring / UnaryOp objects discovered by scanning the
symbols the backend library exports, exactly the way the reference does it
(graphblas/core/operator/base.py:803-893 _initialize; regexes semiring.py:185-219, monoid.py:239-255;
typed-op lookup utils.py:60-157; pair -> INT64 binary.py:387-388; BOOL coercions semiring.py:538-547).

A typed op carries (name, type, return_type, gb_obj, gb_name) like reference TypedOpBase
(operator/base.py:527-620); ``gb_obj`` is what crosses the C boundary (``_carg``).
"""
import re
import types

from . import dtypes
from ._lib import lib
from .dtypes import BOOL, INT64, lookup_dtype, unify

_TYPES = "BOOL|INT8|INT16|INT32|INT64|UINT8|UINT16|UINT32|UINT64|FP32|FP64"


class TypedOp:
    __slots__ = ("parent", "name", "type", "return_type", "gb_obj", "gb_name", "opclass")

    def __init__(self, parent, name, type_, return_type, gb_obj, gb_name, opclass):
        self.parent, self.name, self.type, self.return_type = parent, name, type_, return_type
        self.gb_obj, self.gb_name, self.opclass = gb_obj, gb_name, opclass

    @property
    def _carg(self):
        return self.gb_obj

    def __repr__(self):
        return f"{self.opclass.lower()}.{self.name}[{self.type}]"

    # decomposition like reference TypedBuiltinSemiring.monoid / .binaryop
    @property
    def monoid(self):
        return self.parent.monoid[self.return_type] if self.opclass == "Semiring" else None

    @property
    def binaryop(self):
        if self.opclass == "Semiring":
            return self.parent.binaryop[self.type]
        if self.opclass == "Monoid":
            return self.parent.binaryop[self.type]
        return self

    def __call__(self, expr):
        return self.parent(expr, _typed=self)


class OpBase:
    opclass = None

    def __init__(self, name):
        self.name = name
        self._typed_ops = {}
        self.types = {}
        self.coercions = {}
        self._custom_dtype = None

    def _add(self, typed):
        self._typed_ops[typed.type] = typed
        self.types[typed.type] = typed.return_type

    def __getitem__(self, dtype):
        dtype = lookup_dtype(dtype)
        if dtype not in self._typed_ops:
            raise KeyError(f"{self.name} does not work with {dtype}")
        return self._typed_ops[dtype]

    def __contains__(self, dtype):
        try:
            return lookup_dtype(dtype) in self._typed_ops
        except ValueError:
            return False

    def __repr__(self):
        return f"{self.opclass.lower()}.{self.name}"


class UnaryOp(OpBase):
    opclass = "UnaryOp"


class BinaryOp(OpBase):
    opclass = "BinaryOp"
    monoid = None


class Monoid(OpBase):
    opclass = "Monoid"

    @property
    def binaryop(self):
        return getattr(binary, self.name)


class SelectOp(OpBase):
    """IndexUnaryOps that return BOOL, usable in select (reference core/operator/select.py; builtin names regex-matched from
    dir(lib) at select.py `_parse_config`).  Positional ones (tril, triu, diag, offdiag, colle, colgt, rowle, rowgt) ignore the
    values and take an INT64 thunk; value comparisons (valueeq ... valuele) are typed like the entries."""
    opclass = "SelectOp"
    is_positional = False

    def __call__(self, obj, thunk=None):
        """select.tril(A, -1) == A.select(select.tril, -1) (reference core/operator/select.py:44-61)"""
        return obj.select(self, thunk)


class Semiring(OpBase):
    opclass = "Semiring"

    @property
    def monoid(self):
        return getattr(monoid, self.name.split("_", 1)[0])

    @property
    def binaryop(self):
        return getattr(binary, self.name.split("_", 1)[1])

    def __call__(self, expr, _typed=None):
        """semiring.min_plus(A @ B): reference operator/semiring.py:43-49 -> _call_op."""
        from .infix import MatMulExpr

        if not isinstance(expr, MatMulExpr):
            raise TypeError(f"semiring.{self.name}(...) expects a matmul infix expression (A @ B)")
        return expr._to_expr(_typed if _typed is not None else self)


unary = types.ModuleType("graphblas_b200.unary")
binary = types.ModuleType("graphblas_b200.binary")
monoid = types.ModuleType("graphblas_b200.monoid")
semiring = types.ModuleType("graphblas_b200.semiring")
select = types.ModuleType("graphblas_b200.select")

_RENAME_BINARY = {"oneb": "pair", "div": "cdiv", "rdiv": "rcdiv"}
_STRING_BINARY = {"+": "plus", "-": "minus", "*": "times", "/": "truediv", "&": "land", "|": "lor", "^": "lxor",
                  "==": "eq", "!=": "ne", "<": "lt", ">": "gt", "<=": "le", ">=": "ge"}
_initialized = False


def _get(ns, cls, name):
    op = getattr(ns, name, None)
    if op is None:
        op = cls(name)
        setattr(ns, name, op)
    return op


def initialize():
    """Populate the namespaces from dir(lib).  Called from graphblas_b200.init()."""
    global _initialized
    if _initialized:
        return
    L = lib()
    names = dir(L)
    re_sr_grb = re.compile(rf"^GrB_(PLUS|MIN|MAX|LOR|LAND|LXOR|LXNOR)_([A-Z]+)_SEMIRING_({_TYPES})$")
    re_sr_gxb = re.compile(rf"^GxB_(PLUS|TIMES|MIN|MAX|ANY|LOR|LAND|LXOR|EQ)_([A-Z]+)_({_TYPES})$")
    re_mon_grb = re.compile(rf"^GrB_(PLUS|TIMES|MIN|MAX|LOR|LAND|LXOR|LXNOR)_MONOID_({_TYPES})$")
    re_mon_gxb = re.compile(rf"^GxB_(ANY|EQ)_({_TYPES})_MONOID$")
    re_bin = re.compile(rf"^G[rx]B_(FIRST|SECOND|ONEB|PAIR|MIN|MAX|PLUS|MINUS|RMINUS|TIMES|DIV|RDIV|ANY|LOR|LAND|LXOR|"
                        rf"ISEQ|ISNE|ISGT|ISLT|ISGE|ISLE|POW|EQ|NE|GT|LT|GE|LE)_({_TYPES})$")
    re_bin_bool = re.compile(r"^GrB_(LOR|LAND|LXOR|LXNOR)$")
    re_un = re.compile(rf"^G[rx]B_(IDENTITY|AINV|MINV|ABS|ONE|LNOT|BNOT|SQRT|EXP|LOG|EXP2|LOG2|LOG10|FLOOR|CEIL|ROUND|TRUNC|SIGNUM)_({_TYPES})$")
    re_sel_pos = re.compile(r"^GrB_(TRIL|TRIU|DIAG|OFFDIAG|COLLE|COLGT|ROWLE|ROWGT)$")
    re_sel_val = re.compile(rf"^GrB_(VALUEEQ|VALUENE|VALUEGT|VALUEGE|VALUELT|VALUELE)_({_TYPES})$")
    semiring_names = set()
    for n in names:
        m = re_sr_grb.match(n) or re_sr_gxb.match(n)
        if m and not re_mon_gxb.match(n):
            add, mul, t = m.groups()
            add = {"LXNOR": "lxnor", "EQ": "lxnor"}.get(add, add.lower())
            mul = {"ONEB": "pair"}.get(mul, mul.lower())
            mul = _RENAME_BINARY.get(mul, mul)
            op = _get(semiring, Semiring, f"{add}_{mul}")
            dt = lookup_dtype(t)
            if dt not in op._typed_ops or n.startswith("GrB_"):
                ret = BOOL if mul in ("eq", "ne", "gt", "lt", "ge", "le") else dt   # comparison multiplies: T x T -> BOOL
                op._add(TypedOp(op, op.name, dt, ret, getattr(L, n), n, "Semiring"))
            semiring_names.add(op.name)
            continue
        m = re_mon_grb.match(n) or re_mon_gxb.match(n)
        if m:
            o, t = m.groups()
            o = {"EQ": "lxnor"}.get(o, o.lower())
            op = _get(monoid, Monoid, o)
            dt = lookup_dtype(t)
            op._add(TypedOp(op, op.name, dt, dt, getattr(L, n), n, "Monoid"))
            continue
        m = re_bin.match(n)
        if m:
            o, t = m.groups()
            o = _RENAME_BINARY.get(o.lower(), o.lower())
            op = _get(binary, BinaryOp, o)
            dt = lookup_dtype(t)
            ret = BOOL if o in ("eq", "ne", "gt", "lt", "ge", "le") else dt
            if dt not in op._typed_ops or n.startswith("GrB_"):
                op._add(TypedOp(op, op.name, dt, ret, getattr(L, n), n, "BinaryOp"))
            continue
        m = re_bin_bool.match(n)
        if m:
            o = m.group(1).lower()
            op = _get(binary, BinaryOp, o)
            op._add(TypedOp(op, op.name, BOOL, BOOL, getattr(L, n), n, "BinaryOp"))
            continue
        m = re_un.match(n)
        if m:
            o, t = m.groups()
            op = _get(unary, UnaryOp, o.lower())
            dt = lookup_dtype(t)
            op._add(TypedOp(op, op.name, dt, dt, getattr(L, n), n, "UnaryOp"))
            continue
        m = re_sel_pos.match(n)
        if m:   # one C object serves every entry type
            op = _get(select, SelectOp, m.group(1).lower())
            op.is_positional = True
            for dt in dtypes._ALL:
                op._add(TypedOp(op, op.name, dt, BOOL, getattr(L, n), n, "SelectOp"))
            continue
        m = re_sel_val.match(n)
        if m:
            o, t = m.groups()
            op = _get(select, SelectOp, o.lower())
            dt = lookup_dtype(t)
            op._add(TypedOp(op, op.name, dt, BOOL, getattr(L, n), n, "SelectOp"))
    if hasattr(unary, "lnot") and BOOL not in unary.lnot._typed_ops:
        unary.lnot._add(TypedOp(unary.lnot, "lnot", BOOL, BOOL, L.GrB_LNOT, "GrB_LNOT", "UnaryOp"))
    # ---- coercions (reference operator/semiring.py:468-588, binary.py:387-388)
    notbool = [d for d in dtypes._ALL if d is not BOOL]
    # pair: always INT64 unless BOOL-only semiring
    binary.pair._custom_dtype = lambda op, d1, d2: op[INT64]
    for name in semiring_names:
        op = getattr(semiring, name)
        add, mul = name.split("_", 1)
        if mul == "pair":
            op._custom_dtype = lambda op, d1, d2: op[INT64] if INT64 in op._typed_ops else op[BOOL]
        if BOOL in op._typed_ops and (add in ("lor", "land", "lxor", "lxnor") or mul in ("lor", "land", "lxor", "lxnor")):
            for dt in notbool:          # non-bool inputs run in BOOL
                if dt not in op._typed_ops:
                    op._typed_ops[dt] = op._typed_ops[BOOL]
                    op.types[dt] = BOOL
                    op.coercions[dt] = BOOL
    # boolean forms of arithmetic semirings (reference semiring.py:569-587): plus->lor, times->land, min->land, max->lor
    bool_alias = {"plus": "lor", "times": "land", "min": "land", "max": "lor"}
    for name in list(semiring_names):
        op = getattr(semiring, name)
        add, mul = name.split("_", 1)
        target = f"{bool_alias.get(add, add)}_{bool_alias.get(mul, mul)}"
        if BOOL not in op._typed_ops and hasattr(semiring, target) and BOOL in getattr(semiring, target)._typed_ops:
            op._typed_ops[BOOL] = getattr(semiring, target)._typed_ops[BOOL]
            op.types[BOOL] = BOOL
            op.coercions[BOOL] = BOOL
    for mname, target in (("plus", "lor"), ("times", "land"), ("min", "land"), ("max", "lor")):
        m, t = getattr(monoid, mname), getattr(monoid, target)
        if BOOL not in m._typed_ops:
            m._typed_ops[BOOL] = t._typed_ops[BOOL]
            m.types[BOOL] = BOOL
    # logical binary ops accept any input type by running in BOOL
    for name in ("lor", "land", "lxor", "lxnor"):
        op = getattr(binary, name, None)
        if op is not None and BOOL in op._typed_ops:
            for dt in notbool:
                if dt not in op._typed_ops:
                    op._typed_ops[dt] = op._typed_ops[BOOL]
                    op.types[dt] = BOOL
    # truediv: always floating point (reference binary.truediv): FP types use GrB_DIV, everything else runs in FP64
    td = _get(binary, BinaryOp, "truediv")
    for dt in dtypes._ALL:
        src = binary.cdiv._typed_ops[dtypes.FP32 if dt is dtypes.FP32 else dtypes.FP64]
        td._typed_ops[dt] = TypedOp(td, "truediv", src.type, src.return_type, src.gb_obj, src.gb_name, "BinaryOp")
        td.types[dt] = src.return_type
    _initialized = True


# reference core/operator/select.py `_str_to_select`
_STRING_SELECT = {"==": "valueeq", "!=": "valuene", ">": "valuegt", ">=": "valuege", "<": "valuelt", "<=": "valuele",
                  "col<=": "colle", "col>": "colgt", "row<=": "rowle", "row>": "rowgt", "index<=": "rowle", "index>": "rowgt"}


def from_string(string, kind):
    ns = {"binary": binary, "monoid": monoid, "semiring": semiring, "unary": unary, "select": select}[kind]
    s = string.strip()
    dtype = None
    m = re.match(r"^(.*)\[(\w+)\]$", s)
    if m:
        s, dtype = m.group(1).strip(), m.group(2)
    if kind in ("binary", "monoid"):
        s = _STRING_BINARY.get(s, s)
    if kind == "select":
        s = _STRING_SELECT.get(s.replace(" ", ""), s).lower()
    if kind == "semiring" and "." in s:
        a, b = s.split(".", 1)
        s = f"{_STRING_BINARY.get(a, a)}_{_STRING_BINARY.get(b, b)}"
    op = getattr(ns, s, None)
    if op is None:
        raise ValueError(f"Unknown {kind} string: {string!r}")
    return op[dtype] if dtype else op


def find_opclass(op):
    return getattr(op, "opclass", None)


def get_typed_op(op, dtype, dtype2=None, *, kind=None):
    """reference graphblas/core/operator/utils.py:60-157 (builtin ops only)."""
    if isinstance(op, TypedOp):
        return op
    if isinstance(op, str):
        op = from_string(op, kind)
        if isinstance(op, TypedOp):
            return op
    if not isinstance(op, OpBase):
        raise TypeError(f"Unable to get typed operator from object with type {type(op)}")
    if dtype2 is None or getattr(op, "is_positional", False):
        return op[dtype]
    if op._custom_dtype is not None:
        rv = op._custom_dtype(op, dtype, dtype2)
        if rv is not None:
            return rv
    return op[unify(dtype, dtype2)]
