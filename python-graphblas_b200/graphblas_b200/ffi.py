"""A cffi-shaped ``(ffi, lib)`` pair over libgrb_cuda.so: the adapter python-graphblas needs to load this library as a backend.

The reference keeps its C binding in two objects placed on ``graphblas.core`` (graphblas/__init__.py:195-199):

* ``lib`` -- attribute access to every C function and builtin object by its C-API name, ``dir(lib)`` enumerating them
  (the operator registry regex-scans it: graphblas/core/operator/base.py:690, 803-893);
* ``ffi`` -- ``new / cast / from_buffer / string / sizeof / NULL`` (graphblas/core/{matrix,vector,scalar}.py ``ffi_new``,
  core/utils.py:327 ``_CArray``, exceptions.py:186-188).

Both are cffi objects there (``from suitesparse_graphblas import ffi, lib``, graphblas/__init__.py:143).  Here they are
ctypes objects with the same spelling.  Like cffi, the function signatures come from the C header itself:
``include/grb_cuda.h`` is parsed (its two ``#define``-d declaration families expanded) and every function gets ``argtypes``, so
plain Python ints become ``GrB_Index`` (uint64), floats the right C floating type, and ``ffi.new("GrB_Matrix*")`` cells are
passed by reference -- exactly what ``call()`` (core/base.py:23-54) hands over.

    from graphblas_b200.ffi import ffi, lib            # or install(graphblas.core) inside graphblas._init
    A = ffi.new("GrB_Matrix*")
    lib.GrB_Matrix_new(A, lib.GrB_FP64, 3, 3)
    lib.GrB_mxm(C[0], ffi.NULL, ffi.NULL, lib.GrB_PLUS_TIMES_SEMIRING_FP64, A[0], A[0], ffi.NULL)

There is no CPU fallback behind it: every compute entry point needs a GPU (``GrB_init`` fails loudly without one).
"""
import ctypes
import pathlib
import re

from . import _lib as _base

_HEADER = pathlib.Path(__file__).resolve().parents[2] / "include" / "grb_cuda.h"

_HANDLES = ("GrB_Type", "GrB_UnaryOp", "GrB_BinaryOp", "GrB_Monoid", "GrB_Semiring", "GrB_Descriptor", "GrB_Matrix", "GrB_Vector",
            "GrB_Scalar", "GrB_IndexUnaryOp")
_ENUMS = ("GrB_Info", "GrB_Mode", "GrB_WaitMode", "GrB_Format", "GrB_Desc_Field", "GrB_Desc_Value")
_SCALARS = {
    "bool": ctypes.c_bool, "int8_t": ctypes.c_int8, "int16_t": ctypes.c_int16, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64,
    "uint8_t": ctypes.c_uint8, "uint16_t": ctypes.c_uint16, "uint32_t": ctypes.c_uint32, "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float, "double": ctypes.c_double, "int": ctypes.c_int, "unsigned int": ctypes.c_uint, "size_t": ctypes.c_size_t,
    "GrB_Index": ctypes.c_uint64, "char": ctypes.c_char,
}
for _e in _ENUMS:
    _SCALARS[_e] = ctypes.c_int
for _h in _HANDLES:
    _SCALARS[_h] = ctypes.c_void_p


# ------------------------------------------------------------------ header -> signatures
def _expand_macros(text):
    """Expands the function-like declaration macros of the header (GRB_CUDA_DECLARE_TYPED*), which use ## pasting."""
    text = text.replace("\\\n", " ")
    macros = {}
    for m in re.finditer(r"^[ \t]*#define[ \t]+(\w+)\(([^)]*)\)[ \t]+(.*)$", text, re.M):
        macros[m.group(1)] = ([p.strip() for p in m.group(2).split(",")], m.group(3))
    text = re.sub(r"^[ \t]*#.*$", "", text, flags=re.M)

    def repl(m):
        params, body = macros[m.group(1)]
        args = [a.strip() for a in m.group(2).split(",")]
        for p, a in zip(params, args):
            body = re.sub(rf"\s*##\s*{p}\b", a, body)
            body = re.sub(rf"\b{p}\b", a, body)
        return body

    for name in macros:
        text = re.sub(rf"\b({name})\(([^()]*)\)", repl, text)
    return text


def _ctype_of(decl):
    """C parameter / return declaration -> ctypes type (every pointer is passed as void*, handles included)."""
    decl = re.sub(r"/\*.*?\*/", " ", decl)
    decl = re.sub(r"\bconst\b", " ", decl).strip()
    if decl == "void":
        return None
    if "[" in decl and "*" not in decl:   # array parameter: decays to a pointer
        return ctypes.c_void_p
    if "*" in decl:
        base = decl.split("*")[0].strip()
        return ctypes.c_char_p if base == "char" and decl.count("*") == 1 else ctypes.c_void_p
    words = decl.split()
    for cut in (len(words), len(words) - 1):   # with or without the parameter name
        t = " ".join(words[:cut])
        if t in _SCALARS:
            return _SCALARS[t]
    raise ValueError(f"grb_cuda.h: cannot map C declaration {decl!r}")


def parse_header(path=_HEADER):
    """{function name: (restype, [argtypes])} for every function the header declares."""
    text = re.sub(r"/\*.*?\*/", " ", pathlib.Path(path).read_text(), flags=re.S)
    text = _expand_macros(text)
    sigs = {}
    for m in re.finditer(r"([\w \t\*]+?)\b(Gr[Bx]_\w+)\s*\(([^()]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret or "extern" in ret:
            continue
        args = [] if params in ("", "void") else [_ctype_of(p) for p in params.split(",")]
        sigs[name] = (_ctype_of(ret + " x") if ret != "void" else None, args)
    return sigs


# ------------------------------------------------------------------ ffi
class _Cell:
    """``ffi.new("T*")``: one element of T, indexable with [0], passed to C by reference."""

    def __init__(self, ctype, init=None, _storage=None):
        self._ctype = ctype
        self._c = _storage if _storage is not None else (ctype() if init is None else ctype(init))

    def __getitem__(self, i):
        if i != 0:
            raise IndexError(i)
        return self._c.value

    def __setitem__(self, i, value):
        if i != 0:
            raise IndexError(i)
        self._c.value = getattr(value, "value", value)

    @property
    def _as_parameter_(self):
        return ctypes.byref(self._c)

    def __repr__(self):
        return f"<cell {self._ctype.__name__} {self._c.value!r}>"


class FFI:
    NULL = None

    @staticmethod
    def _base(ctype):
        base = ctype.replace("const", "").strip()
        if base not in _SCALARS:
            raise TypeError(f"unknown C type {ctype!r}")
        return _SCALARS[base]

    def new(self, ctype, init=None):
        ctype = ctype.strip()
        m = re.fullmatch(r"(.+?)\s*\[(\d*)\]", ctype)
        if m:   # arrays: "T[]" with a length or an initialiser, "T[n]"
            base = self._base(m.group(1))
            if base is ctypes.c_char and isinstance(init, (bytes, bytearray)):
                return ctypes.create_string_buffer(bytes(init))
            if m.group(2):
                return (base * int(m.group(2)))()
            if isinstance(init, int):
                return (base * init)()
            init = list(init)
            return (base * len(init))(*init)
        if ctype.endswith("**") and ctype[:-2].strip() == "char":
            return _Cell(ctypes.c_char_p)
        if ctype.endswith("*"):
            return _Cell(self._base(ctype[:-1]), init)
        return _Cell(self._base(ctype), init)   # "GrB_Index" etc.: cffi also hands back a 1-element owner

    def cast(self, ctype, value):
        ctype = ctype.strip()
        if isinstance(value, _Cell) and ctype.endswith("*"):   # view a handle cell as another handle type (Vector <-> Matrix)
            return _Cell(self._base(ctype[:-1]), _storage=value._c)
        if ctype.endswith("*") or ctype.endswith("]"):
            return value if isinstance(value, ctypes.c_void_p) else ctypes.c_void_p(getattr(value, "value", value))
        return self._base(ctype)(getattr(value, "value", value))

    def from_buffer(self, ctype, array=None, require_writable=False):
        array = ctype if array is None else array
        if hasattr(array, "ctypes"):   # numpy
            p = ctypes.c_void_p(array.ctypes.data)
        else:
            p = ctypes.c_void_p(ctypes.addressof(ctypes.c_char.from_buffer(array)))
        p._keepalive = array
        return p

    def string(self, cdata, maxlen=-1):
        if cdata is None:
            raise RuntimeError("ffi.string(NULL)")
        if isinstance(cdata, (bytes, bytearray)):
            return bytes(cdata)
        if isinstance(cdata, ctypes.Array):
            return cdata.value
        return ctypes.string_at(getattr(cdata, "value", cdata), maxlen)

    def sizeof(self, ctype):
        if isinstance(ctype, str):
            return ctypes.sizeof(ctypes.c_void_p if ctype.strip().endswith("*") else self._base(ctype))
        return ctypes.sizeof(getattr(ctype, "_c", ctype))

    def buffer(self, cdata, size=-1):
        c = getattr(cdata, "_c", cdata)
        n = ctypes.sizeof(c) if size < 0 else size
        return (ctypes.c_char * n).from_address(ctypes.addressof(c))


class CLib:
    """``lib``: C functions (with the header's signatures), builtin-object handles and enum constants by C name."""

    def __init__(self, raw=None, header=_HEADER):
        self._raw = raw if raw is not None else _base.lib()
        self._sigs = parse_header(header)
        self._cache = {}

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name in self._cache:
            return self._cache[name]
        val = getattr(self._raw, name)   # AttributeError for unknown names, like cffi
        if name in self._sigs and callable(val):
            restype, argtypes = self._sigs[name]
            val.restype = restype
            val.argtypes = argtypes
            val.__name__ = name
        self._cache[name] = val
        return val

    def __dir__(self):
        return sorted(set(dir(self._raw)) | set(self._sigs))


ffi = FFI()
_clib = None


def __getattr__(name):   # module-level `lib`, created on first use (importing this module must not need the .so)
    global _clib
    if name == "lib":
        if _clib is None:
            _clib = CLib()
        return _clib
    raise AttributeError(name)


def install(core, blocking=False):
    """What ``graphblas._init`` does for a backend (reference graphblas/__init__.py:189-199): initialise the library and
    put ``ffi`` / ``lib`` / ``NULL`` on the ``graphblas.core`` module object passed in."""
    lib = __getattr__("lib")
    info = lib.GrB_init(lib.GrB_BLOCKING if blocking else lib.GrB_NONBLOCKING)
    if info != lib.GrB_SUCCESS:
        raise RuntimeError((lib.GrB_cuda_last_error() or b"GrB_init failed").decode())
    core.ffi, core.lib, core.NULL = ffi, lib, ffi.NULL
    return core
