"""ctypes binding of libgrb_cuda.so -- the `lib` / `ffi`-like pair a python-graphblas backend provides.

Mirrors what the reference puts on ``graphblas.core`` at graphblas/__init__.py:195-199 (``ffi``, ``lib``,
``NULL``): ``lib.<name>`` resolves C functions and builtin-object handles by their GraphBLAS C-API name
and ``dir(lib)`` enumerates them, because the reference's operator registry scans ``dir(lib)``
(graphblas/core/operator/base.py:690,803-893).

There is NO CPU fallback: a missing library raises at import, a missing GPU raises at ``init``.
"""
import ctypes
import os
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
_SO = pathlib.Path(os.environ.get("GRB_CUDA_LIB") or (_HERE / "libgrb_cuda.so"))   # GRB_CUDA_LIB: an A/B build of the same library

GrB_Index = ctypes.c_uint64
_info_funcs_void_p = {"GrB_cuda_lookup", "GrB_cuda_get_stream"}
_restype_override = {
    "GrB_cuda_lookup": ctypes.c_void_p,
    "GrB_cuda_get_stream": ctypes.c_void_p,
    "GrB_cuda_symbol_names": ctypes.c_size_t,
    "GrB_cuda_kernel_names": ctypes.c_size_t,
    "GrB_cuda_launch_count": ctypes.c_uint64,
    "GrB_cuda_memory_in_use": ctypes.c_size_t,
    "GrB_cuda_last_error": ctypes.c_char_p,
    "GrB_cuda_get_option": ctypes.c_char_p,
}

# GrB_Info / enums (include/grb_cuda.h)
CONSTANTS = dict(
    GrB_SUCCESS=0, GrB_NO_VALUE=1, GrB_UNINITIALIZED_OBJECT=-1, GrB_NULL_POINTER=-2, GrB_INVALID_VALUE=-3,
    GrB_INVALID_INDEX=-4, GrB_DOMAIN_MISMATCH=-5, GrB_DIMENSION_MISMATCH=-6, GrB_OUTPUT_NOT_EMPTY=-7,
    GrB_NOT_IMPLEMENTED=-8, GrB_PANIC=-101, GrB_OUT_OF_MEMORY=-102, GrB_INSUFFICIENT_SPACE=-103,
    GrB_INVALID_OBJECT=-104, GrB_INDEX_OUT_OF_BOUNDS=-105, GrB_EMPTY_OBJECT=-106,
    GrB_NONBLOCKING=0, GrB_BLOCKING=1, GrB_COMPLETE=0, GrB_MATERIALIZE=1,
    GrB_CSR_FORMAT=0, GrB_CSC_FORMAT=1, GrB_COO_FORMAT=2,
    GrB_OUTP=0, GrB_MASK=1, GrB_INP0=2, GrB_INP1=3, GrB_DEFAULT=0, GrB_REPLACE=1, GrB_COMP=2, GrB_TRAN=3,
    GrB_STRUCTURE=4, GrB_INDEX_MAX=(1 << 60) - 1,
)


class Lib:
    """Attribute access to functions / handles / constants of libgrb_cuda.so by C name."""

    def __init__(self, path=_SO):
        if not pathlib.Path(path).exists():
            raise ImportError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). This backend has no CPU fallback."
            )
        self._dll = ctypes.CDLL(str(path))
        self._cache = {}
        self._dll.GrB_cuda_lookup.restype = ctypes.c_void_p
        self._dll.GrB_cuda_lookup.argtypes = [ctypes.c_char_p]
        self._dll.GrB_cuda_symbol_names.restype = ctypes.c_size_t
        need = self._dll.GrB_cuda_symbol_names(None, ctypes.c_size_t(0))
        buf = ctypes.create_string_buffer(need)
        self._dll.GrB_cuda_symbol_names(buf, ctypes.c_size_t(need))
        self._handles = [s.decode() for s in buf.raw.split(b"\0") if s]
        self.path = str(path)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        c = self._cache
        if name in c:
            return c[name]
        if name in CONSTANTS:
            val = CONSTANTS[name]
        else:
            h = self._dll.GrB_cuda_lookup(name.encode())
            if h:
                val = ctypes.c_void_p(h)
            elif name == "GrB_ALL":
                val = ctypes.c_void_p.in_dll(self._dll, "GrB_ALL")
            else:
                try:
                    fn = getattr(self._dll, name)
                except AttributeError:
                    raise AttributeError(f"libgrb_cuda has no symbol {name!r}") from None
                fn.restype = _restype_override.get(name, ctypes.c_int)
                val = fn
        c[name] = val
        return val

    def __dir__(self):
        return sorted(set(self._handles) | set(CONSTANTS) | {"GrB_ALL"})


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = Lib(os.environ.get("GRB_CUDA_LIBRARY", _SO))
    return _lib


NULL = None
