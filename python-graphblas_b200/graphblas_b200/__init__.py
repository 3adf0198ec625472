"""graphblas_b200 -- host side of the B200-native GraphBLAS semiring engine.

The package mirrors the slice of python-graphblas's public surface that reaches GrB_mxm / GrB_mxv /
GrB_vxm (Matrix / Vector / TransposedMatrix, masks, `C(mask, accum, replace) << expr`, semiring / binary /
monoid / unary namespaces, dtypes) and funnels every operation through one ctypes call into
libgrb_cuda.so -- the same seam the reference uses (graphblas/core/base.py:23-54).  INTEGRATION.md shows
the few lines that bind the same library under the real `graphblas` package as a third backend.
"""
import sys as _sys

from . import dtypes, exceptions
from ._lib import lib as _libfn

backend = "grb_cuda"
_init_params = None


def init(backend="grb_cuda", blocking=False, device=None):
    """reference graphblas/__init__.py:107-199; re-init with different parameters raises."""
    global _init_params
    params = dict(backend=backend, blocking=blocking)
    if _init_params is not None:
        if _init_params != params:
            raise RuntimeError("graphblas_b200 was already initialised with different parameters")
        return
    if backend != "grb_cuda":
        raise ValueError(f"unknown backend {backend!r}; this package provides only 'grb_cuda'")
    L = _libfn()
    if device is not None:
        L.GrB_cuda_set_device(int(device))
    rc = L.GrB_init(1 if blocking else 0)
    if rc != 0:
        raise exceptions.Panic((L.GrB_cuda_last_error() or b"GrB_init failed").decode())
    from . import operator

    operator.initialize()
    _init_params = params


def is_initialized():
    return _init_params is not None


from . import operator as _operator  # noqa: E402

unary, binary, monoid, semiring, select = _operator.unary, _operator.binary, _operator.monoid, _operator.semiring, _operator.select
for _m in (unary, binary, monoid, semiring, select):
    _sys.modules[_m.__name__] = _m

from .base import Recorder, replace  # noqa: E402
from .matrix import Matrix  # noqa: E402
from .scalar import Scalar  # noqa: E402
from .vector import Vector  # noqa: E402
from . import cuda  # noqa: E402
from . import io  # noqa: E402
from . import agg  # noqa: E402

__all__ = ["Matrix", "Vector", "Scalar", "semiring", "binary", "monoid", "unary", "select", "dtypes", "replace", "init", "cuda", "io", "agg"]
