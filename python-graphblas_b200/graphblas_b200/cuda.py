"""Device-side extensions of the grb_cuda backend (the analogue of the reference's ``A.ss`` / ``gb.ss``
namespace: graphblas/core/ss/matrix.py, graphblas/ss/_core.py): device-pointer import/export, stream
binding, timers, options, kernel accounting.  torch is used only as a device-memory / stream carrier."""
import ctypes

import numpy as np

from ._lib import GrB_Index, lib
from .base import call
from .dtypes import lookup_dtype


def set_option(key, value):
    lib().GrB_cuda_set_option(str(key).encode(), None if value is None else str(value).encode())


def get_option(key):
    return (lib().GrB_cuda_get_option(str(key).encode()) or b"").decode()


def sync():
    call("GrB_cuda_sync", [])


def launch_count():
    return int(lib().GrB_cuda_launch_count())


def memory_in_use():
    return int(lib().GrB_cuda_memory_in_use())


def use_torch_stream():
    """Bind the library to torch's current CUDA stream so torch events / NCCL order with our kernels."""
    import torch

    h = torch.cuda.current_stream().cuda_stream
    call("GrB_cuda_set_stream", [ctypes.c_void_p(h if h else 1)])   # 1 == cudaStreamLegacy (the NULL stream)


def _torch_stream_handle():
    import torch

    h = torch.cuda.current_stream().cuda_stream
    return ctypes.c_void_p(h if h else 1)


def _wait_for_torch():
    """library stream waits for torch's current stream (torch-produced inputs are complete before we read them)"""
    call("GrB_cuda_stream_order", [_torch_stream_handle(), 0])


def _torch_waits():
    """torch's current stream waits for the library stream (our reads of torch buffers finish before torch reuses them;
    library-owned arrays are complete before torch reads them)"""
    call("GrB_cuda_stream_order", [_torch_stream_handle(), 1])


def timer_start():
    call("GrB_cuda_timer_start", [])


def timer_stop():
    ms = ctypes.c_float()
    call("GrB_cuda_timer_stop", [ctypes.byref(ms)])
    return ms.value


def kernel_times(reset=False):
    L = lib()
    need = L.GrB_cuda_kernel_names(None, ctypes.c_size_t(0))
    buf = ctypes.create_string_buffer(max(int(need), 1))
    L.GrB_cuda_kernel_names(buf, ctypes.c_size_t(need))
    out = {}
    for name in [s for s in buf.raw[:need].split(b"\0") if s]:
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        L.GrB_cuda_kernel_time(name, ctypes.byref(ms), ctypes.byref(n))
        out[name.decode()] = (ms.value, n.value)
    if reset:
        L.GrB_cuda_kernel_time(None, None, None)
    return out


def matrix_from_device_csr(indptr, col_indices, values, nrows, ncols, *, sorted=True, name=None):
    """indptr int64 / col_indices int32 / values: torch CUDA tensors (copied, not adopted)."""
    from .matrix import Matrix

    dtype = lookup_dtype(str(values.dtype).replace("torch.", "").replace("float32", "FP32").replace("float64", "FP64")
                         if not str(values.dtype).startswith("torch.") else _torch_dtype(values.dtype))
    assert indptr.dtype.is_floating_point is False and indptr.element_size() == 8
    assert col_indices.element_size() == 4
    h = ctypes.c_void_p()
    out = Matrix._from_handle(h, dtype, nrows, ncols, name)
    _wait_for_torch()
    call("GrB_cuda_Matrix_import_csr32",
         [ctypes.byref(h), dtype, GrB_Index(nrows), GrB_Index(ncols), ctypes.c_void_p(indptr.data_ptr()),
          ctypes.c_void_p(col_indices.data_ptr()), ctypes.c_void_p(values.data_ptr()), GrB_Index(col_indices.numel()),
          1, 1 if sorted else 0])
    _torch_waits()
    return out


def matrix_from_host_csr32(indptr, col_indices, values, nrows, ncols, *, sorted=True, name=None):
    """host numpy arrays (int64 / int32 / typed; ideally pinned) -> device matrix, no uint64 widening."""
    from .matrix import Matrix

    dtype = lookup_dtype(values.dtype)
    h = ctypes.c_void_p()
    out = Matrix._from_handle(h, dtype, nrows, ncols, name)
    call("GrB_cuda_Matrix_import_csr32",
         [ctypes.byref(h), dtype, GrB_Index(nrows), GrB_Index(ncols), ctypes.c_void_p(indptr.ctypes.data),
          ctypes.c_void_p(col_indices.ctypes.data), ctypes.c_void_p(values.ctypes.data), GrB_Index(col_indices.shape[0]),
          0, 1 if sorted else 0])
    return out


def matrix_export_host_csr32(A, indptr, col_indices, values, *, sort=False):
    call("GrB_cuda_Matrix_export_csr32",
         [ctypes.c_void_p(indptr.ctypes.data) if indptr is not None else None,
          ctypes.c_void_p(col_indices.ctypes.data) if col_indices is not None else None,
          ctypes.c_void_p(values.ctypes.data) if values is not None else None,
          GrB_Index(col_indices.shape[0] if col_indices is not None else A.nvals), A, 1 if sort else 0])


def matrix_export_host_csr32_async(A, indptr, col_indices, values):
    """enqueue the D2H copies of A's CSR on the library's copy stream (they overlap later computation); the arrays must be
    pinned and, like A, stay alive until copy_sync()"""
    call("GrB_cuda_Matrix_export_csr32_async",
         [ctypes.c_void_p(indptr.ctypes.data) if indptr is not None else None,
          ctypes.c_void_p(col_indices.ctypes.data) if col_indices is not None else None,
          ctypes.c_void_p(values.ctypes.data) if values is not None else None,
          GrB_Index(col_indices.shape[0] if col_indices is not None else A.nvals), A])


def copy_sync():
    call("GrB_cuda_copy_sync", [])


def mxm_to_host_csr32(A, B, semiring, out_indptr, out_cols, out_vals, *, blocks=8):
    """C = A (+).(x) B straight into host CSR arrays (pinned numpy views: int64 / int32 / values), formed in `blocks` row
    blocks of A with (nearly) equal entry counts: block k leaves the device over PCIe while block k + 1 is being multiplied,
    so the call takes about max(D2H time, multiply time) instead of their sum.  Rows come out unsorted ("jumbled"), like
    every mxm result before it is sorted on demand.  Returns nvals(C)."""
    import torch

    ip, cj, cx = matrix_as_torch(A, sync=False)
    m, n = A.nrows, A.ncols
    nnz = int(ip[-1])
    targets = torch.tensor([nnz * k // blocks for k in range(blocks + 1)], dtype=torch.int64, device=ip.device)
    bounds = torch.searchsorted(ip, targets).tolist()
    bounds[0], bounds[-1] = 0, m
    for k in range(1, len(bounds)):
        bounds[k] = max(bounds[k], bounds[k - 1])
    offs = []
    off = 0
    prev = None   # (block result, copy-stream ticket) of the block that is leaving the device
    for q0, q1 in zip(bounds[:-1], bounds[1:]):
        if q1 == q0:
            continue
        k0, k1 = int(ip[q0]), int(ip[q1])
        Ab = matrix_from_device_csr((ip[q0:q1 + 1] - k0).contiguous(), cj[k0:k1], cx[k0:k1], q1 - q0, n)
        Cb = Ab.mxm(B, semiring).new()
        nv = Cb.nvals
        if off + nv > out_cols.shape[0]:
            copy_sync()
            raise ValueError(f"output arrays too small: need more than {out_cols.shape[0]} entries")
        matrix_export_host_csr32_async(Cb, out_indptr[q0:q1 + 1], out_cols[off:off + nv], out_vals[off:off + nv])
        t = ctypes.c_int(-1)
        call("GrB_cuda_copy_fence", [ctypes.byref(t)])
        if prev is not None:   # the previous block has (nearly) drained while this one was multiplied: release it, two blocks alive at most
            call("GrB_cuda_copy_wait", [ctypes.c_int(prev[1])])
        prev = (Cb, t.value)
        offs.append((q0, q1, off))
        off += nv
    copy_sync()
    prev = None
    for q0, q1, o in offs:   # block-local row pointers -> global ones
        if o:
            out_indptr[q0:q1] += o
    out_indptr[m] = off
    return off


# ------------------------------------------------------------------ DLPack (any array library that speaks it -- CuPy, JAX, RAPIDS ...)
def vector_to_dlpack(v, sync=True):
    """(values, present) of the vector as DLPack capsules over the library's own device arrays (zero-copy; valid until v changes)"""
    from torch.utils.dlpack import to_dlpack

    vals, present = vector_as_torch(v, sync=sync)
    return to_dlpack(vals), to_dlpack(present)


def vector_from_dlpack(values, present=None, *, name=None):
    """a Vector from DLPack capsules / objects with __dlpack__ (dense values + optional uint8 presence; copied onto the library's arrays)"""
    import torch

    vals = torch.from_dlpack(values)
    pres = None if present is None else torch.from_dlpack(present)
    return vector_from_torch(vals, pres, name=name)


def matrix_to_dlpack(A, sync=True):
    """(indptr int64, col_indices int32, values) of the matrix's device CSR as DLPack capsules (zero-copy, read-only; rows may be
    unsorted after mxm -- matrix_sort(A) first when order matters)"""
    from torch.utils.dlpack import to_dlpack

    return tuple(to_dlpack(t) for t in matrix_as_torch(A, sync=sync))


def matrix_from_dlpack(indptr, col_indices, values, nrows, ncols, *, sorted=True, name=None):
    """a Matrix from DLPack capsules / objects of a device CSR (indptr int64, col_indices int32, values; copied)"""
    import torch

    return matrix_from_device_csr(torch.from_dlpack(indptr), torch.from_dlpack(col_indices), torch.from_dlpack(values), nrows, ncols,
                                  sorted=sorted, name=name)


def _torch_dtype(dt):
    import torch

    return {torch.bool: "BOOL", torch.int8: "INT8", torch.int16: "INT16", torch.int32: "INT32", torch.int64: "INT64",
            torch.uint8: "UINT8", torch.float32: "FP32", torch.float64: "FP64"}[dt]


def _np_to_torch_dtype(np_dtype):
    import torch

    return {"bool": torch.bool, "int8": torch.int8, "int16": torch.int16, "int32": torch.int32, "int64": torch.int64,
            "uint8": torch.uint8, "uint16": torch.int16, "uint32": torch.int32, "uint64": torch.int64,
            "float32": torch.float32, "float64": torch.float64}[np.dtype(np_dtype).name]


def vector_from_torch(values, present=None, *, name=None):
    """dense torch CUDA tensor (+ optional uint8 presence) -> Vector (copied)."""
    from .vector import Vector

    dtype = lookup_dtype(_torch_dtype(values.dtype))
    h = ctypes.c_void_p()
    out = Vector._from_handle(h, dtype, values.numel(), name)
    _wait_for_torch()
    call("GrB_cuda_Vector_import_dense",
         [ctypes.byref(h), dtype, GrB_Index(values.numel()), ctypes.c_void_p(values.data_ptr()),
          None if present is None else ctypes.c_void_p(present.data_ptr()), 1])
    _torch_waits()
    return out


def vector_from_numpy(values, present=None, *, name=None):
    from .vector import Vector

    values = np.ascontiguousarray(values)
    dtype = lookup_dtype(values.dtype)
    h = ctypes.c_void_p()
    out = Vector._from_handle(h, dtype, values.shape[0], name)
    call("GrB_cuda_Vector_import_dense",
         [ctypes.byref(h), dtype, GrB_Index(values.shape[0]), ctypes.c_void_p(values.ctypes.data),
          None if present is None else ctypes.c_void_p(np.ascontiguousarray(present, dtype=np.uint8).ctypes.data), 0])
    return out


def vector_to_numpy(v, values=None, present=None):
    values = np.empty(v.size, dtype=v.dtype.np_type) if values is None else values
    present = np.empty(v.size, dtype=np.uint8) if present is None else present
    call("GrB_cuda_Vector_export_dense", [ctypes.c_void_p(values.ctypes.data), ctypes.c_void_p(present.ctypes.data), v])
    return values, present


def vector_device_pointers(v):
    vals, pres = ctypes.c_void_p(), ctypes.c_void_p()
    call("GrB_cuda_Vector_device_arrays", [v, ctypes.byref(vals), ctypes.byref(pres)])
    return vals.value, pres.value


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}


def vector_as_torch(v, sync=True):
    """zero-copy torch views (values, present) of the vector's device arrays; call vector_touch(v) after writing.
    The library enqueues work on its own stream: unless torch shares that stream (use_torch_stream), the views are
    only safe to read after a synchronisation, which `sync=True` performs."""
    import torch

    if sync:
        globals()["sync"]()
    else:
        _torch_waits()

    pv, pp = vector_device_pointers(v)
    n = v.size
    vals = torch.as_tensor(_CudaArray(pv, (n,), np.dtype(v.dtype.np_type).str if v.dtype.name != "BOOL" else "|u1"), device="cuda")
    pres = torch.as_tensor(_CudaArray(pp, (n,), "|u1"), device="cuda")
    return vals, pres


def matrix_as_torch(A, sync=True):
    """zero-copy torch views (indptr int64, col_indices int32, values) of the matrix's device CSR (GrB_cuda_Matrix_device_csr).
    Rows may be unsorted after mxm -- call matrix_sort(A) first when order matters.  Read-only; valid until A changes."""
    import torch

    ap, aj, ax = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    call("GrB_cuda_Matrix_device_csr", [A, ctypes.byref(ap), ctypes.byref(aj), ctypes.byref(ax)])
    if sync:
        globals()["sync"]()
    else:
        _torch_waits()
    nv = A.nvals
    indptr = torch.as_tensor(_CudaArray(ap.value, (A.nrows + 1,), "<i8"), device="cuda")
    if nv == 0:
        return indptr, torch.empty(0, dtype=torch.int32, device="cuda"), torch.empty(0, dtype=_torch_dtype_of(A.dtype), device="cuda")
    cols = torch.as_tensor(_CudaArray(aj.value, (nv,), "<i4"), device="cuda")
    vals = torch.as_tensor(_CudaArray(ax.value, (nv,), np.dtype(A.dtype.np_type).str if A.dtype.name != "BOOL" else "|u1"), device="cuda")
    return indptr, cols, vals


def _torch_dtype_of(dtype):
    import torch

    return torch.uint8 if dtype.name == "BOOL" else getattr(torch, np.dtype(dtype.np_type).name)


def vector_touch(v):
    call("GrB_cuda_Vector_touch", [v])


def vector_assume_full(v):
    """the arrays were refilled from outside and every position holds an entry: spares the recount (a kernel + a host round trip)"""
    call("GrB_cuda_Vector_assume_full", [v])


def matrix_sort(A):
    call("GrB_cuda_Matrix_sort", [A])


def matrix_compact(A):
    """Squeeze a row-end product into the compact CSR now (export, sort, SpMV and element-wise ops do it on demand)."""
    call("GrB_cuda_Matrix_compact", [A])


def matrix_build_transpose(A):
    call("GrB_cuda_Matrix_build_transpose", [A])


def mxm_symbolic(A, B):
    """(flops, nvals) of A (+).(x) B without forming it."""
    f, n = GrB_Index(), GrB_Index()
    call("GrB_cuda_mxm_symbolic", [ctypes.byref(f), ctypes.byref(n), A, B, None])
    return f.value, n.value


# ------------------------------------------------------------------ fused multiply + exchange over peer memory
def peer_alloc(nbytes):
    """zeroed cudaMalloc buffer that other ranks of the node can map with CUDA IPC; returns the device pointer (int)"""
    p = ctypes.c_void_p()
    call("GrB_cuda_peer_alloc", [ctypes.byref(p), ctypes.c_size_t(int(nbytes))])
    return p.value


def peer_free(ptr):
    call("GrB_cuda_peer_free", [ctypes.c_void_p(ptr)])


def ipc_handle(ptr):
    h = (ctypes.c_ubyte * 64)()
    call("GrB_cuda_ipc_get", [ctypes.c_void_p(ptr), h])
    return bytes(h)


def ipc_open(handle):
    p = ctypes.c_void_p()
    h = (ctypes.c_ubyte * 64).from_buffer_copy(handle)
    call("GrB_cuda_ipc_open", [h, ctypes.byref(p)])
    return p.value


def ipc_close(ptr):
    call("GrB_cuda_ipc_close", [ctypes.c_void_p(ptr)])


def vector_wrap(dtype, n, vals_ptr, present_ptr, *, name=None):
    """a library Vector over caller-owned device arrays (never freed by the library)"""
    from .vector import Vector

    dtype = lookup_dtype(dtype)
    h = ctypes.c_void_p()
    out = Vector._from_handle(h, dtype, n, name)
    call("GrB_cuda_Vector_wrap", [ctypes.byref(h), dtype, GrB_Index(n), ctypes.c_void_p(vals_ptr), ctypes.c_void_p(present_ptr)])
    return out


class peer_targets:
    """with peer_targets(vals_ptrs, present_ptrs, offset, scale=None): w(accum) << A.mxv(x, semiring)
    -- inside the block every finished output position i of a GrB_mxv / GrB_vxm is also stored at offset + i of the target
    vectors (optionally multiplied by scale[i], a device pointer to values of the result type)."""

    def __init__(self, vals_ptrs, present_ptrs, offset, scale=None):
        n = len(vals_ptrs)
        self._v = (ctypes.c_void_p * n)(*vals_ptrs)
        self._p = (ctypes.c_void_p * n)(*present_ptrs) if present_ptrs is not None else None
        self._n, self._off, self._scale = n, int(offset), scale

    def __enter__(self):
        call("GrB_cuda_set_peer_targets", [self._n, self._v, self._p, GrB_Index(self._off), ctypes.c_void_p(self._scale) if self._scale else None])
        return self

    def __exit__(self, *exc):
        call("GrB_cuda_set_peer_targets", [0, None, None, GrB_Index(0), None])
