"""Converters around the hot path: scipy.sparse and Matrix Market, host side only.

Mirrors reference graphblas/io/_scipy.py:8-119 (``from_scipy_sparse`` / ``to_scipy_sparse``) and io/_matrixmarket.py:8-140
(``mmread`` / ``mmwrite`` through ``scipy.io``): CSR / CSC inputs go through ``GrB_Matrix_import_<T>`` unchanged, everything
else through ``build`` with an optional ``dup_op``.  Values stay typed: the reference's iso-valued import (one stored value
for a structure-only graph, io/_scipy.py:31-46) is a SuiteSparse extension; here structure-only semirings (any_pair,
plus_second, ...) already skip the value arrays inside the kernels."""
import numpy as np

from .dtypes import lookup_dtype
from .matrix import Matrix
from .vector import Vector


def from_scipy_sparse(A, *, dup_op=None, name=None):
    nrows, ncols = A.shape
    dtype = lookup_dtype(A.dtype)
    if A.nnz == 0:
        return Matrix(dtype, nrows=nrows, ncols=ncols, name=name)
    if A.format == "csr":
        return Matrix.from_csr(A.indptr, A.indices, A.data, ncols=ncols, name=name)
    if A.format == "csc":
        return Matrix.from_csc(A.indptr, A.indices, A.data, nrows=nrows, name=name)
    if A.format != "coo":
        A = A.tocoo()
    return Matrix.from_coo(A.row, A.col, A.data, nrows=nrows, ncols=ncols, dtype=dtype, dup_op=dup_op, name=name)


def to_scipy_sparse(A, format="csr"):
    """reference io/_scipy.py:68-119; a Vector becomes an n x 1 array"""
    import scipy.sparse as ss

    format = format.lower()
    if isinstance(A, Vector):
        idx, vals = A.to_coo()
        rv = ss.coo_array((vals, (idx.astype(np.int64), np.zeros(idx.size, dtype=np.int64))), shape=(A.size, 1))
        return rv.asformat(format)
    if format == "csc":
        indptr, rows, vals = A.to_csc()
        return ss.csc_array((vals, rows.astype(np.int64), indptr.astype(np.int64)), shape=(A.nrows, A.ncols))
    indptr, cols, vals = A.to_csr()
    rv = ss.csr_array((vals, cols.astype(np.int64), indptr.astype(np.int64)), shape=(A.nrows, A.ncols))
    return rv if format == "csr" else rv.asformat(format)


def mmread(source, *, dup_op=None, name=None, **kwargs):
    """reference io/_matrixmarket.py:8-97 (engine "scipy")"""
    from scipy.io import mmread as _mmread

    array = _mmread(source, **kwargs)
    if hasattr(array, "format"):
        return from_scipy_sparse(array, dup_op=dup_op, name=name)
    array = np.asarray(array)   # dense Matrix Market file
    r, c = np.nonzero(np.ones_like(array, dtype=bool))
    return Matrix.from_coo(r, c, array[r, c], nrows=array.shape[0], ncols=array.shape[1], name=name)


def mmwrite(target, matrix, *, comment="", field=None, precision=None, symmetry=None, **kwargs):
    """reference io/_matrixmarket.py:100-140"""
    from scipy.io import mmwrite as _mmwrite

    _mmwrite(target, to_scipy_sparse(matrix, "coo"), comment=comment, field=field, precision=precision, symmetry=symmetry, **kwargs)
