"""The cffi-shaped (ffi, lib) adapter a python-graphblas maintainer binds (graphblas_b200/ffi.py), driven the way the
reference drives its binding: handles in 1-element cells from ``ffi.new("GrB_X*")`` passed to ``lib.GrB_X_new`` and
dereferenced with ``[0]`` (reference graphblas/core/matrix.py:190-203, core/scalar.py:76-84, core/base.py:23-54),
numpy buffers through ``ffi.from_buffer`` (core/utils.py:315-328), error strings through ``ffi.new("char**")`` /
``ffi.string`` (exceptions.py:171-189), operators found by scanning ``dir(lib)`` (core/operator/base.py:690, 803-893).

CPU part: signatures parsed from include/grb_cuda.h, symbol discovery, the host-only GrB_Scalar family.  GPU part
(-m gpu): one mxm / reduce / select through nothing but (ffi, lib), checked against the oracle."""
import ctypes
import re

import numpy as np
import pytest

from graphblas_b200 import ffi as F

TYPES = ["BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP32", "FP64"]


@pytest.fixture(scope="module")
def lib():
    return F.__getattr__("lib")


def test_every_header_function_resolves_with_a_signature(lib):
    sigs = F.parse_header()
    assert len(sigs) > 300
    for name, (restype, argtypes) in sigs.items():
        fn = getattr(lib, name)
        assert fn.argtypes == argtypes and fn.restype is restype, name
    # spot checks of the mapping cffi would derive from the same declarations
    assert sigs["GrB_Matrix_new"] == (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64])
    assert sigs["GrB_Vector_setElement_FP32"][1][1] is ctypes.c_float
    assert sigs["GrB_Scalar_extractElement_INT16"][1] == [ctypes.c_void_p, ctypes.c_void_p]
    assert sigs["GrB_cuda_last_error"] == (ctypes.c_char_p, [])


def test_names_the_reference_formats_at_run_time_exist(lib):
    """lifecycle / data / typed names; reference call sites in the comments"""
    plain = ["GrB_mxm", "GrB_mxv", "GrB_vxm", "GrB_transpose",
             "GrB_Matrix_new", "GrB_Matrix_free", "GrB_Matrix_dup", "GrB_Matrix_clear", "GrB_Matrix_nrows", "GrB_Matrix_ncols",
             "GrB_Matrix_nvals", "GrB_Matrix_wait", "GrB_Matrix_error", "GrB_Matrix_exportSize",
             "GrB_Vector_new", "GrB_Vector_free", "GrB_Vector_dup", "GrB_Vector_clear", "GrB_Vector_size", "GrB_Vector_nvals",
             "GrB_Vector_wait", "GrB_Vector_error",
             "GrB_Scalar_new", "GrB_Scalar_free", "GrB_Scalar_dup", "GrB_Scalar_clear", "GrB_Scalar_nvals", "GrB_Scalar_wait",   # core/scalar.py:83, 235-280
             "GrB_Vector_reduce_Monoid_Scalar", "GrB_Matrix_reduce_Monoid_Scalar",                                               # core/vector.py:1670, core/matrix.py:2750
             "GrB_Vector_apply_BinaryOp1st_Scalar", "GrB_Vector_apply_BinaryOp2nd_Scalar", "GrB_Matrix_apply_BinaryOp1st_Scalar",
             "GrB_Matrix_apply_BinaryOp2nd_Scalar", "GrB_Vector_select_Scalar", "GrB_Matrix_select_Scalar",
             "GrB_Vector_assign", "GrB_Vector_assign_Scalar", "GrB_Matrix_assign",                                                # core/vector.py:1928
             "GrB_Vector_eWiseAdd_BinaryOp", "GrB_Vector_eWiseMult_BinaryOp", "GrB_Matrix_eWiseAdd_BinaryOp",
             "GrB_Matrix_eWiseMult_BinaryOp", "GrB_Vector_apply", "GrB_Matrix_apply"]
    stems = ["GrB_Matrix_import", "GrB_Matrix_export", "GrB_Matrix_build", "GrB_Matrix_extractTuples", "GrB_Matrix_extractElement",
             "GrB_Matrix_setElement", "GrB_Vector_build", "GrB_Vector_extractTuples", "GrB_Vector_setElement", "GrB_Vector_extractElement",
             "GrB_Vector_reduce", "GrB_Matrix_reduce", "GrB_Vector_assign", "GrB_Scalar_setElement", "GrB_Scalar_extractElement",
             "GrB_Vector_apply_BinaryOp1st", "GrB_Vector_apply_BinaryOp2nd", "GrB_Matrix_apply_BinaryOp1st", "GrB_Matrix_apply_BinaryOp2nd",  # core/vector.py:1477, core/matrix.py:2472
             "GrB_Vector_select", "GrB_Matrix_select"]                                                                            # core/vector.py:1622, core/matrix.py:2621
    names = plain + [f"{s}_{t}" for s in stems for t in TYPES]
    missing = [n for n in names if not callable(getattr(lib, n, None))]
    assert not missing, missing


def test_dir_lib_feeds_the_operator_registry(lib):
    """the registry regex-scans dir(lib) (reference core/operator/base.py:690); every family it looks for must be enumerable"""
    names = [n for n in dir(lib) if n[0] != "_"]
    T = "|".join(TYPES)
    families = {
        "semiring GrB": rf"^GrB_(PLUS|MIN|MAX)_(TIMES|PLUS|MIN|MAX|FIRST|SECOND)_SEMIRING_({T})$",
        "semiring GxB": rf"^GxB_(PLUS|TIMES|MIN|MAX|ANY)_(FIRST|SECOND|PAIR|MIN|MAX|PLUS|MINUS|RMINUS|TIMES|DIV|RDIV)_({T})$",
        "monoid": rf"^GrB_(PLUS|TIMES|MIN|MAX)_MONOID_({T})$",
        "binary": rf"^GrB_(FIRST|SECOND|MIN|MAX|PLUS|MINUS|TIMES|DIV|EQ|NE|GT|LT|GE|LE|ONEB)_({T})$",
        "unary": rf"^GrB_(IDENTITY|AINV|MINV|ABS)_({T})$",
        "select positional": r"^GrB_(TRIL|TRIU|DIAG|OFFDIAG|COLLE|COLGT|ROWLE|ROWGT)$",
        "select value": rf"^GrB_VALUE(EQ|NE|GT|GE|LT|LE)_({T})$",
        "descriptor": r"^GrB_DESC_(R?S?C?(T0)?(T1)?)$",
    }
    counts = {k: sum(1 for n in names if re.match(rx, n)) for k, rx in families.items()}
    assert counts["semiring GrB"] == 12 * 10 and counts["monoid"] == 40 and counts["unary"] == 44
    assert counts["binary"] == 15 * 11 and counts["select positional"] == 8 and counts["select value"] == 66
    assert counts["descriptor"] == 31 and counts["semiring GxB"] > 400
    # the names the reference hard-references while importing (last bullet)
    for n in ["GrB_ALL", "GrB_MATERIALIZE", "GrB_COMPLETE", "GrB_CSR_FORMAT", "GrB_CSC_FORMAT", "GrB_COO_FORMAT", "GrB_INDEX_MAX",
              "GrB_BLOCKING", "GrB_NONBLOCKING", "GxB_ANY_PAIR_INT64", "GrB_VALUENE_INT64", "GxB_ONE_INT64", "GrB_LAND", "GrB_LOR",
              "GrB_ONEB_INT64", "GrB_MIN_PLUS_SEMIRING_INT32"]:
        assert n in names, n
        assert getattr(lib, n) is not None
    # handles are cached objects, so identity checks like reference tests/test_op.py:97-98 hold
    assert lib.GrB_MIN_PLUS_SEMIRING_INT32 is lib.GrB_MIN_PLUS_SEMIRING_INT32


def test_scalar_family_through_ffi(lib):
    """GrB_Scalar lives on the host, so its whole life cycle runs without a GPU (reference core/scalar.py:76-84, 199-280)"""
    ffi = F.ffi
    s = ffi.new("GrB_Scalar*")
    assert lib.GrB_Scalar_new(s, lib.GrB_INT64) == lib.GrB_SUCCESS and s[0]
    n = ffi.new("GrB_Index*")
    assert lib.GrB_Scalar_nvals(n, s[0]) == 0 and n[0] == 0
    x = ffi.new("int64_t*")
    assert lib.GrB_Scalar_extractElement_INT64(x, s[0]) == lib.GrB_NO_VALUE
    assert lib.GrB_Scalar_setElement_INT64(s[0], -(2**40) - 5) == 0
    assert lib.GrB_Scalar_extractElement_INT64(x, s[0]) == 0 and x[0] == -(2**40) - 5
    d = ffi.new("double*")
    assert lib.GrB_Scalar_extractElement_FP64(d, s[0]) == 0 and d[0] == float(-(2**40) - 5)   # typecast on the way out
    b = ffi.new("bool*")
    assert lib.GrB_Scalar_extractElement_BOOL(b, s[0]) == 0 and b[0] is True
    assert lib.GrB_Scalar_setElement_FP32(s[0], 2.75) == 0                                       # ... and on the way in (C cast)
    assert lib.GrB_Scalar_extractElement_INT64(x, s[0]) == 0 and x[0] == 2
    t = ffi.new("GrB_Scalar*")
    assert lib.GrB_Scalar_dup(t, s[0]) == 0
    assert lib.GrB_Scalar_clear(s[0]) == 0 and lib.GrB_Scalar_nvals(n, s[0]) == 0 and n[0] == 0
    assert lib.GrB_Scalar_nvals(n, t[0]) == 0 and n[0] == 1
    assert lib.GrB_Scalar_extractElement_INT64(x, t[0]) == 0 and x[0] == 2
    msg = ffi.new("char**")
    assert lib.GrB_Scalar_error(msg, t[0]) == 0 and ffi.string(msg[0]) == b""
    assert lib.GrB_Scalar_free(s) == 0 and lib.GrB_Scalar_free(t) == 0 and not s[0] and not t[0]
    assert lib.GrB_Scalar_nvals(n, ffi.NULL) == lib.GrB_UNINITIALIZED_OBJECT


def test_ffi_spellings():
    ffi = F.ffi
    a = np.arange(5, dtype=np.uint64)
    p = ffi.cast("GrB_Index*", ffi.from_buffer(a))
    assert p.value == a.ctypes.data
    arr = ffi.new("GrB_Index[]", 4)
    assert len(arr) == 4 and ffi.sizeof("GrB_Index") == 8
    assert ffi.string(ffi.new("char[]", b"abc")) == b"abc"
    cell = ffi.new("GrB_Vector*")
    as_matrix = ffi.cast("GrB_Matrix*", cell)      # reference core/vector.py:203: the same storage seen as another handle type
    cell[0] = 1234
    assert as_matrix[0] == 1234
    with pytest.raises(TypeError):
        ffi.new("no_such_type*")


def test_host_scalar_object_uses_grb_scalar():
    """graphblas_b200.Scalar is a GrB_Scalar by default, like the reference's (core/scalar.py:50-84)"""
    import graphblas_b200 as gb

    s = gb.Scalar(gb.dtypes.FP64)
    assert s.is_grbscalar and s.is_empty and s.value is None and s.nvals == 0
    s.value = 3.5
    assert s.value == 3.5 and s.nvals == 1 and not s.is_empty and s == 3.5
    t = s.dup(gb.dtypes.INT32)
    assert t.value == 3 and t.dtype == gb.dtypes.INT32
    s.clear()
    assert s.is_empty and t.value == 3
    c = gb.Scalar.from_value(7, is_cscalar=True)
    assert c.is_cscalar and c.value == 7 and c.dtype == gb.dtypes.INT64


# ------------------------------------------------------------------ GPU: the binding executes
@pytest.mark.gpu
def test_mxm_reduce_select_through_ffi_only(lib):
    from oracle import semantics as S

    ffi = F.ffi
    assert lib.GrB_init(lib.GrB_NONBLOCKING) == 0
    rng = np.random.default_rng(5)
    n = 40
    key = np.unique(rng.integers(0, n * n, 300))
    I, J = (key // n).astype(np.uint64), (key % n).astype(np.uint64)
    X = rng.integers(-4, 5, key.size).astype(np.float64)
    A = ffi.new("GrB_Matrix*")
    assert lib.GrB_Matrix_new(A, lib.GrB_FP64, n, n) == 0
    assert lib.GrB_Matrix_build_FP64(A[0], ffi.from_buffer(I), ffi.from_buffer(J), ffi.from_buffer(X), key.size, ffi.NULL) == 0
    C = ffi.new("GrB_Matrix*")
    assert lib.GrB_Matrix_new(C, lib.GrB_FP64, n, n) == 0
    # reference core/base.py:496-503: [C, mask, accum, op, A, B, desc]
    assert lib.GrB_mxm(C[0], ffi.NULL, ffi.NULL, lib.GrB_PLUS_TIMES_SEMIRING_FP64, A[0], A[0], lib.GrB_DESC_T1) == 0
    nv = ffi.new("GrB_Index*")
    assert lib.GrB_Matrix_nvals(nv, C[0]) == 0
    oi, oj, ox = np.empty(nv[0], np.uint64), np.empty(nv[0], np.uint64), np.empty(nv[0], np.float64)
    assert lib.GrB_Matrix_extractTuples_FP64(ffi.from_buffer(oi), ffi.from_buffer(oj), ffi.from_buffer(ox), nv, C[0]) == 0
    Ao = S.SpMat.from_coo(I.astype(np.int64), J.astype(np.int64), X, n, n)
    want = S.mxm(S.SpMat(n, n, np.float64), None, None, "plus_times", Ao, Ao, t1=True)
    wi, wj, wx = want.to_coo()
    order = np.lexsort((oj, oi))
    assert np.array_equal(oi[order].astype(np.int64), wi) and np.array_equal(oj[order].astype(np.int64), wj)
    assert np.array_equal(ox[order], wx)
    # reduce into a GrB_Scalar (reference core/matrix.py:2750), then with an accumulator
    s = ffi.new("GrB_Scalar*")
    assert lib.GrB_Scalar_new(s, lib.GrB_FP64) == 0
    assert lib.GrB_Matrix_reduce_Monoid_Scalar(s[0], ffi.NULL, lib.GrB_PLUS_MONOID_FP64, A[0], ffi.NULL) == 0
    d = ffi.new("double*")
    assert lib.GrB_Scalar_extractElement_FP64(d, s[0]) == 0 and d[0] == X.sum()
    assert lib.GrB_Matrix_reduce_Monoid_Scalar(s[0], lib.GrB_PLUS_FP64, lib.GrB_PLUS_MONOID_FP64, A[0], ffi.NULL) == 0
    assert lib.GrB_Scalar_extractElement_FP64(d, s[0]) == 0 and d[0] == 2 * X.sum()
    # select tril through the typed name (reference core/matrix.py:2621)
    L = ffi.new("GrB_Matrix*")
    assert lib.GrB_Matrix_new(L, lib.GrB_FP64, n, n) == 0
    assert lib.GrB_Matrix_select_INT64(L[0], ffi.NULL, ffi.NULL, lib.GrB_TRIL, A[0], -1, ffi.NULL) == 0
    assert lib.GrB_Matrix_nvals(nv, L[0]) == 0 and nv[0] == int(np.sum(J.astype(np.int64) <= I.astype(np.int64) - 1))
    # a dimension error comes back synchronously with a message (reference exceptions.py:171-189)
    B = ffi.new("GrB_Matrix*")
    assert lib.GrB_Matrix_new(B, lib.GrB_FP64, n + 1, n) == 0
    assert lib.GrB_mxm(C[0], ffi.NULL, ffi.NULL, lib.GrB_PLUS_TIMES_SEMIRING_FP64, A[0], B[0], ffi.NULL) == lib.GrB_DIMENSION_MISMATCH
    msg = ffi.new("char**")
    assert lib.GrB_Matrix_error(msg, C[0]) == 0 and b"GrB_mxm" in ffi.string(msg[0])
    for h in (A, B, C, L):
        assert lib.GrB_Matrix_free(h) == 0
    assert lib.GrB_Scalar_free(s) == 0
