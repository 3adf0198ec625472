"""CPU-only: the C-ABI library loads and exports every symbol include/grb_cuda.h declares, plus the typed
C-API names and the builtin-object data symbols the reference's registry looks for.  No compute calls."""
import ctypes
import pathlib
import re

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
SO = ROOT / "python-graphblas_b200" / "graphblas_b200" / "libgrb_cuda.so"
TYPES = ["BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP32", "FP64"]


@pytest.fixture(scope="module")
def dll():
    if not SO.exists():
        import __graft_entry__ as g

        g.build()
    return ctypes.CDLL(str(SO))


def test_header_functions_exported(dll):
    text = (ROOT / "include" / "grb_cuda.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(GrB_\w+|GxB_\w+)\s*\(", text)
    names = [n for n in names if not n.endswith("_opaque")]
    assert len(names) > 60
    missing = [n for n in set(names) if not hasattr(dll, n)]
    assert not missing, missing


def test_typed_api_names_exported(dll):
    stems = ["GrB_Matrix_import", "GrB_Matrix_export", "GrB_Matrix_build", "GrB_Matrix_extractTuples",
             "GrB_Matrix_extractElement", "GrB_Vector_build", "GrB_Vector_extractTuples", "GrB_Vector_setElement",
             "GrB_Vector_extractElement", "GrB_Vector_reduce", "GrB_Vector_assign"]
    missing = [f"{s}_{t}" for s in stems for t in TYPES if not hasattr(dll, f"{s}_{t}")]
    assert not missing, missing


def test_builtin_objects_are_data_symbols(dll):
    wanted = ["GrB_BOOL", "GrB_FP64", "GrB_ALL", "GrB_DESC_RSC", "GrB_DESC_T0", "GrB_DESC_T1", "GrB_DESC_ST0",
              "GrB_LOR_LAND_SEMIRING_BOOL", "GrB_PLUS_MONOID_INT64", "GrB_MIN_INT64", "GrB_LNOT", "GrB_ONEB_INT64"]
    for t in TYPES[1:]:
        wanted += [f"GrB_PLUS_TIMES_SEMIRING_{t}", f"GrB_MIN_PLUS_SEMIRING_{t}", f"GxB_ANY_PAIR_{t}",
                   f"GxB_PLUS_SECOND_{t}", f"GrB_PLUS_{t}", f"GrB_MIN_{t}"]
    for name in wanted:
        ctypes.c_void_p.in_dll(dll, name)  # raises ValueError if the symbol is missing
    dll.GrB_cuda_lookup.restype = ctypes.c_void_p
    assert dll.GrB_cuda_lookup(b"GrB_MIN_PLUS_SEMIRING_INT32") == ctypes.c_void_p.in_dll(dll, "GrB_MIN_PLUS_SEMIRING_INT32").value
    assert dll.GrB_cuda_lookup(b"no_such_symbol") is None


def test_host_registry_finds_hot_semirings():
    import graphblas_b200 as gb
    from graphblas_b200 import operator

    operator.initialize()
    # reference tests/test_op.py:97-98: min_plus[INT32].gb_obj is lib.GrB_MIN_PLUS_SEMIRING_INT32
    assert gb.semiring.min_plus["INT32"].gb_name == "GrB_MIN_PLUS_SEMIRING_INT32"
    assert gb.semiring.plus_times["FP32"].gb_name == "GrB_PLUS_TIMES_SEMIRING_FP32"
    assert gb.semiring.any_pair["INT64"].gb_name == "GxB_ANY_PAIR_INT64"
    assert gb.semiring.plus_second["FP64"].gb_name == "GxB_PLUS_SECOND_FP64"
    assert gb.semiring.lor_land["BOOL"].gb_name == "GrB_LOR_LAND_SEMIRING_BOOL"
    # pair forces INT64 (reference operator/binary.py:387-388); lor_land on ints runs in BOOL (semiring.py:538-547)
    assert operator.get_typed_op(gb.semiring.any_pair, gb.dtypes.BOOL, gb.dtypes.FP32, kind="semiring").type == gb.dtypes.INT64
    assert gb.semiring.lor_land["INT64"].gb_name == "GrB_LOR_LAND_SEMIRING_BOOL"
    # monoid / binaryop decomposition (reference tests/test_op.py:909-912)
    sr = gb.semiring.min_plus["INT64"]
    assert sr.monoid.gb_name == "GrB_MIN_MONOID_INT64" and sr.binaryop.gb_name == "GrB_PLUS_INT64"


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import graphblas_b200 as gb

    with pytest.raises(gb.exceptions.Panic, match="no CPU fallback"):
        gb.init()
