"""Builtin data types; mirrors reference graphblas/core/dtypes.py:329-420 (DataType), :527-549 (lookup_dtype),
:552-568 (unify = numpy promote_types)."""
import ctypes

import numpy as np

from ._lib import lib


class DataType:
    __slots__ = ("name", "gb_name", "c_type", "np_type", "ctype", "_carg_cache")

    def __init__(self, name, gb_name, c_type, np_type, ctype):
        self.name, self.gb_name, self.c_type, self.np_type, self.ctype = name, gb_name, c_type, np.dtype(np_type), ctype
        self._carg_cache = None

    @property
    def gb_obj(self):
        if self._carg_cache is None:
            self._carg_cache = getattr(lib(), self.gb_name)
        return self._carg_cache

    _carg = gb_obj

    def __repr__(self):
        return self.name

    def __eq__(self, other):
        try:
            return lookup_dtype(other) is self
        except Exception:
            return False

    def __hash__(self):
        return hash(self.name)


BOOL = DataType("BOOL", "GrB_BOOL", "bool", np.bool_, ctypes.c_bool)
INT8 = DataType("INT8", "GrB_INT8", "int8_t", np.int8, ctypes.c_int8)
INT16 = DataType("INT16", "GrB_INT16", "int16_t", np.int16, ctypes.c_int16)
INT32 = DataType("INT32", "GrB_INT32", "int32_t", np.int32, ctypes.c_int32)
INT64 = DataType("INT64", "GrB_INT64", "int64_t", np.int64, ctypes.c_int64)
UINT8 = DataType("UINT8", "GrB_UINT8", "uint8_t", np.uint8, ctypes.c_uint8)
UINT16 = DataType("UINT16", "GrB_UINT16", "uint16_t", np.uint16, ctypes.c_uint16)
UINT32 = DataType("UINT32", "GrB_UINT32", "uint32_t", np.uint32, ctypes.c_uint32)
UINT64 = DataType("UINT64", "GrB_UINT64", "uint64_t", np.uint64, ctypes.c_uint64)
FP32 = DataType("FP32", "GrB_FP32", "float", np.float32, ctypes.c_float)
FP64 = DataType("FP64", "GrB_FP64", "double", np.float64, ctypes.c_double)

_ALL = [BOOL, INT8, INT16, INT32, INT64, UINT8, UINT16, UINT32, UINT64, FP32, FP64]
_registry = {}
for _dt in _ALL:
    _registry[_dt.name] = _registry[_dt.name.lower()] = _registry[_dt.gb_name] = _dt
    _registry[_dt.np_type] = _registry[_dt.np_type.name] = _dt
_registry[bool] = _registry["bool"] = BOOL
_registry[int] = _registry["int"] = INT64
_registry[float] = _registry["float"] = FP64
_INDEX = UINT64


def lookup_dtype(key):
    if isinstance(key, DataType):
        return key
    try:
        return _registry[key]
    except (KeyError, TypeError):
        pass
    try:
        return _registry[np.dtype(key)]
    except Exception:
        raise ValueError(f"Unknown dtype: {key} of type {type(key)}") from None


def unify(type1, type2):
    if type1 is type2:
        return type1
    return lookup_dtype(np.promote_types(type1.np_type, type2.np_type))
