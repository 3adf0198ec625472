"""GrB_Info -> exception map; same names and behaviour as reference graphblas/exceptions.py:123-189."""
import ctypes

from ._lib import CONSTANTS, lib


class GraphblasException(Exception):
    pass


class NoValue(GraphblasException):
    pass


class UninitializedObject(GraphblasException):
    pass


class InvalidObject(GraphblasException):
    pass


class NullPointer(GraphblasException):
    pass


class InvalidValue(GraphblasException):
    pass


class InvalidIndex(GraphblasException):
    pass


class DomainMismatch(GraphblasException):
    pass


class DimensionMismatch(GraphblasException):
    pass


class OutputNotEmpty(GraphblasException):
    pass


class EmptyObject(GraphblasException):
    pass


class OutOfMemory(GraphblasException):
    pass


class InsufficientSpace(GraphblasException):
    pass


class IndexOutOfBound(GraphblasException):
    pass


class Panic(GraphblasException):
    pass


class NotImplementedException(GraphblasException):
    pass


_error_code_lookup = {
    CONSTANTS["GrB_UNINITIALIZED_OBJECT"]: UninitializedObject,
    CONSTANTS["GrB_INVALID_OBJECT"]: InvalidObject,
    CONSTANTS["GrB_NULL_POINTER"]: NullPointer,
    CONSTANTS["GrB_INVALID_VALUE"]: InvalidValue,
    CONSTANTS["GrB_INVALID_INDEX"]: InvalidIndex,
    CONSTANTS["GrB_DOMAIN_MISMATCH"]: DomainMismatch,
    CONSTANTS["GrB_DIMENSION_MISMATCH"]: DimensionMismatch,
    CONSTANTS["GrB_OUTPUT_NOT_EMPTY"]: OutputNotEmpty,
    CONSTANTS["GrB_EMPTY_OBJECT"]: EmptyObject,
    CONSTANTS["GrB_OUT_OF_MEMORY"]: OutOfMemory,
    CONSTANTS["GrB_INSUFFICIENT_SPACE"]: InsufficientSpace,
    CONSTANTS["GrB_INDEX_OUT_OF_BOUNDS"]: IndexOutOfBound,
    CONSTANTS["GrB_PANIC"]: Panic,
    CONSTANTS["GrB_NOT_IMPLEMENTED"]: NotImplementedException,
}


def check_status(response_code, args):
    """reference graphblas/exceptions.py:153-189: 0 ok, GrB_NO_VALUE returned, else fetch the text and raise."""
    if response_code == 0:
        return None
    if response_code == 1:
        return NoValue
    arg = args[0] if isinstance(args, (list, tuple)) and args else args
    text = None
    type_name = type(arg).__name__
    carg = getattr(arg, "_carg", None)
    if carg is not None and type_name in ("Matrix", "Vector"):
        msg = ctypes.c_char_p()
        getattr(lib(), f"GrB_{type_name}_error")(ctypes.byref(msg), carg)
        text = msg.value.decode() if msg.value else None
    if not text:
        text = (lib().GrB_cuda_last_error() or b"").decode()
    raise _error_code_lookup.get(response_code, GraphblasException)(text)
