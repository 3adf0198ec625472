#!/usr/bin/env python
"""Generates builtins_gen.inc: every builtin GraphBLAS object this library exports as a C data
symbol (type, unary/binary op, monoid, semiring, descriptor), plus the name->handle table.

The names follow the GraphBLAS C API 2.0 / SuiteSparse GxB conventions because the reference's
operator registry finds operators by regex over dir(lib)
(reference graphblas/core/operator/semiring.py:185-219, monoid.py:239-255, binary.py _parse_config).
Run:  python gen_builtins.py > builtins_gen.inc
"""
TYPES = ["BOOL", "INT8", "INT16", "INT32", "INT64", "UINT8", "UINT16", "UINT32", "UINT64", "FP32", "FP64"]
NUM = TYPES[1:]
INTS = TYPES[1:9]
out = []
table = []


def emit(kind, sym, init):
    out.append(f"static {kind}_opaque obj_{sym} = {init};")
    out.append(f'extern "C" {{ {kind} {sym} = &obj_{sym}; }}')
    table.append(sym)


for t in TYPES:
    pass  # types are defined by hand in builtins.cu (GrB_BOOL ...)

# ---- unary
for t in TYPES:
    emit("GrB_UnaryOp", f"GrB_IDENTITY_{t}", f'{{UOP_IDENTITY, TC_{t}, "GrB_IDENTITY_{t}"}}')
    emit("GrB_UnaryOp", f"GrB_AINV_{t}", f'{{UOP_AINV, TC_{t}, "GrB_AINV_{t}"}}')
    emit("GrB_UnaryOp", f"GrB_MINV_{t}", f'{{UOP_MINV, TC_{t}, "GrB_MINV_{t}"}}')
    emit("GrB_UnaryOp", f"GrB_ABS_{t}", f'{{UOP_ABS, TC_{t}, "GrB_ABS_{t}"}}')
    emit("GrB_UnaryOp", f"GxB_ONE_{t}", f'{{UOP_ONE, TC_{t}, "GxB_ONE_{t}"}}')
    emit("GrB_UnaryOp", f"GxB_LNOT_{t}", f'{{UOP_LNOT, TC_{t}, "GxB_LNOT_{t}"}}')
emit("GrB_UnaryOp", "GrB_LNOT", '{UOP_LNOT, TC_BOOL, "GrB_LNOT"}')
for t in INTS:
    emit("GrB_UnaryOp", f"GrB_BNOT_{t}", f'{{UOP_BNOT, TC_{t}, "GrB_BNOT_{t}"}}')
for t in ["FP32", "FP64"]:   # the floating-point GxB unary ops the aggregator finalizers and PageRank-style recipes use
    for op in ["SQRT", "EXP", "LOG", "EXP2", "LOG2", "LOG10", "FLOOR", "CEIL", "ROUND", "TRUNC", "SIGNUM"]:
        emit("GrB_UnaryOp", f"GxB_{op}_{t}", f'{{UOP_{op}, TC_{t}, "GxB_{op}_{t}"}}')

# ---- binary
GRB_BIN = ["FIRST", "SECOND", "MIN", "MAX", "PLUS", "MINUS", "TIMES", "DIV"]
GXB_BIN = ["RMINUS", "RDIV", "PAIR", "ANY", "LOR", "LAND", "LXOR", "ISEQ", "ISNE", "POW"]
CMP = ["EQ", "NE", "GT", "LT", "GE", "LE"]
IS_CMP = {"ISGT": "GT", "ISLT": "LT", "ISGE": "GE", "ISLE": "LE"}
for t in TYPES:
    for op in GRB_BIN:
        emit("GrB_BinaryOp", f"GrB_{op}_{t}", f'{{OP_{op}, TC_{t}, TC_{t}, "GrB_{op}_{t}"}}')
    emit("GrB_BinaryOp", f"GrB_ONEB_{t}", f'{{OP_PAIR, TC_{t}, TC_{t}, "GrB_ONEB_{t}"}}')
    for op in GXB_BIN:
        emit("GrB_BinaryOp", f"GxB_{op}_{t}", f'{{OP_{op}, TC_{t}, TC_{t}, "GxB_{op}_{t}"}}')
    for op, code in IS_CMP.items():   # GxB_ISGT ...: the comparison as 1 / 0 in the operand type (same device code as GT ..., other result type)
        emit("GrB_BinaryOp", f"GxB_{op}_{t}", f'{{OP_{code}, TC_{t}, TC_{t}, "GxB_{op}_{t}"}}')
    for op in CMP:
        emit("GrB_BinaryOp", f"GrB_{op}_{t}", f'{{OP_{op}, TC_{t}, TC_BOOL, "GrB_{op}_{t}"}}')
for op in ["LOR", "LAND", "LXOR", "LXNOR"]:
    emit("GrB_BinaryOp", f"GrB_{op}", f'{{OP_{op}, TC_BOOL, TC_BOOL, "GrB_{op}"}}')

# ---- monoids
for t in NUM:
    for op in ["PLUS", "TIMES", "MIN", "MAX"]:
        emit("GrB_Monoid", f"GrB_{op}_MONOID_{t}", f'{{OP_{op}, TC_{t}, "GrB_{op}_MONOID_{t}"}}')
for t in TYPES:
    emit("GrB_Monoid", f"GxB_ANY_{t}_MONOID", f'{{OP_ANY, TC_{t}, "GxB_ANY_{t}_MONOID"}}')
for op in ["LOR", "LAND", "LXOR", "LXNOR"]:
    emit("GrB_Monoid", f"GrB_{op}_MONOID_BOOL", f'{{OP_{op}, TC_BOOL, "GrB_{op}_MONOID_BOOL"}}')
emit("GrB_Monoid", "GxB_EQ_BOOL_MONOID", '{OP_LXNOR, TC_BOOL, "GxB_EQ_BOOL_MONOID"}')

# ---- semirings
GRB_SR = {("PLUS", "TIMES"), ("PLUS", "MIN"), ("MIN", "PLUS"), ("MIN", "TIMES"), ("MIN", "FIRST"), ("MIN", "SECOND"),
          ("MIN", "MAX"), ("MAX", "PLUS"), ("MAX", "TIMES"), ("MAX", "FIRST"), ("MAX", "SECOND"), ("MAX", "MIN")}
ADDS = ["PLUS", "TIMES", "MIN", "MAX", "ANY"]
MULS = ["FIRST", "SECOND", "PAIR", "MIN", "MAX", "PLUS", "MINUS", "RMINUS", "TIMES", "DIV", "RDIV", "LOR", "LAND", "LXOR", "ISEQ", "ISNE", "POW"]
for t in NUM:
    for a in ADDS:
        for m in MULS:
            if (a, m) in GRB_SR:
                emit("GrB_Semiring", f"GrB_{a}_{m}_SEMIRING_{t}", f'{{OP_{a}, OP_{m}, TC_{t}, "GrB_{a}_{m}_SEMIRING_{t}"}}')
            else:
                emit("GrB_Semiring", f"GxB_{a}_{m}_{t}", f'{{OP_{a}, OP_{m}, TC_{t}, "GxB_{a}_{m}_{t}"}}')
        for m, code in IS_CMP.items():
            emit("GrB_Semiring", f"GxB_{a}_{m}_{t}", f'{{OP_{a}, OP_{code}, TC_{t}, "GxB_{a}_{m}_{t}"}}')
# comparison multiplies under a logical monoid (GxB_LOR_GT_INT32 ...: T x T -> BOOL, monoid over BOOL).  The kernels are typed by ONE
# type, so these run in T: the comparison yields 1 / 0 in T, LOR of booleans is MAX of those, LAND is MIN, ANY is ANY; the write-back
# casts the 1 / 0 result into the output type (BOOL for `.new()`).  LXOR / EQ monoids would need the parity of a sum: not offered.
for t in NUM:
    for a, acode in (("LOR", "OP_MAX"), ("LAND", "OP_MIN"), ("ANY", "OP_ANY")):
        for m in CMP:
            emit("GrB_Semiring", f"GxB_{a}_{m}_{t}", f'{{{acode}, OP_{m}, TC_{t}, "GxB_{a}_{m}_{t}"}}')
BADDS = {"LOR": "OP_LOR", "LAND": "OP_LAND", "LXOR": "OP_LXOR", "EQ": "OP_LXNOR", "ANY": "OP_ANY"}
BMULS = ["FIRST", "SECOND", "PAIR", "LOR", "LAND", "LXOR", "EQ", "NE", "GT", "LT", "GE", "LE"]
GRB_BSR = {("LOR", "LAND"): "GrB_LOR_LAND_SEMIRING_BOOL", ("LAND", "LOR"): "GrB_LAND_LOR_SEMIRING_BOOL",
           ("LXOR", "LAND"): "GrB_LXOR_LAND_SEMIRING_BOOL", ("EQ", "LOR"): "GrB_LXNOR_LOR_SEMIRING_BOOL"}
for a, acode in BADDS.items():
    for m in BMULS:
        name = GRB_BSR.get((a, m), f"GxB_{a}_{m}_BOOL")
        emit("GrB_Semiring", name, f'{{{acode}, OP_{m}, TC_BOOL, "{name}"}}')

# ---- index-unary ops for select (reference graphblas/core/operator/indexunary.py, select.py: regex over dir(lib))
for op in ["TRIL", "TRIU", "DIAG", "OFFDIAG", "COLLE", "COLGT", "ROWLE", "ROWGT"]:
    emit("GrB_IndexUnaryOp", f"GrB_{op}", f'{{IOP_{op}, -1, "GrB_{op}"}}')
for t in TYPES:
    for op in ["VALUEEQ", "VALUENE", "VALUEGT", "VALUEGE", "VALUELT", "VALUELE"]:
        emit("GrB_IndexUnaryOp", f"GrB_{op}_{t}", f'{{IOP_{op}, TC_{t}, "GrB_{op}_{t}"}}')
for t in ["INT32", "INT64"]:
    for op in ["ROWINDEX", "COLINDEX", "DIAGINDEX"]:
        emit("GrB_IndexUnaryOp", f"GrB_{op}_{t}", f'{{IOP_{op}, -1, "GrB_{op}_{t}"}}')

# ---- descriptors  (reference graphblas/core/descriptor.py:51-84)
for r in (0, 1):
    for s in (0, 1):
        for c in (0, 1):
            for t0 in (0, 1):
                for t1 in (0, 1):
                    if not (r or s or c or t0 or t1):
                        continue
                    name = "GrB_DESC_" + ("R" if r else "") + ("S" if s else "") + ("C" if c else "") + \
                        ("T0" if t0 else "") + ("T1" if t1 else "")
                    emit("GrB_Descriptor", name,
                         "{%s, %s, %s, %s, %s, \"%s\"}" % tuple(["true" if x else "false" for x in (r, c, s, t0, t1)] + [name]))

print("// GENERATED by gen_builtins.py -- do not edit")
print("\n".join(out))
print("static const struct { const char *name; void *handle; } g_symbol_table[] = {")
for s in table:
    print(f'    {{"{s}", (void *)&obj_{s}}},')
print("};")
