// scalar.cu -- GrB_Scalar objects and the C-API names built on them or on a C scalar: reduce to a GrB_Scalar, the typed
// bind-1st / bind-2nd apply names, typed Matrix reduce, select, whole-object assign, Matrix_setElement.
//
// These are thin host-side adapters: every one of them forwards to a type-generic core in vecops.cu / io.cu (pointer + GrB_Type
// for the scalar), which is where the kernels live.  They exist because the reference formats these exact names at run time:
//   GrB_Scalar_*                         graphblas/core/scalar.py:83, 199-216, 235-280
//   GrB_{Vector,Matrix}_reduce_Monoid_Scalar   core/vector.py:1670, core/matrix.py:2750
//   GrB_{Vector,Matrix}_apply_BinaryOp{1st,2nd}_{T,Scalar}   core/vector.py:1477-1525, core/matrix.py:2472-2520
//   GrB_Matrix_reduce_<T>                core/matrix.py:2754
//   GrB_{Vector,Matrix}_select_{T,Scalar}      core/vector.py:1622-1624, core/matrix.py:2621-2623
//   GrB_Vector_assign / GrB_Matrix_assign      core/vector.py:1928, core/matrix.py:3300
#include <new>

#include "grb_ops.cuh"

extern "C" const GrB_Index *GrB_ALL;

// ------------------------------------------------------------------ host cast
namespace {
struct Wide { double d; int64_t l; uint64_t ul; bool isf, isu; };
Wide widen(const void *src, int st) {
    Wide w{0, 0, 0, false, false};
    switch (st) {
        case TC_BOOL: w.l = *(const uint8_t *)src != 0; break;
        case TC_INT8: w.l = *(const int8_t *)src; break;
        case TC_INT16: w.l = *(const int16_t *)src; break;
        case TC_INT32: w.l = *(const int32_t *)src; break;
        case TC_INT64: w.l = *(const int64_t *)src; break;
        case TC_UINT8: w.ul = *(const uint8_t *)src; w.isu = true; break;
        case TC_UINT16: w.ul = *(const uint16_t *)src; w.isu = true; break;
        case TC_UINT32: w.ul = *(const uint32_t *)src; w.isu = true; break;
        case TC_UINT64: w.ul = *(const uint64_t *)src; w.isu = true; break;
        case TC_FP32: w.d = *(const float *)src; w.isf = true; break;
        default: w.d = *(const double *)src; w.isf = true; break;
    }
    return w;
}
template <typename CT> void narrow(void *dst, const Wide &w) {
    CT x = w.isf ? (CT)w.d : w.isu ? (CT)w.ul : (CT)w.l;
    memcpy(dst, &x, sizeof x);
}
}   // namespace

void host_cast(void *dst, int dt, const void *src, int st) {
    const Wide w = widen(src, st);
    switch (dt) {
        case TC_BOOL: *(uint8_t *)dst = (w.isf ? w.d != 0 : w.isu ? w.ul != 0 : w.l != 0) ? 1 : 0; break;
        case TC_INT8: narrow<int8_t>(dst, w); break;
        case TC_INT16: narrow<int16_t>(dst, w); break;
        case TC_INT32: narrow<int32_t>(dst, w); break;
        case TC_INT64: narrow<int64_t>(dst, w); break;
        case TC_UINT8: narrow<uint8_t>(dst, w); break;
        case TC_UINT16: narrow<uint16_t>(dst, w); break;
        case TC_UINT32: narrow<uint32_t>(dst, w); break;
        case TC_UINT64: narrow<uint64_t>(dst, w); break;
        case TC_FP32: narrow<float>(dst, w); break;
        default: narrow<double>(dst, w); break;
    }
}

// ------------------------------------------------------------------ GrB_Scalar
extern "C" GrB_Info GrB_Scalar_new(GrB_Scalar *s, GrB_Type type) {   // host-only object: needs no device, hence no CHECK_INIT
    if (!s || !type) return set_error(nullptr, GrB_NULL_POINTER, "GrB_Scalar_new: null argument");
    GrB_Scalar x = new (std::nothrow) GrB_Scalar_opaque();
    if (!x) return GrB_OUT_OF_MEMORY;
    x->magic = GRB_MAGIC_SCALAR;
    x->type = type->code;
    x->has = false;
    memset(x->buf, 0, sizeof x->buf);
    *s = x;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Scalar_free(GrB_Scalar *s) {
    if (!s || !*s || !valid(*s)) return GrB_SUCCESS;
    (*s)->magic = GRB_MAGIC_FREED;
    delete *s;
    *s = nullptr;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Scalar_dup(GrB_Scalar *t, const GrB_Scalar s) {
    if (!t) return GrB_NULL_POINTER;
    if (!valid(s)) return GrB_UNINITIALIZED_OBJECT;
    GrB_Scalar x = new (std::nothrow) GrB_Scalar_opaque();
    if (!x) return GrB_OUT_OF_MEMORY;
    x->magic = GRB_MAGIC_SCALAR;
    x->type = s->type;
    x->has = s->has;
    memcpy(x->buf, s->buf, sizeof x->buf);
    *t = x;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Scalar_clear(GrB_Scalar s) {
    if (!valid(s)) return GrB_UNINITIALIZED_OBJECT;
    s->has = false;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Scalar_nvals(GrB_Index *nvals, const GrB_Scalar s) {
    if (!nvals) return GrB_NULL_POINTER;
    if (!valid(s)) return GrB_UNINITIALIZED_OBJECT;
    *nvals = s->has ? 1 : 0;
    return GrB_SUCCESS;
}
extern "C" GrB_Info GrB_Scalar_wait(GrB_Scalar s, GrB_WaitMode) { return valid(s) ? GrB_SUCCESS : GrB_UNINITIALIZED_OBJECT; }
extern "C" GrB_Info GrB_Scalar_error(const char **error, const GrB_Scalar s) {
    if (!error) return GrB_NULL_POINTER;
    *error = valid(s) ? s->err.c_str() : GrB_cuda_last_error();
    return GrB_SUCCESS;
}
static GrB_Info scalar_set(GrB_Scalar s, const void *x, int xtype) {
    if (!valid(s)) return GrB_UNINITIALIZED_OBJECT;
    host_cast(s->buf, s->type, x, xtype);
    s->has = true;
    return GrB_SUCCESS;
}
static GrB_Info scalar_get(void *x, int xtype, const GrB_Scalar s) {
    if (!x) return GrB_NULL_POINTER;
    if (!valid(s)) return GrB_UNINITIALIZED_OBJECT;
    if (!s->has) return GrB_NO_VALUE;
    host_cast(x, xtype, s->buf, s->type);
    return GrB_SUCCESS;
}

// s (accum)= t where t is a host value of type ttype, or nothing (has_t false)
template <typename CT> static void accum_into(GrB_Scalar s, int accum_op, const void *t, int ttype) {
    CT x;
    if constexpr (is_gbool<CT>::value) { uint8_t b; host_cast(&b, TC_BOOL, t, ttype); x.v = b; }
    else host_cast(&x, type_code_of<CT>(), t, ttype);
    if (accum_op != OP_NONE && s->has) {
        CT old;
        memcpy(&old, s->buf, sizeof(CT));
        x = binop<CT>(accum_op, old, x);
    }
    memcpy(s->buf, &x, sizeof(CT));
    s->has = true;
}
static GrB_Info scalar_store_result(GrB_Scalar s, const GrB_BinaryOp accum, const void *t, int ttype, bool has_t) {
    if (accum && accum->ztype != accum->type)
        return set_error(&s->err, GrB_DOMAIN_MISMATCH, "accumulator %s does not return its input type", accum->name);
    if (!has_t) {
        if (!accum) s->has = false;   // no accumulator: the (empty) result replaces the scalar
        return GrB_SUCCESS;
    }
    GRB_DISPATCH_TYPE(s->type, CT, accum_into<CT>(s, accum ? accum->opcode : OP_NONE, t, ttype));
    return GrB_SUCCESS;
}

extern "C" GrB_Info GrB_Vector_reduce_Monoid_Scalar(GrB_Scalar s, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Vector u,
                                                    const GrB_Descriptor) {
    CHECK_INIT();
    if (!valid(s)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "reduce: output scalar is not initialised");
    if (!op) return set_error(&s->err, GrB_NULL_POINTER, "reduce: null monoid");
    unsigned char t[8] = {0};
    GrB_Index nv = 0;
    GRB_TRY(GrB_cuda_Vector_reduce(t, const_cast<GrB_Type>(type_of_code(op->type)), nullptr, op, u, &nv));
    return scalar_store_result(s, accum, t, op->type, nv > 0);
}
extern "C" GrB_Info GrB_Matrix_reduce_Monoid_Scalar(GrB_Scalar s, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Matrix A,
                                                    const GrB_Descriptor) {
    CHECK_INIT();
    if (!valid(s)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "reduce: output scalar is not initialised");
    if (!op) return set_error(&s->err, GrB_NULL_POINTER, "reduce: null monoid");
    unsigned char t[8] = {0};
    GrB_Index nv = 0;
    GRB_TRY(GrB_cuda_Matrix_reduce(t, const_cast<GrB_Type>(type_of_code(op->type)), nullptr, op, A, &nv));
    return scalar_store_result(s, accum, t, op->type, nv > 0);
}

// ------------------------------------------------------------------ apply / select with a GrB_Scalar
#define SCALAR_ARG(sc, what)                                                                              \
    if (!valid(sc)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, what ": scalar is not initialised"); \
    if (!(sc)->has) return set_error(nullptr, GrB_EMPTY_OBJECT, what ": scalar is empty")
extern "C" GrB_Info GrB_Vector_apply_BinaryOp1st_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                        const GrB_Scalar x, const GrB_Vector u, const GrB_Descriptor desc) {
    SCALAR_ARG(x, "apply");
    return GrB_cuda_Vector_apply_binop(w, mask, accum, op, u, x->buf, const_cast<GrB_Type>(type_of_code(x->type)), 1, desc);
}
extern "C" GrB_Info GrB_Vector_apply_BinaryOp2nd_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                        const GrB_Vector u, const GrB_Scalar y, const GrB_Descriptor desc) {
    SCALAR_ARG(y, "apply");
    return GrB_cuda_Vector_apply_binop(w, mask, accum, op, u, y->buf, const_cast<GrB_Type>(type_of_code(y->type)), 0, desc);
}
extern "C" GrB_Info GrB_Matrix_apply_BinaryOp1st_Scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                        const GrB_Scalar x, const GrB_Matrix A, const GrB_Descriptor desc) {
    SCALAR_ARG(x, "apply");
    return GrB_cuda_Matrix_apply_binop(C, Mask, accum, op, A, x->buf, const_cast<GrB_Type>(type_of_code(x->type)), 1, desc);
}
extern "C" GrB_Info GrB_Matrix_apply_BinaryOp2nd_Scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_BinaryOp op,
                                                        const GrB_Matrix A, const GrB_Scalar y, const GrB_Descriptor desc) {
    SCALAR_ARG(y, "apply");
    return GrB_cuda_Matrix_apply_binop(C, Mask, accum, op, A, y->buf, const_cast<GrB_Type>(type_of_code(y->type)), 0, desc);
}
extern "C" GrB_Info GrB_Vector_select_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                             const GrB_Vector u, const GrB_Scalar y, const GrB_Descriptor desc) {
    SCALAR_ARG(y, "select");
    return GrB_cuda_Vector_select(w, mask, accum, op, u, y->buf, const_cast<GrB_Type>(type_of_code(y->type)), desc);
}
extern "C" GrB_Info GrB_Matrix_select_Scalar(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op,
                                             const GrB_Matrix A, const GrB_Scalar y, const GrB_Descriptor desc) {
    SCALAR_ARG(y, "select");
    return GrB_cuda_Matrix_select(C, Mask, accum, op, A, y->buf, const_cast<GrB_Type>(type_of_code(y->type)), desc);
}
extern "C" GrB_Info GrB_Vector_assign_Scalar(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Scalar s,
                                             const GrB_Index *indices, GrB_Index, const GrB_Descriptor desc) {
    if (!valid(s)) return GrB_UNINITIALIZED_OBJECT;
    if (indices != nullptr && indices != GrB_ALL) return GrB_NOT_IMPLEMENTED;   // only GrB_ALL is on the path
    if (s->has) return GrB_cuda_Vector_assign_scalar(w, mask, accum, s->buf, const_cast<GrB_Type>(type_of_code(s->type)), desc);
    // an empty scalar assigns "no entry": the same as assigning an empty vector
    if (!valid(w)) return GrB_UNINITIALIZED_OBJECT;
    GrB_Vector e = nullptr;
    GRB_TRY(GrB_Vector_new(&e, const_cast<GrB_Type>(type_of_code(w->type)), (GrB_Index)w->n));
    GrB_Info info = GrB_Vector_assign(w, mask, accum, e, GrB_ALL, (GrB_Index)w->n, desc);
    GrB_Vector_free(&e);
    return info;
}

// ------------------------------------------------------------------ Matrix_setElement: C(i,j) = x as C (second)= T, T the 1-entry matrix
static GrB_Info matrix_set_element(GrB_Matrix C, const void *x, GrB_Type xtype, GrB_Index i, GrB_Index j) {
    CHECK_INIT();
    if (!valid(C)) return GrB_UNINITIALIZED_OBJECT;
    if (i >= (GrB_Index)C->nrows || j >= (GrB_Index)C->ncols)
        return set_error(&C->err, GrB_INVALID_INDEX, "setElement: (%llu, %llu) outside %lld x %lld", (unsigned long long)i, (unsigned long long)j,
                         (long long)C->nrows, (long long)C->ncols);
    GrB_Matrix T = nullptr;
    GRB_TRY(matrix_new_shell(&T, C->type, C->nrows, C->ncols));
    GrB_Info info = GrB_cuda_Matrix_build(T, &i, &j, x, xtype, 1, nullptr);
    if (!info) {
        char name[32];
        snprintf(name, sizeof name, "GrB_SECOND_%s", type_of_code(C->type)->name + 4);
        const GrB_BinaryOp second = (GrB_BinaryOp)GrB_cuda_lookup(name);
        info = second ? matrix_write_back(C, T, nullptr, second, nullptr) : GrB_PANIC;
    }
    GrB_Matrix_free(&T);
    return info;
}

// ------------------------------------------------------------------ typed names
#define GRB_TYPED2(SFX, CT, TOBJ)                                                                                                        \
    extern "C" GrB_Info GrB_Scalar_setElement_##SFX(GrB_Scalar s, CT x) { return scalar_set(s, &x, TOBJ->code); }                        \
    extern "C" GrB_Info GrB_Scalar_extractElement_##SFX(CT *x, const GrB_Scalar s) { return scalar_get(x, TOBJ->code, s); }              \
    extern "C" GrB_Info GrB_Matrix_setElement_##SFX(GrB_Matrix C, CT x, GrB_Index i, GrB_Index j) {                                      \
        return matrix_set_element(C, &x, TOBJ, i, j);                                                                                    \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Matrix_reduce_##SFX(CT *val, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Matrix A,              \
                                                const GrB_Descriptor) {                                                                  \
        return GrB_cuda_Matrix_reduce(val, TOBJ, accum, op, A, nullptr);                                                                 \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Vector_apply_BinaryOp1st_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,                \
                                                           const GrB_BinaryOp op, CT x, const GrB_Vector u, const GrB_Descriptor desc) { \
        return GrB_cuda_Vector_apply_binop(w, mask, accum, op, u, &x, TOBJ, 1, desc);                                                    \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Vector_apply_BinaryOp2nd_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum,                \
                                                           const GrB_BinaryOp op, const GrB_Vector u, CT y, const GrB_Descriptor desc) { \
        return GrB_cuda_Vector_apply_binop(w, mask, accum, op, u, &y, TOBJ, 0, desc);                                                    \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Matrix_apply_BinaryOp1st_##SFX(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum,                \
                                                           const GrB_BinaryOp op, CT x, const GrB_Matrix A, const GrB_Descriptor desc) { \
        return GrB_cuda_Matrix_apply_binop(C, Mask, accum, op, A, &x, TOBJ, 1, desc);                                                    \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Matrix_apply_BinaryOp2nd_##SFX(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum,                \
                                                           const GrB_BinaryOp op, const GrB_Matrix A, CT y, const GrB_Descriptor desc) { \
        return GrB_cuda_Matrix_apply_binop(C, Mask, accum, op, A, &y, TOBJ, 0, desc);                                                    \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Vector_select_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op, \
                                                const GrB_Vector u, CT y, const GrB_Descriptor desc) {                                   \
        return GrB_cuda_Vector_select(w, mask, accum, op, u, &y, TOBJ, desc);                                                            \
    }                                                                                                                                    \
    extern "C" GrB_Info GrB_Matrix_select_##SFX(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_IndexUnaryOp op, \
                                                const GrB_Matrix A, CT y, const GrB_Descriptor desc) {                                   \
        return GrB_cuda_Matrix_select(C, Mask, accum, op, A, &y, TOBJ, desc);                                                            \
    }

GRB_TYPED2(BOOL, bool, GrB_BOOL)
GRB_TYPED2(INT8, int8_t, GrB_INT8)
GRB_TYPED2(INT16, int16_t, GrB_INT16)
GRB_TYPED2(INT32, int32_t, GrB_INT32)
GRB_TYPED2(INT64, int64_t, GrB_INT64)
GRB_TYPED2(UINT8, uint8_t, GrB_UINT8)
GRB_TYPED2(UINT16, uint16_t, GrB_UINT16)
GRB_TYPED2(UINT32, uint32_t, GrB_UINT32)
GRB_TYPED2(UINT64, uint64_t, GrB_UINT64)
GRB_TYPED2(FP32, float, GrB_FP32)
GRB_TYPED2(FP64, double, GrB_FP64)
