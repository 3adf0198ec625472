// api.cu -- the three hot entry points and the typed C-API name wrappers.
//
//   GrB_mxm : reference graphblas/core/matrix.py:2319-2328 ; GrB_mxv : core/matrix.py:2252-2259 ;
//   GrB_vxm : core/vector.py:1368-1375.  Argument order [C, mask, accum, op, A, B, desc] as assembled at
//   core/base.py:496-503; descriptor bits as in core/descriptor.py:51-84.
//
// Dimension / domain errors are detected from metadata and returned synchronously (the reference relies
// on that: `expr.new(name="")  # raise now`, core/matrix.py:2260-2261); the output object is only
// replaced after the multiply succeeded, so a failed call leaves it untouched.
#include "grb_ops.cuh"

extern "C" const GrB_Index *GrB_ALL;

// Every builtin semiring this library exports has kernels: multiplies that return their operand type run as they are, comparison
// multiplies (GxB_LOR_GT_INT32 ...) run in the operand type with the comparison as 1 / 0 and the logical monoid mapped onto it
// (csrc/gen_builtins.py), the result cast by the write-back.
static bool semiring_supported(const GrB_Semiring op) { return op != nullptr; }


// shared tail of GrB_mxv / GrB_vxm: multiply (with the write-back fused into the kernel when possible), then write back
static GrB_Info mat_vec_common(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring op, GrB_Matrix A,
                               bool use_transpose, GrB_Vector u, bool flip, const GrB_Descriptor desc) {
    const bool comp = desc && desc->comp, structure = desc && desc->structure, replace = desc && desc->replace;
    const uint8_t *mbytes = nullptr;
    void *mtmp = nullptr;
    if (mask) {
        GRB_TRY(vector_ensure_arrays(mask));
        GRB_TRY(mask_effective_bytes(&mbytes, &mtmp, mask->present, mask->vals, mask->type, mask->n, structure, &w->err));
    }
    const PeerTargets *peer = g_peer.n > 0 ? &g_peer : nullptr;
    const bool needs_epi = mask != nullptr || accum != nullptr || comp || peer;
    // (a pow accumulator takes the separate write-back pass: the fused epilogue is compiled without pow, grb_ops.cuh binop<T, false>)
    const bool can_fuse = needs_epi && w->type == op->type &&
                          (!accum || (accum->type == w->type && accum->ztype == accum->type && accum->opcode != OP_POW && accum->opcode != OP_RPOW)) &&
                          (peer || opt_get_int("fuse_epilogue", 1) != 0);
    if (peer && !can_fuse) {
        dev_free(mtmp);
        return set_error(&w->err, GrB_NOT_IMPLEMENTED, "fused multiply + peer exchange needs output, semiring and accumulator of one type");
    }
    VecEpiHost epi{w->vals, w->present, mbytes, mask != nullptr, comp, replace, accum ? accum->opcode : OP_NONE, peer};
    void *tv = nullptr;
    uint8_t *tp = nullptr;
    int64_t tl = 0;
    bool fused = false;
    GrB_Info info = multiply_mat_vec_impl(&tv, &tp, &tl, op, A, use_transpose, u, flip, mbytes, comp, &w->err, can_fuse ? &epi : nullptr, &fused);
    dev_free(mtmp);
    GRB_TRY(info);
    if (peer && !fused) { dev_free(tv); dev_free(tp); return set_error(&w->err, GrB_PANIC, "fused peer exchange did not run"); }
    if (fused) {
        vector_take_arrays(w, tv, tp, -1);
        return GrB_SUCCESS;
    }
    return vector_write_back(w, tv, tp, op->type, mask, accum, desc, true);
}

extern "C" GrB_Info GrB_mxv(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring op,
                            const GrB_Matrix A, const GrB_Vector u, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(w)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "GrB_mxv: output vector is not initialised");
    if (!op || !valid(A) || !valid(u)) return set_error(&w->err, GrB_NULL_POINTER, "GrB_mxv: null or uninitialised argument");
    if (mask && !valid(mask)) return set_error(&w->err, GrB_UNINITIALIZED_OBJECT, "GrB_mxv: bad mask");
    if (!semiring_supported(op)) return set_error(&w->err, GrB_NOT_IMPLEMENTED, "semiring %s has no kernel in this backend", op->name);
    const bool t0 = desc && desc->t0;
    const int64_t out_len = t0 ? A->ncols : A->nrows, in_len = t0 ? A->nrows : A->ncols;
    if (w->n != out_len || u->n != in_len || (mask && mask->n != w->n))
        return set_error(&w->err, GrB_DIMENSION_MISMATCH, "GrB_mxv: w(%lld) = A(%lldx%lld%s) * u(%lld), mask(%lld)", (long long)w->n,
                         (long long)A->nrows, (long long)A->ncols, t0 ? ")'" : ")", (long long)u->n, (long long)(mask ? mask->n : w->n));
    return mat_vec_common(w, mask, accum, op, A, t0, u, /*flip=*/false, desc);
}

extern "C" GrB_Info GrB_vxm(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, const GrB_Semiring op,
                            const GrB_Vector u, const GrB_Matrix A, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(w)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "GrB_vxm: output vector is not initialised");
    if (!op || !valid(A) || !valid(u)) return set_error(&w->err, GrB_NULL_POINTER, "GrB_vxm: null or uninitialised argument");
    if (mask && !valid(mask)) return set_error(&w->err, GrB_UNINITIALIZED_OBJECT, "GrB_vxm: bad mask");
    if (!semiring_supported(op)) return set_error(&w->err, GrB_NOT_IMPLEMENTED, "semiring %s has no kernel in this backend", op->name);
    const bool t1 = desc && desc->t1;   // only INP1 (the matrix) can be transposed: core/vector.py:1374
    // w' = u' A  <=>  w = A' u with the multiply's operands swapped
    const int64_t out_len = t1 ? A->nrows : A->ncols, in_len = t1 ? A->ncols : A->nrows;
    if (w->n != out_len || u->n != in_len || (mask && mask->n != w->n))
        return set_error(&w->err, GrB_DIMENSION_MISMATCH, "GrB_vxm: w(%lld) = u(%lld) * A(%lldx%lld%s, mask(%lld)", (long long)w->n,
                         (long long)u->n, (long long)A->nrows, (long long)A->ncols, t1 ? ")'" : ")", (long long)(mask ? mask->n : w->n));
    return mat_vec_common(w, mask, accum, op, A, /*use_transpose=*/!t1, u, /*flip=*/true, desc);
}

extern "C" GrB_Info GrB_mxm(GrB_Matrix C, const GrB_Matrix Mask, const GrB_BinaryOp accum, const GrB_Semiring op,
                            const GrB_Matrix A, const GrB_Matrix B, const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(C)) return set_error(nullptr, GrB_UNINITIALIZED_OBJECT, "GrB_mxm: output matrix is not initialised");
    if (!op || !valid(A) || !valid(B)) return set_error(&C->err, GrB_NULL_POINTER, "GrB_mxm: null or uninitialised argument");
    if (Mask && !valid(Mask)) return set_error(&C->err, GrB_UNINITIALIZED_OBJECT, "GrB_mxm: bad mask");
    if (!semiring_supported(op)) return set_error(&C->err, GrB_NOT_IMPLEMENTED, "semiring %s has no kernel in this backend", op->name);
    const bool t0 = desc && desc->t0, t1 = desc && desc->t1;
    const int64_t m = t0 ? A->ncols : A->nrows, k = t0 ? A->nrows : A->ncols;
    const int64_t k2 = t1 ? B->ncols : B->nrows, n = t1 ? B->nrows : B->ncols;
    if (k != k2 || C->nrows != m || C->ncols != n || (Mask && (Mask->nrows != m || Mask->ncols != n)))
        return set_error(&C->err, GrB_DIMENSION_MISMATCH, "GrB_mxm: C(%lldx%lld) = A(%lldx%lld) * B(%lldx%lld)", (long long)C->nrows,
                         (long long)C->ncols, (long long)m, (long long)k, (long long)k2, (long long)n);
    GrB_Matrix T = nullptr;
    GRB_TRY(spgemm(&T, op, A, t0, B, t1, Mask, desc && desc->comp, desc && desc->structure, &C->err, false, nullptr, nullptr));
    GrB_Info info = matrix_write_back(C, T, Mask, accum, desc);
    GrB_Matrix_free(&T);
    return info;
}

extern "C" GrB_Info GrB_cuda_mxm_symbolic(GrB_Index *flops, GrB_Index *nvals_out, const GrB_Matrix A, const GrB_Matrix B,
                                          const GrB_Descriptor desc) {
    CHECK_INIT();
    if (!valid(A) || !valid(B)) return GrB_UNINITIALIZED_OBJECT;
    uint64_t f = 0, nv = 0;
    GrB_Matrix T = nullptr;
    GRB_TRY(spgemm(&T, nullptr, A, desc && desc->t0, B, desc && desc->t1, nullptr, false, false, &A->err, true, &f, &nv));
    if (flops) *flops = f;
    if (nvals_out) *nvals_out = nv;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ typed C-API names (forwarders)
#define GRB_TYPED(SFX, CT, TOBJ)                                                                                              \
    extern "C" GrB_Info GrB_Matrix_import_##SFX(GrB_Matrix *A, GrB_Type type, GrB_Index nrows, GrB_Index ncols,              \
                                                 const GrB_Index *Ap, const GrB_Index *Ai, const CT *Ax, GrB_Index Ap_len,     \
                                                 GrB_Index Ai_len, GrB_Index Ax_len, GrB_Format format) {                      \
        return GrB_cuda_Matrix_import(A, type ? type : TOBJ, TOBJ, nrows, ncols, Ap, Ai, Ax, Ap_len, Ai_len, Ax_len, format);                \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Matrix_export_##SFX(GrB_Index *Ap, GrB_Index *Ai, CT *Ax, GrB_Index *Ap_len, GrB_Index *Ai_len,  \
                                                 GrB_Index *Ax_len, GrB_Format format, GrB_Matrix A) {                         \
        return GrB_cuda_Matrix_export(Ap, Ai, Ax, TOBJ, Ap_len, Ai_len, Ax_len, format, A);                                   \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Matrix_build_##SFX(GrB_Matrix C, const GrB_Index *I, const GrB_Index *J, const CT *X,            \
                                                GrB_Index nvals, const GrB_BinaryOp dup) {                                     \
        return GrB_cuda_Matrix_build(C, I, J, X, TOBJ, nvals, dup);                                                           \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Matrix_extractTuples_##SFX(GrB_Index *I, GrB_Index *J, CT *X, GrB_Index *nvals,                  \
                                                        const GrB_Matrix A) {                                                  \
        return GrB_cuda_Matrix_extractTuples(I, J, X, TOBJ, nvals, A);                                                        \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Matrix_extractElement_##SFX(CT *x, const GrB_Matrix A, GrB_Index i, GrB_Index j) {               \
        return GrB_cuda_Matrix_extractElement(x, TOBJ, A, i, j);                                                              \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Vector_build_##SFX(GrB_Vector w, const GrB_Index *I, const CT *X, GrB_Index nvals,               \
                                                const GrB_BinaryOp dup) {                                                      \
        return GrB_cuda_Vector_build(w, I, X, TOBJ, nvals, dup);                                                              \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Vector_extractTuples_##SFX(GrB_Index *I, CT *X, GrB_Index *nvals, const GrB_Vector v) {          \
        return GrB_cuda_Vector_extractTuples(I, X, TOBJ, nvals, v);                                                           \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Vector_setElement_##SFX(GrB_Vector w, CT x, GrB_Index i) {                                       \
        return GrB_cuda_Vector_setElement(w, &x, TOBJ, i);                                                                    \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Vector_extractElement_##SFX(CT *x, const GrB_Vector v, GrB_Index i) {                            \
        return GrB_cuda_Vector_extractElement(x, TOBJ, v, i);                                                                 \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Vector_reduce_##SFX(CT *val, const GrB_BinaryOp accum, const GrB_Monoid op, const GrB_Vector u,  \
                                                 const GrB_Descriptor desc) {                                                  \
        (void)desc;                                                                                                           \
        return GrB_cuda_Vector_reduce(val, TOBJ, accum, op, u, nullptr);                                                      \
    }                                                                                                                         \
    extern "C" GrB_Info GrB_Vector_assign_##SFX(GrB_Vector w, const GrB_Vector mask, const GrB_BinaryOp accum, CT val,       \
                                                 const GrB_Index *indices, GrB_Index ni, const GrB_Descriptor desc) {          \
        if (indices != nullptr && indices != GrB_ALL)                                                                        \
            return GrB_NOT_IMPLEMENTED; /* only GrB_ALL is on the path */                                                     \
        return GrB_cuda_Vector_assign_scalar(w, mask, accum, &val, TOBJ, desc);                                               \
    }

static const GrB_Index grb_all_sentinel = 0;
extern "C" { const GrB_Index *GrB_ALL = &grb_all_sentinel; }

GRB_TYPED(BOOL, bool, GrB_BOOL)
GRB_TYPED(INT8, int8_t, GrB_INT8)
GRB_TYPED(INT16, int16_t, GrB_INT16)
GRB_TYPED(INT32, int32_t, GrB_INT32)
GRB_TYPED(INT64, int64_t, GrB_INT64)
GRB_TYPED(UINT8, uint8_t, GrB_UINT8)
GRB_TYPED(UINT16, uint16_t, GrB_UINT16)
GRB_TYPED(UINT32, uint32_t, GrB_UINT32)
GRB_TYPED(UINT64, uint64_t, GrB_UINT64)
GRB_TYPED(FP32, float, GrB_FP32)
GRB_TYPED(FP64, double, GrB_FP64)

// ------------------------------------------------------------------ fused multiply + exchange over peer memory
PeerTargets g_peer = {0, {nullptr}, {nullptr}, 0, nullptr};
extern "C" GrB_Info GrB_cuda_set_peer_targets(int n, void *const *vals, uint8_t *const *present, GrB_Index offset, const void *scale) {
    if (n < 0 || n > MAX_PEERS || (n > 0 && !vals)) return set_error(nullptr, GrB_INVALID_VALUE, "peer targets: 0..%d vectors", MAX_PEERS);
    g_peer.n = n;
    for (int k = 0; k < n; k++) { g_peer.vals[k] = vals[k]; g_peer.present[k] = present ? present[k] : nullptr; }
    g_peer.offset = (int64_t)offset;
    g_peer.scale = scale;
    return GrB_SUCCESS;
}
