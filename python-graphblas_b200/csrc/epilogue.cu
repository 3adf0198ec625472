// epilogue.cu -- the GraphBLAS write-back  C<M, replace> (accum)= T  for vectors and matrices.
//
//   Z = accum ? (C u T with accum on the intersection) : T
//   m(i) = mask entry counts (structure: present; value: present and truthy), flipped by COMP
//   m  -> C(i) := Z(i) (deleted when Z(i) is absent);  !m -> REPLACE ? delete : keep
//
// Semantics block of SURVEY.md section 8(c); flags decoded as the reference does at
// graphblas/core/base.py:458-475 (mask.complement / mask.structure / replace -> GrB_DESC_*).
#include "grb_ops.cuh"

// ------------------------------------------------------------------ mask -> bytes
template <typename T>
__global__ void mask_truthy_kernel(const uint8_t *__restrict__ present, const T *__restrict__ vals, int64_t n,
                                   uint8_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = (present ? present[i] != 0 : true) && truthy<T>(vals[i]);
}

// byte i != 0  <=>  mask entry i "counts" before complement.  *tmp must be dev_free'd by the caller.
GrB_Info mask_effective_bytes(const uint8_t **out, void **tmp, const uint8_t *present, const void *vals, int type,
                              int64_t n, bool structure, std::string *err) {
    *tmp = nullptr;
    if (structure || n == 0) {
        *out = present;
        return GrB_SUCCESS;
    }
    uint8_t *eff = (uint8_t *)dev_alloc((size_t)n);
    if (!eff) return set_error(err, GrB_OUT_OF_MEMORY, "mask bytes");
    int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)g_num_sms * 16);
    LAUNCH_NOTE("mask_truthy");
    GRB_DISPATCH_TYPE(type, T, (mask_truthy_kernel<T><<<blocks, 256, 0, g_stream>>>(present, (const T *)vals, n, eff)));
    CUDA_TRY(err, cudaGetLastError());
    *tmp = eff;
    *out = eff;
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ vector write-back (in place on T's buffers)
template <typename T>
__global__ void vec_epilogue_kernel(int64_t n, const T *__restrict__ c_vals, const uint8_t *__restrict__ c_present,
                                    T *t_vals, uint8_t *t_present, const uint8_t *__restrict__ mask, bool has_mask,
                                    bool comp, bool replace, int accum) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        bool m = has_mask ? ((mask[i] != 0) != comp) : !comp;
        bool cp = c_present ? c_present[i] != 0 : false;
        bool tp = t_present[i] != 0;
        T c = cp ? c_vals[i] : T();
        T z = t_vals[i];
        bool zp = tp;
        if (accum != OP_NONE && cp) {
            z = tp ? binop<T>(accum, c, z) : c;
            zp = true;
        }
        if (!m) {
            if (replace) zp = false;
            else { z = c; zp = cp; }
        }
        t_vals[i] = zp ? z : T();
        t_present[i] = zp ? 1 : 0;
    }
}

GrB_Info vector_write_back(GrB_Vector w, void *t_vals, uint8_t *t_present, int t_type, const GrB_Vector mask,
                           const GrB_BinaryOp accum, const GrB_Descriptor desc, bool t_owned) {
    (void)t_owned;
    const bool comp = desc && desc->comp, structure = desc && desc->structure, replace = desc && desc->replace;
    const int64_t n = w->n;
    // T into the output's type
    if (t_type != w->type) {
        void *cast = dev_alloc((size_t)(n > 0 ? n : 1) * type_size(w->type));
        if (!cast) { dev_free(t_vals); dev_free(t_present); return set_error(&w->err, GrB_OUT_OF_MEMORY, "typecast of result"); }
        GrB_Info info = cast_array(cast, w->type, t_vals, t_type, n, &w->err);
        dev_free(t_vals);
        t_vals = cast;
        if (info) { dev_free(t_vals); dev_free(t_present); return info; }
    }
    if (accum && accum->ztype != accum->type) {
        dev_free(t_vals); dev_free(t_present);
        return set_error(&w->err, GrB_DOMAIN_MISMATCH, "accumulator %s does not return its input type", accum->name);
    }
    if (accum && accum->type != w->type) {
        dev_free(t_vals); dev_free(t_present);
        return set_error(&w->err, GrB_NOT_IMPLEMENTED, "accumulator %s must be typed like the output (%s)", accum->name,
                         type_of_code(w->type)->name);
    }
    const bool has_mask = mask != nullptr;
    if (!has_mask && !accum && !comp) {   // C = T
        vector_take_arrays(w, t_vals, t_present, -1);
        return GrB_SUCCESS;
    }
    const uint8_t *mbytes = nullptr;
    void *mtmp = nullptr;
    if (has_mask) {
        GrB_Info info = vector_ensure_arrays(mask);
        if (!info) info = mask_effective_bytes(&mbytes, &mtmp, mask->present, mask->vals, mask->type, n, structure, &w->err);
        if (info) { dev_free(t_vals); dev_free(t_present); return info; }
    }
    if (n > 0) {
        int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)g_num_sms * 16);
        LAUNCH_NOTE("vec_epilogue");
        GRB_DISPATCH_TYPE(w->type, T,
                          (vec_epilogue_kernel<T><<<blocks, 256, 0, g_stream>>>(n, (const T *)w->vals, w->present, (T *)t_vals,
                                                                            t_present, mbytes, has_mask, comp, replace,
                                                                            accum ? accum->opcode : OP_NONE)));
    }
    cudaError_t e = cudaGetLastError();
    dev_free(mtmp);
    if (e != cudaSuccess) { dev_free(t_vals); dev_free(t_present); return cuda_fail(&w->err, e, "vec_epilogue"); }
    vector_take_arrays(w, t_vals, t_present, -1);
    return GrB_SUCCESS;
}

// ------------------------------------------------------------------ matrix write-back (two passes over sorted rows)
template <typename T, bool FILL>
__global__ void mat_epilogue_kernel(int64_t nrows, const int64_t *__restrict__ cp, const int32_t *__restrict__ cj,
                                    const T *__restrict__ cx, const int64_t *__restrict__ tp,
                                    const int32_t *__restrict__ tj, const T *__restrict__ tx,
                                    const int64_t *__restrict__ mp, const int32_t *__restrict__ mj,
                                    const uint8_t *__restrict__ meff, bool has_mask, bool comp, bool replace, int accum,
                                    int64_t *__restrict__ out_ptr, int32_t *__restrict__ out_j, T *__restrict__ out_x) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    int64_t a = cp ? cp[i] : 0, ae = cp ? cp[i + 1] : 0;
    int64_t b = tp[i], be = tp[i + 1];
    int64_t q = has_mask ? mp[i] : 0, qe = has_mask ? mp[i + 1] : 0;
    int64_t o = FILL ? out_ptr[i] : 0;
    int64_t cnt = 0;
    while (a < ae || b < be) {
        int32_t ja = a < ae ? cj[a] : INT32_MAX, jb = b < be ? tj[b] : INT32_MAX;
        int32_t j = ja < jb ? ja : jb;
        bool hc = ja == j, ht = jb == j;
        bool m;
        if (has_mask) {
            while (q < qe && mj[q] < j) q++;
            m = (q < qe && mj[q] == j && (!meff || meff[q] != 0)) != comp;
        } else {
            m = !comp;
        }
        T z = T();
        bool zp = false;
        if (ht) { z = tx[b]; zp = true; }
        if (accum != OP_NONE && hc) {
            z = ht ? binop<T>(accum, cx[a], z) : cx[a];
            zp = true;
        }
        if (!m) {
            if (replace) zp = false;
            else if (hc) { z = cx[a]; zp = true; }
            else zp = false;
        }
        if (zp) {
            if (FILL) { out_j[o + cnt] = j; out_x[o + cnt] = z; }
            cnt++;
        }
        if (hc) a++;
        if (ht) b++;
    }
    if (!FILL) out_ptr[i] = cnt;
}

GrB_Info matrix_write_back(GrB_Matrix C, GrB_Matrix T, const GrB_Matrix M, const GrB_BinaryOp accum,
                           const GrB_Descriptor desc) {
    const bool comp = desc && desc->comp, structure = desc && desc->structure, replace = desc && desc->replace;
    if (accum && accum->ztype != accum->type)
        return set_error(&C->err, GrB_DOMAIN_MISMATCH, "accumulator %s does not return its input type", accum->name);
    if (accum && accum->type != C->type)
        return set_error(&C->err, GrB_NOT_IMPLEMENTED, "accumulator %s must be typed like the output (%s)", accum->name,
                         type_of_code(C->type)->name);
    if (T->csr.end && (T->type != C->type || M || accum || comp)) GRB_TRY(matrix_materialize(T));
    if (T->type != C->type && T->nvals > 0) {
        void *cast = dev_alloc((size_t)T->nvals * type_size(C->type));
        if (!cast) return set_error(&C->err, GrB_OUT_OF_MEMORY, "typecast of result");
        GrB_Info info = cast_array(cast, C->type, T->csr.val, T->type, T->nvals, &C->err);
        if (info) { dev_free(cast); return info; }
        dev_free(T->csr.val);
        T->csr.val = cast;
    }
    T->type = C->type;
    if (!M && !accum && !comp) {
        matrix_take(C, T);
        return GrB_SUCCESS;
    }
    GRB_TRY(matrix_ensure_sorted(T));
    GRB_TRY(matrix_ensure_sorted(C));
    const uint8_t *meff = nullptr;
    void *mtmp = nullptr;
    if (M) {
        GRB_TRY(matrix_materialize(M));
        GRB_TRY(matrix_ensure_sorted(M));
        if (!structure && M->nvals > 0)
            GRB_TRY(mask_effective_bytes(&meff, &mtmp, nullptr, M->csr.val, M->type, M->nvals, false, &C->err));
    }
    const int64_t nrows = C->nrows;
    GrB_Matrix R = nullptr;
    {   // every exit below releases the mask bytes and the shell (round-1 advisor finding: early returns leaked them)
        GrB_Info ri = matrix_new_shell(&R, C->type, C->nrows, C->ncols);
        if (ri) { dev_free(mtmp); return ri; }
    }
    int64_t *optr = dev_alloc_t<int64_t>((size_t)nrows + 1);
    if (!optr) { dev_free(mtmp); GrB_Matrix_free(&R); return set_error(&C->err, GrB_OUT_OF_MEMORY, "write-back row pointers"); }
    GrB_Info info = GrB_SUCCESS;
    cudaMemsetAsync(optr + nrows, 0, sizeof(int64_t), g_stream);
    unsigned blocks = (unsigned)((nrows + 127) / 128);
    if (nrows > 0) {
        LAUNCH_NOTE("mat_epilogue_count");
        GRB_DISPATCH_TYPE(C->type, V,
                          (mat_epilogue_kernel<V, false><<<blocks, 128, 0, g_stream>>>(
                              nrows, C->csr.ptr, C->csr.idx, (const V *)C->csr.val, T->csr.ptr, T->csr.idx, (const V *)T->csr.val,
                              M ? M->csr.ptr : nullptr, M ? M->csr.idx : nullptr, meff, M != nullptr, comp, replace,
                              accum ? accum->opcode : OP_NONE, optr, nullptr, nullptr)));
    }
    info = exclusive_scan_i64(optr, nrows + 1, &C->err);
    int64_t total = info ? 0 : read_i64(optr + nrows);
    if (!info) {
        R->csr.ptr = optr;
        size_t nv = (size_t)(total > 0 ? total : 1);
        R->csr.idx = dev_alloc_t<int32_t>(nv);
        R->csr.val = dev_alloc(nv * type_size(C->type));
        R->nvals = total;
        if (!R->csr.idx || !R->csr.val) info = set_error(&C->err, GrB_OUT_OF_MEMORY, "write-back result (%lld entries)", (long long)total);
    }
    if (!info && nrows > 0) {
        LAUNCH_NOTE("mat_epilogue_fill");
        GRB_DISPATCH_TYPE(C->type, V,
                          (mat_epilogue_kernel<V, true><<<blocks, 128, 0, g_stream>>>(
                              nrows, C->csr.ptr, C->csr.idx, (const V *)C->csr.val, T->csr.ptr, T->csr.idx, (const V *)T->csr.val,
                              M ? M->csr.ptr : nullptr, M ? M->csr.idx : nullptr, meff, M != nullptr, comp, replace,
                              accum ? accum->opcode : OP_NONE, optr, R->csr.idx, (V *)R->csr.val)));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) info = cuda_fail(&C->err, e, "mat_epilogue");
    }
    dev_free(mtmp);
    if (info) {
        if (!R->csr.ptr) dev_free(optr);
        GrB_Matrix_free(&R);
        return info;
    }
    matrix_take(C, R);
    GrB_Matrix_free(&R);
    return GrB_SUCCESS;
}
